// tcgen05 complex transform along a leading axis:  Y[o, j, i] = sum_h Mat[j, h] * X[o, h, i]   (all complex64).
//
// The matrix is the same for every plane o and every trailing index i, so the problem is ONE real GEMM
//     D[r, n] = sum_k' A[r, k'] * B[k', n]       r = (o, i)   k' = (h, re|im)   n = (j, re|im)
// whose rows are gathered from X (row r reads X[o, h, i] for all h: 8-byte elements at a stride of I complex
// numbers) and whose B is the real 2x2 expansion of Mat,  [[Re, Im], [-Im, Re]]  per (h, j), built once per plan
// matrix on the host (tf32 hi/lo split, chunk-major UMMA K-major interleave image, backend_cuda.cu
// tc_get_mid_image).  The accumulator row of a thread then holds (re, im) pairs of consecutive j, and the 32 lanes of
// a warp hold 32 consecutive rows = consecutive i of one plane: every store instruction writes 256 contiguous bytes.
//
// Pipeline = tc_kpipe.cuh (persistent CTAs, S-stage ring, 8 loader warps with a 4-deep register prefetch ring, one
// elected MMA lane issuing hi*hi + hi*lo + lo*hi, two TMEM accumulators, 4 epilogue warps); only the loaders'
// addressing and the epilogue differ.  Output columns beyond 256 are split over blockIdx.y (each tile re-streams the
// small, L2-resident input).
//
// STATUS: opt-in (UNO_B200_MID_TC=1) until it has been run against the SIMT kernel on a B200; the default leading-axis
// transform is mid2_kernel (backend_cuda.cu).
#pragma once
#include "backend.h"
#include "tc_common.cuh"
#include "tc_kpipe.cuh"

namespace uno {
namespace tc {

struct MidTcParams {
    const float* X;        // complex64 interleaved [O, H, I]
    float* Y;              // complex64 interleaved [O, J, I]
    const float* Bimg;     // [n_tiles][n_chunks][hi | lo], each half (KC/4) x N_t x 16 bytes
    int O, H, J, I;
    long R;                // O * I rows
    int K;                 // 2 * H
    int N_t, n_tiles, n_chunks, stages;
    long m_tiles;
    int tmem_cols;
};

// LW = loader warps (8 or 16): LW / 4 groups of 128 threads share the eight h pairs of a chunk
constexpr int kMidLoadWarps = 16;
constexpr int kMidThreads = (kMidLoadWarps + kKpEpiWarps + 2) * 32;     // + MMA issuer + B-chunk copy warp
__global__ void __launch_bounds__(kMidThreads, 1) mid_tc_kernel(const MidTcParams p) {
    constexpr int kKpLoadWarps = kMidLoadWarps;          // shadows the analysis kernel's constant
    constexpr int G = kMidLoadWarps / 4, JP = 8 / G;     // thread groups; h pairs per thread and chunk
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    const int nt = blockIdx.y;
    const uint32_t b_half = (uint32_t)p.N_t * kKC * 4;
    const uint32_t stage_bytes = 2 * kKpAHalf + 2 * b_half;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;            // [S] loaders (+ bulk copy bytes) -> mma
    uint64_t* empty = bars + 8;       // [S] mma -> loaders
    uint64_t* d_full = bars + 16;     // [2]
    uint64_t* d_empty = bars + 18;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    constexpr int kMmaWarp = kKpLoadWarps + kKpEpiWarps, kCopyWarp = kMmaWarp + 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], kKpLoadWarps * 32 + 1);   // +1: the arrive.expect_tx of the B copy
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&d_full[s], 1);
            mbar_init(&d_empty[s], kKpEpiWarps * 32);
        }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t buf_cols = (uint32_t)p.tmem_cols / 2;
    const int NKC = p.n_chunks;

    if (warp == kMmaWarp) {
        // ------------------------------------------------------------------ MMA issuer (as tc_kpipe.cuh)
        const uint32_t idesc = make_idesc_tf32(128, p.N_t, 0, 0);
        const uint32_t lbo_b = (uint32_t)p.N_t * 16;
        const uint32_t smem_base = smem_u32(smem);
        const uint64_t a_hi0 = make_smem_desc(smem_base, kLboA, 128);
        const uint64_t b_hi0 = make_smem_desc(smem_base + 2 * kKpAHalf, lbo_b, 128);
        const uint32_t a_lo_off = kKpAHalf >> 4, b_lo_off = b_half >> 4;     // descriptor units (16 bytes)
        const uint32_t a_step = (2 * kLboA) >> 4, b_step = (2 * lbo_b) >> 4, stage_step = stage_bytes >> 4;
        const int last_nks = (p.K - (NKC - 1) * kKC + 7) / 8;               // k-steps of the ragged last chunk
        int s = 0;
        uint32_t ph = 0;
        uint64_t a_st = a_hi0, b_st = b_hi0;
        int it = 0;
        for (long tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait(&d_empty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)buf * buf_cols;
            uint32_t acc = 0;
            for (int kc = 0; kc < NKC; ++kc) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const int nks = (kc == NKC - 1) ? last_nks : kKC / 8;
                uint64_t da = a_st, db = b_st;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < kKC / 8; ++ks) {
                        if (ks < nks) {
                            mma_tf32(d_tmem, da, db, idesc, ks ? 1u : acc);
                            mma_tf32(d_tmem, da, db + b_lo_off, idesc, 1u);
                            mma_tf32(d_tmem, da + a_lo_off, db, idesc, 1u);
                        }
                        da += a_step;
                        db += b_step;
                    }
                    tc_commit(&empty[s]);
                }
                acc = 1u;
                __syncwarp();
                a_st += stage_step;
                b_st += stage_step;
                if (++s == S) { s = 0; ph ^= 1u; a_st = a_hi0; b_st = b_hi0; }
            }
            if (elect_one()) tc_commit(&d_full[buf]);
            __syncwarp();
        }
    } else if (warp < kKpLoadWarps) {
        // ------------------------------------------------------------------ loaders: chunk = 128 rows x 16 h (32 real k)
        // thread -> row (ltid % 128) of the tile, h pairs hh + 2j (j = 0..3): two 8-byte loads make the 16 bytes
        // (re, im, re', im') of one K-major core-matrix row, so a warp's stores cover 512 contiguous bytes.
        const int ltid = threadIdx.x;
        const int rl = ltid & 127, hh = ltid >> 7;
        long n_my_tiles = 0;
        for (long tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) ++n_my_tiles;
        const long total = n_my_tiles * NKC;
        long i_tile = blockIdx.x;
        int i_kc = 0;
        int p_s = 0;
        uint32_t p_ph = 0;
        const long h_stride = 2L * p.I;                      // floats between X[o, h, i] and X[o, h+1, i]
        const uint32_t so = (uint32_t)hh * kLboA + (uint32_t)rl * 16;
        const float* row_ptr = nullptr;                      // &X[o, 0, i] of this thread's row in tile i_tile (null: row >= R)
        auto seek_row = [&]() {
            const long r = i_tile * 128 + rl;
            if (i_tile < p.m_tiles && r < p.R) {
                const long o = r / p.I;
                const long i = r - o * p.I;
                row_ptr = p.X + 2 * (o * (long)p.H * p.I + i);
            } else row_ptr = nullptr;
        };
        seek_row();
        float4 ring[kKpDepth][JP];
        auto issue = [&](float4 (&v)[JP]) {
            const int h0 = i_kc * (kKC / 2) + 2 * hh;
#pragma unroll
            for (int j = 0; j < JP; ++j) {
                const int h = h0 + 2 * G * j;
                float2 e0 = make_float2(0.f, 0.f), e1 = make_float2(0.f, 0.f);
                if (row_ptr) {
                    const float* q = row_ptr + (long)h * h_stride;
                    if (h < p.H) e0 = __ldg(reinterpret_cast<const float2*>(q));
                    if (h + 1 < p.H) e1 = __ldg(reinterpret_cast<const float2*>(q + h_stride));
                }
                v[j] = make_float4(e0.x, e0.y, e1.x, e1.y);
            }
            if (++i_kc == NKC) { i_kc = 0; i_tile += gridDim.x; seek_row(); }
        };
        auto process = [&](const float4 (&v)[JP]) {
            mbar_wait(&empty[p_s], p_ph ^ 1u);
            uint8_t* st = smem + (size_t)p_s * stage_bytes + so;
#pragma unroll
            for (int j = 0; j < JP; ++j) {
                float4 hi, lo;
                split_tf32(v[j].x, hi.x, lo.x);
                split_tf32(v[j].y, hi.y, lo.y);
                split_tf32(v[j].z, hi.z, lo.z);
                split_tf32(v[j].w, hi.w, lo.w);
                *reinterpret_cast<float4*>(st + (uint32_t)(G * j) * kLboA) = hi;
                *reinterpret_cast<float4*>(st + kKpAHalf + (uint32_t)(G * j) * kLboA) = lo;
            }
            fence_proxy_async();
            mbar_arrive(&full[p_s]);
            if (++p_s == S) { p_s = 0; p_ph ^= 1u; }
        };
#pragma unroll
        for (int d = 0; d < kKpDepth - 1; ++d)
            if (d < total) issue(ring[d]);
        for (long g = 0; g < total; g += kKpDepth) {
#pragma unroll
            for (int d = 0; d < kKpDepth; ++d) {
                if (g + d + kKpDepth - 1 < total) issue(ring[(d + kKpDepth - 1) % kKpDepth]);
                if (g + d < total) process(ring[d]);
            }
        }
    } else if (warp == kCopyWarp) {
        // ------------------------------------------------------------------ B-chunk copies (as tc_kpipe.cuh): one lane issues, per
        // chunk, the bulk copy of the matrix image into the stage the MMAs have released
        if (lane == 0) {
            long n_my_tiles = 0;
            for (long tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) ++n_my_tiles;
            const long total = n_my_tiles * NKC;
            const uint32_t img_chunk_floats = 2 * b_half / 4;
            const float* bimg = p.Bimg + (size_t)nt * NKC * img_chunk_floats;
            int s = 0, kc = 0;
            uint32_t ph = 0;
            for (long g = 0; g < total; ++g) {
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full[s], 2 * b_half);
                bulk_g2s(smem + (size_t)s * stage_bytes + 2 * kKpAHalf, bimg + (size_t)kc * img_chunk_floats, 2 * b_half, &full[s]);
                if (++s == S) { s = 0; ph ^= 1u; }
                if (++kc == NKC) kc = 0;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: one warp per TMEM lane quarter
        // thread = one row (o, i); 32 accumulator columns per round = 16 complex outputs Y[o, j0 .. j0+15, i]
        const int q = warp - kKpLoadWarps;
        const long j_stride = 2L * p.I;
        const int jt0 = nt * (p.N_t / 2);                 // first output index j of this column tile
        const int jn = min(p.N_t / 2, p.J - jt0);         // valid j in this tile
        int it = 0;
        for (long tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait_relaxed(&d_full[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_base = tmem_base + (uint32_t)buf * buf_cols + ((uint32_t)(q * 32) << 16);
            const long r = tile * 128 + q * 32 + lane;
            const bool rok = r < p.R;
            const long o = rok ? r / p.I : 0;
            const long i = rok ? r - o * p.I : 0;
            float* yrow = p.Y + 2 * ((o * p.J + jt0) * (long)p.I + i);
            for (int c0 = 0; c0 < 2 * jn; c0 += 32) {
                uint32_t v[2][16];
                tmem_ld_32x32b_x16(t_base + (uint32_t)c0, v[0]);
                const bool second = c0 + 16 < p.N_t;      // N_t is a multiple of 16: never read beyond the accumulator
                if (second) tmem_ld_32x32b_x16(t_base + (uint32_t)(c0 + 16), v[1]);
                tmem_ld_wait();
                if (rok) {
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        const int j = c0 / 2 + t;
                        if (j < jn && (t < 8 || second))
                            *reinterpret_cast<float2*>(yrow + (long)j * j_stride) =
                                make_float2(__uint_as_float(v[t >> 3][2 * (t & 7)]), __uint_as_float(v[t >> 3][2 * (t & 7) + 1]));
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&d_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace tc
}  // namespace uno
