// tcgen05 1x1 channel mix:  Y[b, n, p] = sum_k Wm[n, k] * X[b, k, p] (+ bias[n]),  p = pixel (contiguous).
//
// Used for Conv{2,3}d(k=1) forward (Wm = weight [Co,Ci]) and for its input gradient (Wm = weight^T).
// UMMA view: D[128 pixels, N] = A[128 pixels, K] * B[K, N] with
//   A = X_b^T : pixel index contiguous in memory  -> "MN-major" operand (no transpose pass needed)
//   B = Wm    : [N, K] row-major                  -> K-major operand, tf32 hi/lo image resident in smem
// The weights are parameters (they change every optimiser step), so their image is rebuilt on the fly by
// conv_weight_image_kernel (a few KB) right before the main kernel.
//
//   * persistent CTAs, tiles = (sample, 128-pixel block), static round-robin
//   * 16 loader warps stream 32-channel chunks of the tile (float4 along pixels, hi/lo split in registers,
//     one 16-byte st.shared per value group into the MN-major interleave layout; register ping-pong keeps
//     two chunks in flight)
//   * one lane issues 3 MMAs per 8-channel k-step into one of two TMEM accumulators
//   * 8 epilogue warps: tcgen05.ld 32x32b (lane = pixel) -> + bias -> stores that are 128 B contiguous per warp
#pragma once
#include "backend.h"
#include "tc_common.cuh"
#include "tc_kpipe.cuh"

namespace uno {
namespace tc {

struct ConvTcParams {
    const float* X; long sXb; long npix;
    const float* Bimg;         // [hi | lo], each (K_pad/4) x N_t x 16 bytes
    const float* bias;         // [N] or null
    float* Y; long sYb;
    int K, N, N_t, n_chunks, stages, batch;
    long tiles_per_b, n_tiles;
    int tmem_cols;
};

// MN-major tf32 operands must use the "128-byte swizzle with 32-byte base" canonical layout:
//   atom = 32 pixels (128 B, contiguous) x 4 channels (rows 128 B apart); inside an atom the 32-byte granule g of
//   channel row r sits at granule g ^ r.  Channel groups of 4 follow at SBO, pixel blocks of 32 at LBO.
constexpr uint32_t kCvSbo = 512;                        // next group of 4 channels
constexpr uint32_t kCvLbo = (kKC / 4) * kCvSbo;         // next block of 32 pixels (32 channels per stage)
constexpr uint32_t kCvAHalf = 4 * kCvLbo;               // one A image (hi or lo) per stage: 32 channels x 128 pixels
constexpr uint32_t kCvLayout = 1;                       // UMMA::LayoutType::SWIZZLE_128B_BASE32B
constexpr int kCvEpiWarps = 8;
__host__ __device__ constexpr int cv_threads(int LW) { return (LW + kCvEpiWarps + 1) * 32; }

__host__ __device__ inline size_t conv_tc_smem_bytes(int N_t, int n_chunks, int stages) {
    return 1024 + (size_t)2 * N_t * n_chunks * kKC * 4 + (size_t)stages * 2 * kCvAHalf + 32 * 8 + 16;
}

// Wm(n, k) = W[n * w_rs + k * w_cs]  ->  resident K-major image (hi | lo), zero padded to N_t x K_pad
__global__ void conv_weight_image_kernel(const float* __restrict__ W, long w_rs, long w_cs, int N, int K, int N_t, int K_pad,
                                         float* __restrict__ img) {
    const int total = N_t * K_pad;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int n = idx / K_pad, k = idx - n * K_pad;
        float v = 0.f;
        if (n < N && k < K) v = W[n * w_rs + k * w_cs];
        // round hi to nearest tf32 (instead of truncating): the weights get the better split
        const uint32_t u = __float_as_uint(v);
        const float hi = __uint_as_float((u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u);
        const float lo = v - hi;
        const size_t o = (size_t)(k / 4) * N_t * 4 + (size_t)n * 4 + (k % 4);
        img[o] = hi;
        img[(size_t)N_t * K_pad + o] = lo;
    }
}

// LW = loader warps (8 or 16): a thread owns 32 / LW channels of every chunk
template <int LW>
__global__ void __launch_bounds__(cv_threads(LW), 1) conv1x1_tc_kernel(const ConvTcParams p) {
    constexpr int kCvLoadWarps = LW;
    constexpr int CPT = 32 / LW;                   // channels per thread and chunk
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    const int K_pad = p.n_chunks * kKC;
    const uint32_t b_half = (uint32_t)p.N_t * K_pad * 4;
    uint8_t* sB = smem;
    // the swizzled A stages must start on a 1024-byte boundary of the shared address space
    uint8_t* sA = smem + 2 * b_half;
    sA += (1024u - (smem_u32(sA) & 1023u)) & 1023u;
    const uint32_t stage_bytes = 2 * kCvAHalf;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)S * stage_bytes);
    uint64_t* full = bars;            // [S]
    uint64_t* empty = bars + 8;       // [S]
    uint64_t* d_full = bars + 16;     // [2]
    uint64_t* d_empty = bars + 18;    // [2]
    uint64_t* b_full = bars + 20;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
    constexpr int kMmaWarp = kCvLoadWarps + kCvEpiWarps;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], kCvLoadWarps * 32);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&d_full[s], 1);
            mbar_init(&d_empty[s], kCvEpiWarps * 32);
        }
        mbar_init(b_full, 1);
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
        // ------------------------------------------------------------------ MMA issuer (whole warp loops, one elected lane
        // issues; descriptors are advanced, never rebuilt: tc_common.cuh)
        if (lane == 0) {
            // resident weight image, in pieces of <= 64 KB
            mbar_arrive_expect_tx(b_full, 2 * b_half);
            for (uint32_t off = 0; off < 2 * b_half; off += 65536u) {
                const uint32_t n = min(65536u, 2 * b_half - off);
                bulk_g2s(sB + off, reinterpret_cast<const uint8_t*>(p.Bimg) + off, n, b_full);
            }
            mbar_wait(b_full, 0);
        }
        __syncwarp();
        {
            const uint32_t idesc = make_idesc_tf32(128, p.N_t, /*a MN-major*/ 1, /*b K-major*/ 0);
            const uint32_t lbo_b = (uint32_t)p.N_t * 16;
            const uint64_t a_hi0 = make_smem_desc(smem_u32(sA), kCvLbo, kCvSbo, kCvLayout);
            const uint64_t b_hi0 = make_smem_desc(smem_u32(sB), lbo_b, 128);
            const uint32_t a_lo_off = kCvAHalf >> 4, b_lo_off = b_half >> 4;
            const uint32_t a_step = (2 * kCvSbo) >> 4, b_step = (2 * lbo_b) >> 4, stage_step = (uint32_t)stage_bytes >> 4;
            const int last_nks = (p.K - (NKC - 1) * kKC + 7) / 8;
            int s = 0;
            uint32_t ph = 0;
            uint64_t a_st = a_hi0;
            int it = 0;
            for (long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                mbar_wait(&d_empty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)buf * buf_cols;
                uint64_t db = b_hi0;                     // the weight image is walked once per tile
                uint32_t acc = 0;
                for (int kc = 0; kc < NKC; ++kc) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const int nks = (kc == NKC - 1) ? last_nks : kKC / 8;
                    if (elect_one()) {
                        uint64_t da = a_st, dbk = db;
#pragma unroll
                        for (int ks = 0; ks < kKC / 8; ++ks) {
                            if (ks < nks) {
                                mma_tf32(d_tmem, da, dbk, idesc, ks ? 1u : acc);
                                mma_tf32(d_tmem, da, dbk + b_lo_off, idesc, 1u);
                                mma_tf32(d_tmem, da + a_lo_off, dbk, idesc, 1u);
                            }
                            da += a_step;
                            dbk += b_step;
                        }
                        tc_commit(&empty[s]);
                    }
                    __syncwarp();
                    acc = 1u;
                    db += (uint64_t)(kKC / 8) * b_step;
                    a_st += stage_step;
                    if (++s == S) { s = 0; ph ^= 1u; a_st = a_hi0; }
                }
                if (elect_one()) tc_commit(&d_full[buf]);
                __syncwarp();
            }
        }
    } else if (warp < kCvLoadWarps) {
        // ------------------------------------------------------------------ loaders: chunk = 32 channels x 128 pixels
        // incremental cursors, one 32-bit division per TILE, one 64-bit multiply per chunk
        const int ltid = threadIdx.x;
        const int pg = ltid & 31, cb = ltid >> 5;      // pixel group (4 px), channel base
        const int n_tiles = (int)p.n_tiles, tpb = (int)p.tiles_per_b;
        int n_my = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++n_my;
        const long total = (long)n_my * NKC;
        const long stride_lw = (long)LW * p.npix;
        int i_tile = blockIdx.x, i_kc = 0;
        const float* i_base = nullptr;                 // X + b*sXb + px of the issue cursor's tile
        bool i_pxok = false;
        auto retile = [&]() {
            const int b = i_tile / tpb;
            const long px = (long)(i_tile - b * tpb) * 128 + pg * 4;
            i_pxok = px < p.npix;
            i_base = p.X + (long)b * p.sXb + (i_pxok ? px : 0);
        };
        if (total > 0) retile();
        int p_s = 0;
        uint32_t p_ph = 0;
        // swizzled MN-major destination of this thread inside a stage (channel cl = cb + LW*i -> group (cl>>2) = (LW/4)*i + (cb>>2), row r4 = cb&3)
        const uint32_t r4 = (uint32_t)cb & 3u;
        const uint32_t so = (uint32_t)(pg >> 3) * kCvLbo + (uint32_t)(cb >> 2) * kCvSbo + r4 * 128u + ((((uint32_t)(pg & 7) >> 1) ^ r4) << 5) +
                            ((uint32_t)pg & 1u) * 16u;
        float4 ring[kKpDepth][CPT];
        auto issue = [&](float4 (&v)[CPT]) {
            const int c0 = i_kc * kKC + cb;
            const float* src = i_base + (long)c0 * p.npix;
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i_pxok && c0 + LW * i < p.K) v[i] = __ldg(reinterpret_cast<const float4*>(src + i * stride_lw));
            }
            if (++i_kc == NKC) {
                i_kc = 0;
                i_tile += gridDim.x;
                if (i_tile < n_tiles) retile();
            }
        };
        auto process = [&](const float4 (&v)[CPT]) {
            mbar_wait(&empty[p_s], p_ph ^ 1u);
            uint8_t* st = sA + (size_t)p_s * stage_bytes + so;
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                float4 hi, lo;
                split_tf32(v[i].x, hi.x, lo.x);
                split_tf32(v[i].y, hi.y, lo.y);
                split_tf32(v[i].z, hi.z, lo.z);
                split_tf32(v[i].w, hi.w, lo.w);
                *reinterpret_cast<float4*>(st + i * (LW / 4) * kCvSbo) = hi;
                *reinterpret_cast<float4*>(st + kCvAHalf + i * (LW / 4) * kCvSbo) = lo;
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
    } else {
        // ------------------------------------------------------------------ epilogue: warp e -> lane quarter e%4, column half e/4
        const int e = warp - kCvLoadWarps;
        const int q = e & 3, half = e >> 2;
        const int n_tiles = (int)p.n_tiles, tpb = (int)p.tiles_per_b;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait_relaxed(&d_full[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const int b = tile / tpb;
            const long px = (long)(tile - b * tpb) * 128 + q * 32 + lane;
            const bool pxok = px < p.npix;
            const uint32_t t_base = tmem_base + (uint32_t)buf * buf_cols + ((uint32_t)(q * 32) << 16);
            for (int c0 = half * 16; c0 < p.N_t; c0 += 32) {
                if (c0 >= p.N) break;
                uint32_t r[16];
                tmem_ld_32x32b_x16(t_base + (uint32_t)c0, r);
                float bv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) bv[j] = (p.bias && c0 + j < p.N) ? __ldg(p.bias + c0 + j) : 0.f;
                float* yp = p.Y + (long)b * p.sYb + (long)c0 * p.npix + (pxok ? px : 0);
                tmem_ld_wait();
                if (pxok) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (c0 + j < p.N) *yp = __uint_as_float(r[j]) + bv[j];
                        yp += p.npix;
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
