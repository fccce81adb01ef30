// tcgen05 "row GEMM" with a small contraction dimension:  C[R, N] (op)= A[R, K] * B[K, N],  K <= 64.
//
// This is the synthesis stage of the truncated inverse DFT along the last axis (kept modes -> real
// samples, K = 2*modes) with the block epilogue fused:  C += acc, optional exact-erf GELU to a second
// tensor.  With K this small the kernel is a pure HBM streamer, so it is organised around the stores:
//
//   * persistent CTAs (one per SM), static round-robin over 128-row tiles
//   * B (the constant twiddle matrix, pre-split into tf32 hi/lo images in UMMA canonical layout on the
//     host) is loaded ONCE per CTA with a bulk async copy and stays resident in shared memory
//   * warps 4-5 stream the A tile (global -> registers -> hi/lo split -> st.shared in the UMMA
//     K-major "interleave" layout), double buffered
//   * warp 6 (one lane) issues 3 x K/8 tcgen05.mma.kind::tf32 per tile into one of two TMEM accumulators
//   * warps 0-3 drain the other accumulator: tcgen05.ld 16x256b -> fused epilogue -> global
//     (each quad of lanes covers one full 32-byte sector of a row)
//   so the tensor core, the loaders and the epilogue of consecutive tiles overlap.
#pragma once
#include "backend.h"
#include "tc_common.cuh"

namespace uno {
namespace tc {

struct RowGemmParams {
    const float* A; long lda; long R;
    const float* Bimg;            // per n-tile: [hi image | lo image]; image = (K_pad/4) x N_t x 16 bytes
    float* C; float* C2; long ldc;
    int N, K, K_pad;
    int n_tiles, N_t;
    long m_tiles;
    int epi;
    int tmem_cols;                // power of two >= 2*N_t
    int a_vec_ok;                 // A rows are 16-byte aligned and K % 4 == 0
};

constexpr int kRowGemmEpiWarps = 8;                      // two per TMEM lane quarter, splitting the columns
constexpr int kRowGemmLoadWarps = 2;
constexpr int kRowGemmThreads = (kRowGemmEpiWarps + kRowGemmLoadWarps + 1) * 32;
constexpr uint32_t kLboA = 128 * 16 + 16;   // +16 B: the loaders' 16-byte stores of one warp fall in distinct bank groups

__host__ __device__ inline size_t rowgemm_smem_bytes(int K_pad, int N_t) {
    const size_t b = (size_t)2 * K_pad * N_t * 4;
    const size_t a = (size_t)4 * (K_pad / 4) * kLboA;
    return b + a + 16 * 8 + 16;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Epilogue of one 128-row tile for one warp: TMEM lane quarter `q`, 16-column chunks c0 = 16*(2*i + half).
// All loads of a 32-column group are issued before any store so that they are in flight together.
template <int EPI, bool VEC2>
__device__ __forceinline__ void rowgemm_epilogue_tile(const RowGemmParams& p, uint32_t t_base, long row0, int n_base, int half, int lane) {
    constexpr bool kAccum = (EPI != EPI_STORE);
    float* __restrict__ C = p.C;
    float* __restrict__ C2 = p.C2;
    // pairs of chunks: chunk indices (2*i + half) for i = 0.. ; process two chunks (i, i+1) per iteration
    for (int ci = half; ci * 16 < p.N_t; ci += 4) {
        const int c0a = ci * 16, c0b = (ci + 2) * 16;
        const bool has_b = c0b < p.N_t && n_base + c0b < p.N;
        if (n_base + c0a >= p.N) break;
        uint32_t r[4][8];
        tmem_ld_16x256b_x2(t_base + (uint32_t)c0a, r[0]);
        tmem_ld_16x256b_x2(t_base + (16u << 16) + (uint32_t)c0a, r[1]);
        if (has_b) {
            tmem_ld_16x256b_x2(t_base + (uint32_t)c0b, r[2]);
            tmem_ld_16x256b_x2(t_base + (16u << 16) + (uint32_t)c0b, r[3]);
        }
        // addresses: element pair e = (chunk j, half hh, repeat rep, row-pair rr)
        long off[16];
        bool ok0[16], ok1[16];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int rep = 0; rep < 2; ++rep)
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int e = ((j * 2 + hh) * 2 + rep) * 2 + rr;
                        const long grow = row0 + hh * 16 + rr * 8 + (lane >> 2);
                        const int col = n_base + (j ? c0b : c0a) + rep * 8 + 2 * (lane & 3);
                        const bool live = (j == 0 || has_b) && grow < p.R;
                        ok0[e] = live && col < p.N;
                        ok1[e] = live && col + 1 < p.N;
                        off[e] = grow * p.ldc + col;
                    }
        float2 cz[16];
        if (kAccum) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                cz[e] = make_float2(0.f, 0.f);
                if (VEC2) {
                    if (ok1[e]) cz[e] = *reinterpret_cast<const float2*>(C + off[e]);
                    else if (ok0[e]) cz[e].x = C[off[e]];
                } else {
                    if (ok0[e]) cz[e].x = C[off[e]];
                    if (ok1[e]) cz[e].y = C[off[e] + 1];
                }
            }
        }
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int j = e >> 3, hh = (e >> 2) & 1, rep = (e >> 1) & 1, rr = e & 1;
            float2 acc = make_float2(__uint_as_float(r[j * 2 + hh][rep * 4 + rr * 2 + 0]), __uint_as_float(r[j * 2 + hh][rep * 4 + rr * 2 + 1]));
            if (kAccum) { acc.x += cz[e].x; acc.y += cz[e].y; }
            float2 act = acc;
            if (EPI == EPI_ACCUM_GELU || EPI == EPI_ACCUM_GELU_INPLACE) { act.x = gelu_erf(acc.x); act.y = gelu_erf(acc.y); }
            const float2 out1 = (EPI == EPI_ACCUM_GELU_INPLACE) ? act : acc;
            if (VEC2 && ok1[e]) {
                *reinterpret_cast<float2*>(C + off[e]) = out1;
                if (EPI == EPI_ACCUM_GELU) *reinterpret_cast<float2*>(C2 + off[e]) = act;
            } else {
                if (ok0[e]) { C[off[e]] = out1.x; if (EPI == EPI_ACCUM_GELU) C2[off[e]] = act.x; }
                if (ok1[e]) { C[off[e] + 1] = out1.y; if (EPI == EPI_ACCUM_GELU) C2[off[e] + 1] = act.y; }
            }
        }
    }
}

template <int EPI>
__global__ void __launch_bounds__(kRowGemmThreads, 1) rowgemm_smallk_kernel(const RowGemmParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nt = blockIdx.y;
    const int CH = p.K_pad / 4;
    const uint32_t b_half = (uint32_t)p.K_pad * p.N_t * 4;
    const uint32_t a_bytes = (uint32_t)CH * kLboA;
    uint8_t* sB = smem;
    uint8_t* sA = smem + 2 * b_half;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + 4 * a_bytes);
    uint64_t* a_full = bars;        // [2] loaders -> mma
    uint64_t* a_empty = bars + 2;   // [2] mma -> loaders
    uint64_t* d_full = bars + 4;    // [2] mma -> epilogue
    uint64_t* d_empty = bars + 6;   // [2] epilogue -> mma
    uint64_t* b_full = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
    constexpr int kMmaWarp = kRowGemmEpiWarps + kRowGemmLoadWarps;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a_full[s], kRowGemmLoadWarps * 32);
            mbar_init(&a_empty[s], 1);
            mbar_init(&d_full[s], 1);
            mbar_init(&d_empty[s], kRowGemmEpiWarps * 32);
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

    if (warp == kMmaWarp) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            mbar_arrive_expect_tx(b_full, 2 * b_half);
            bulk_g2s(sB, p.Bimg + (size_t)nt * (2 * b_half / 4), 2 * b_half, b_full);
            mbar_wait(b_full, 0);
            const uint32_t idesc = make_idesc_tf32(128, p.N_t, 0, 0);
            const uint32_t lbo_b = (uint32_t)p.N_t * 16;
            const uint32_t sB_addr = smem_u32(sB);
            int it = 0;
            for (long tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
                const int s = it & 1;
                const uint32_t ph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(&a_full[s], ph);
                mbar_wait(&d_empty[s], ph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)s * buf_cols;
                const uint32_t a_hi = smem_u32(sA + (size_t)(s * 2 + 0) * a_bytes);
                const uint32_t a_lo = smem_u32(sA + (size_t)(s * 2 + 1) * a_bytes);
                for (int ks = 0; ks < p.K_pad / 8; ++ks) {
                    const uint64_t da_hi = make_smem_desc(a_hi + ks * 2 * kLboA, kLboA, 128);
                    const uint64_t da_lo = make_smem_desc(a_lo + ks * 2 * kLboA, kLboA, 128);
                    const uint64_t db_hi = make_smem_desc(sB_addr + ks * 2 * lbo_b, lbo_b, 128);
                    const uint64_t db_lo = make_smem_desc(sB_addr + b_half + ks * 2 * lbo_b, lbo_b, 128);
                    mma_tf32(d_tmem, da_hi, db_hi, idesc, ks > 0 ? 1u : 0u);
                    mma_tf32(d_tmem, da_hi, db_lo, idesc, 1u);
                    mma_tf32(d_tmem, da_lo, db_hi, idesc, 1u);
                }
                tc_commit(&a_empty[s]);
                tc_commit(&d_full[s]);
            }
        }
    } else if (warp >= kRowGemmEpiWarps) {
        // ------------------------------------------------------------------ A loaders
        constexpr int NL = kRowGemmLoadWarps * 32;
        constexpr int BATCH = 5;
        const int ltid = threadIdx.x - kRowGemmEpiWarps * 32;
        const int total = 128 * CH;
        int it = 0;
        for (long tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
            const int s = it & 1;
            const uint32_t ph = (uint32_t)(it >> 1) & 1u;
            mbar_wait(&a_empty[s], ph ^ 1u);
            uint8_t* dst_hi = sA + (size_t)(s * 2 + 0) * a_bytes;
            uint8_t* dst_lo = sA + (size_t)(s * 2 + 1) * a_bytes;
            const long row0 = tile * 128;
            for (int cb = ltid; cb < total; cb += NL * BATCH) {
                float4 v[BATCH];
                uint32_t o[BATCH];
#pragma unroll
                for (int u = 0; u < BATCH; ++u) {
                    const int c = cb + u * NL;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    o[u] = 0xFFFFFFFFu;
                    if (c < total) {
                        const int row = c / CH, kc = c - row * CH;
                        o[u] = (uint32_t)kc * kLboA + (uint32_t)row * 16;
                        const long grow = row0 + row;
                        if (grow < p.R) {
                            const float* src = p.A + grow * p.lda + kc * 4;
                            if (p.a_vec_ok && kc * 4 + 4 <= p.K) {
                                v[u] = __ldg(reinterpret_cast<const float4*>(src));
                            } else {
                                if (kc * 4 + 0 < p.K) v[u].x = __ldg(src + 0);
                                if (kc * 4 + 1 < p.K) v[u].y = __ldg(src + 1);
                                if (kc * 4 + 2 < p.K) v[u].z = __ldg(src + 2);
                                if (kc * 4 + 3 < p.K) v[u].w = __ldg(src + 3);
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < BATCH; ++u) {
                    if (o[u] == 0xFFFFFFFFu) continue;
                    float4 hi, lo;
                    split_tf32(v[u].x, hi.x, lo.x);
                    split_tf32(v[u].y, hi.y, lo.y);
                    split_tf32(v[u].z, hi.z, lo.z);
                    split_tf32(v[u].w, hi.w, lo.w);
                    *reinterpret_cast<float4*>(dst_hi + o[u]) = hi;
                    *reinterpret_cast<float4*>(dst_lo + o[u]) = lo;
                }
            }
            fence_proxy_async();
            mbar_arrive(&a_full[s]);
        }
    } else {
        // ------------------------------------------------------------------ epilogue: warp e -> TMEM lane quarter e%4, column half e/4
        const bool vec2 = (p.ldc % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 7) == 0) &&
                          (EPI != EPI_ACCUM_GELU || (reinterpret_cast<uintptr_t>(p.C2) & 7) == 0);
        const int q = warp & 3, half = warp >> 2;
        const int n_base = nt * p.N_t;
        int it = 0;
        for (long tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
            const int s = it & 1;
            const uint32_t ph = (uint32_t)(it >> 1) & 1u;
            mbar_wait(&d_full[s], ph);
            tc_fence_after();
            const long row0 = tile * 128 + q * 32;
            const uint32_t t_base = tmem_base + (uint32_t)s * buf_cols + ((uint32_t)(q * 32) << 16);
            if (vec2) rowgemm_epilogue_tile<EPI, true>(p, t_base, row0, n_base, half, lane);
            else rowgemm_epilogue_tile<EPI, false>(p, t_base, row0, n_base, half, lane);
            tc_fence_before();
            mbar_arrive(&d_empty[s]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace tc
}  // namespace uno
