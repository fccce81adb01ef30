// tcgen05 per-mode complex channel contraction (SpectralConv*.compl_mul*, integral_operators.py:45, :179, :383):
//     C[m, n, q] = sum_k opA(A[m, k, q]) * opB(B[k, n, q])          one small complex GEMM per kept mode q
// (forward: m = sample, k = in channel, n = out channel; dX and dW are the same call with other strides / conj flags).
//
// Per mode the complex product is ONE real GEMM with the contraction index doubled, k' = (k, re|im):
//     D[(n, re|im), m] = sum_k' Wimg[(n, re|im), k'] * Ximg[m, k']
//     Ximg[m, (k,re)] = Re A      Ximg[m, (k,im)] = sa Im A                      (sa, sb = -1 when conjugated)
//     Wimg[(n,re), (k,re)] = Re B   Wimg[(n,re), (k,im)] = -sb Im B   Wimg[(n,im), (k,re)] = sb Im B   Wimg[(n,im), (k,im)] = Re B
// The operand with the many rows -- the (n, re|im) rows of the weights, 128 - 384 of them -- sits on the UMMA M side
// (128 rows per tile, no padding waste) and the sample index on the N side (16 - 64 columns).  Both images are K-major
// "interleave" layouts written by the loader warps from 8-byte global loads (tf32 hi/lo split on the way, as in
// tc_kpipe.cuh); MMA issue, stage ring and TMEM double buffering are tc_kpipe.cuh's.
//
// A work item = (row tile, column tile, corner, mode); items are numbered with the mode index fastest, so the CTAs
// running at the same time touch neighbouring modes and share the 32-byte sectors their 8-byte loads pull into L2.
// The epilogue pairs the (re, im) rows held by neighbouring lanes with one shuffle per output and stores float2.
//
// STATUS: opt-in (UNO_B200_CMM_TC=1) until it has been run against the SIMT kernels on a B200; the default contraction
// kernels are cmm2_kernel / cmm_kernel (backend_cuda.cu).
#pragma once
#include "backend.h"
#include "tc_common.cuh"
#include "tc_kpipe.cuh"

namespace uno {
namespace tc {

struct CmmTcParams {
    CmmArgs a;
    int N_t;          // columns of a tile (samples m), multiple of 16, <= 64 (two sample-side tasks per loader thread)
    int ns_tiles;     // column tiles
    int ms_tiles;     // 128-row tiles over the 2*N rows (n, re|im)
    int n_chunks;     // ceil(K / 16): a chunk is 16 complex k = 32 real k'
    int stages, tmem_cols;
    long items;       // ms_tiles * ns_tiles * ncorner * q_outer * q_inner
};

constexpr int kCmKC = kKC / 2;        // complex k per chunk
constexpr int kCmDepth = 4;           // chunks of global loads in flight per loader thread
constexpr int kCmXT = 2;              // sample-side tasks per loader thread: 8 * N_t <= 256 * kCmXT

struct CmmItem {
    int ms, ns, corner, qo, qi;
};
__device__ __forceinline__ CmmItem cmm_item(const CmmTcParams& p, long w) {
    CmmItem it;
    it.qi = (int)(w % p.a.q_inner); w /= p.a.q_inner;
    it.qo = (int)(w % p.a.q_outer); w /= p.a.q_outer;
    it.corner = (int)(w % p.a.ncorner); w /= p.a.ncorner;
    it.ns = (int)(w % p.ns_tiles);
    it.ms = (int)(w / p.ns_tiles);
    return it;
}

__global__ void __launch_bounds__(kKpThreads, 1) cmm_tc_kernel(const CmmTcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    const uint32_t b_half = (uint32_t)p.N_t * kKC * 4;
    const uint32_t stage_bytes = 2 * kKpAHalf + 2 * b_half;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;            // [S] loaders -> mma
    uint64_t* empty = bars + 8;       // [S] mma -> loaders
    uint64_t* d_full = bars + 16;     // [2]
    uint64_t* d_empty = bars + 18;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    constexpr int kMmaWarp = kKpLoadWarps + kKpEpiWarps;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], kKpLoadWarps * 32);
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
        const uint32_t a_lo_off = kKpAHalf >> 4, b_lo_off = b_half >> 4;
        const uint32_t a_step = (2 * kLboA) >> 4, b_step = (2 * lbo_b) >> 4, stage_step = stage_bytes >> 4;
        const int last_nks = (2 * p.a.K - (NKC - 1) * kKC + 7) / 8;
        int s = 0;
        uint32_t ph = 0;
        uint64_t a_st = a_hi0, b_st = b_hi0;
        int it = 0;
        for (long w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
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
        // ------------------------------------------------------------------ loaders
        // A thread owns the same tasks in every chunk.  Weight side: (n_l, k pair kp) -> the two image rows (n_l, re) and
        // (n_l, im), one 16-byte row each.  Sample side: (m_l, kp) -> one 16-byte row.  A task = two 8-byte global loads.
        const int ltid = threadIdx.x;
        const float sa = p.a.conjA ? -1.f : 1.f, sb = p.a.conjB ? -1.f : 1.f;
        long n_my = 0;
        for (long w = blockIdx.x; w < p.items; w += gridDim.x) ++n_my;
        const long total = n_my * NKC;
        // fixed per-thread task geometry
        const int wn = ltid & 63;                       // weight row pair within the tile
        const int wkp0 = ltid >> 6;                     // k pairs wkp0, wkp0 + 4
        const uint32_t w_so = (uint32_t)wkp0 * kLboA + (uint32_t)(2 * wn) * 16;
        int xm[kCmXT], xkp[kCmXT];
        uint32_t x_so[kCmXT];
        const int n_xtasks = 8 * p.N_t;
#pragma unroll
        for (int j = 0; j < kCmXT; ++j) {
            const int idx = ltid + 256 * j;
            xkp[j] = idx / p.N_t;
            xm[j] = idx - xkp[j] * p.N_t;
            x_so[j] = 2 * kKpAHalf + (uint32_t)xkp[j] * (uint32_t)(p.N_t * 16) + (uint32_t)xm[j] * 16;
            if (idx >= n_xtasks) xkp[j] = -1;
        }
        // issue-side cursor
        long i_w = blockIdx.x;
        int i_kc = 0;
        const float2* gB = nullptr;       // &B[k = 0, n of this thread, q]   (null: row beyond N)
        const float2* gA[kCmXT];              // &A[m of task j, k = 0, q]        (null: row beyond M or no task)
        auto seek = [&]() {
            gB = nullptr;
#pragma unroll
            for (int j = 0; j < kCmXT; ++j) gA[j] = nullptr;
            if (i_w >= p.items) return;
            const CmmItem it = cmm_item(p, i_w);
            const int n = it.ms * 64 + wn;
            if (n < p.a.N)
                gB = reinterpret_cast<const float2*>(p.a.B[it.corner]) + (long)it.qo * p.a.b_sqo + it.qi + (long)n * p.a.b_sn;
#pragma unroll
            for (int j = 0; j < kCmXT; ++j) {
                const int m = it.ns * p.N_t + xm[j];
                if (xkp[j] >= 0 && m < p.a.M)
                    gA[j] = reinterpret_cast<const float2*>(p.a.A[it.corner]) + (long)it.qo * p.a.a_sqo + it.qi + (long)m * p.a.a_sm;
            }
        };
        seek();
        int p_s = 0;
        uint32_t p_ph = 0;
        struct Slot { float4 w[2]; float4 x[kCmXT]; };
        Slot ring[kCmDepth];
        auto issue = [&](Slot& v) {
            const int k0 = i_kc * kCmKC;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int k = k0 + 2 * (wkp0 + 4 * j);
                float2 e0 = make_float2(0.f, 0.f), e1 = make_float2(0.f, 0.f);
                if (gB) {
                    if (k < p.a.K) e0 = __ldg(gB + (long)k * p.a.b_sk);
                    if (k + 1 < p.a.K) e1 = __ldg(gB + (long)(k + 1) * p.a.b_sk);
                }
                v.w[j] = make_float4(e0.x, e0.y, e1.x, e1.y);
            }
#pragma unroll
            for (int j = 0; j < kCmXT; ++j) {
                float2 e0 = make_float2(0.f, 0.f), e1 = make_float2(0.f, 0.f);
                if (gA[j]) {
                    const int k = k0 + 2 * xkp[j];
                    if (k < p.a.K) e0 = __ldg(gA[j] + (long)k * p.a.a_sk);
                    if (k + 1 < p.a.K) e1 = __ldg(gA[j] + (long)(k + 1) * p.a.a_sk);
                }
                v.x[j] = make_float4(e0.x, e0.y, e1.x, e1.y);
            }
            if (++i_kc == NKC) { i_kc = 0; i_w += gridDim.x; seek(); }
        };
        auto put = [&](uint8_t* dst, uint32_t lo_off, float4 v) {
            float4 hi, lo;
            split_tf32(v.x, hi.x, lo.x);
            split_tf32(v.y, hi.y, lo.y);
            split_tf32(v.z, hi.z, lo.z);
            split_tf32(v.w, hi.w, lo.w);
            *reinterpret_cast<float4*>(dst) = hi;
            *reinterpret_cast<float4*>(dst + lo_off) = lo;
        };
        auto process = [&](const Slot& v) {
            mbar_wait(&empty[p_s], p_ph ^ 1u);
            uint8_t* st = smem + (size_t)p_s * stage_bytes;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float4 e = v.w[j];                                   // (Br, Bi, Br', Bi') of k, k+1
                uint8_t* d = st + w_so + (uint32_t)(4 * j) * kLboA;
                put(d, kKpAHalf, make_float4(e.x, -sb * e.y, e.z, -sb * e.w));        // row (n, re)
                put(d + 16, kKpAHalf, make_float4(sb * e.y, e.x, sb * e.w, e.z));     // row (n, im)
            }
#pragma unroll
            for (int j = 0; j < kCmXT; ++j) {
                if (xkp[j] < 0) continue;
                const float4 e = v.x[j];
                put(st + x_so[j], b_half, make_float4(e.x, sa * e.y, e.z, sa * e.w));
            }
            fence_proxy_async();
            mbar_arrive(&full[p_s]);
            if (++p_s == S) { p_s = 0; p_ph ^= 1u; }
        };
#pragma unroll
        for (int d = 0; d < kCmDepth - 1; ++d)
            if (d < total) issue(ring[d]);
        for (long g = 0; g < total; g += kCmDepth) {
#pragma unroll
            for (int d = 0; d < kCmDepth; ++d) {
                if (g + d + kCmDepth - 1 < total) issue(ring[(d + kCmDepth - 1) % kCmDepth]);
                if (g + d < total) process(ring[d]);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: one warp per TMEM lane quarter
        // lane pair (2t, 2t+1) holds the (re, im) rows of one n; even lanes store the even columns (samples), odd lanes the
        // odd ones, after exchanging the missing half with the neighbour.
        const int q = warp - kKpLoadWarps;
        const int odd = lane & 1;
        int it = 0;
        for (long w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
            const int buf = it & 1;
            const CmmItem im = cmm_item(p, w);
            mbar_wait_relaxed(&d_full[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_base = tmem_base + (uint32_t)buf * buf_cols + ((uint32_t)(q * 32) << 16);
            const int n = im.ms * 64 + ((q * 32 + lane) >> 1);
            const bool nok = n < p.a.N;
            const int m0 = im.ns * p.N_t;
            float2* crow = reinterpret_cast<float2*>(p.a.C[im.corner]) + (long)im.qo * p.a.c_sqo + im.qi + (long)(nok ? n : 0) * p.a.c_sn;
            const int mcols = min(p.N_t, p.a.M - m0);          // valid columns of this tile
            for (int c0 = 0; c0 < mcols; c0 += 16) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(t_base + (uint32_t)c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float mine_keep = __uint_as_float(odd ? v[2 * u + 1] : v[2 * u]);
                    const float send = __uint_as_float(odd ? v[2 * u] : v[2 * u + 1]);
                    const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
                    const int c = c0 + 2 * u + odd;
                    if (nok && c < mcols)
                        crow[(long)(m0 + c) * p.a.c_sm] = odd ? make_float2(recv, mine_keep) : make_float2(mine_keep, recv);
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
