// tcgen05 weight gradient of the 1x1 channel mix:  dW[m, n] += sum_b sum_p G[b, m, p] * X[b, n, p]
//   (m = output channel, n = input channel, p = pixel; both operands have the pixel index contiguous).
//
// UMMA view: D[128 (m, zero padded), N] += A[128, 8] * B[8, N] per k-step with K = pixels, both operands K-major.
// Split-K: the (sample, 32-pixel chunk) sequence is divided evenly over persistent CTAs; each CTA accumulates
// its whole range in ONE TMEM accumulator (fp32) and flushes it once with atomicAdd -- 148 x M x N atomics in
// total instead of one per split-K tile.
//   * 16 loader warps: float4 along pixels, tf32 hi/lo split in registers, K-major interleave layout, 4-deep
//     register prefetch ring, rows beyond the real channel count are zeroed once and never touched again
//     (ncu on the 8-warp form: 14 % of the warp slots occupied, 3.3 long-scoreboard stalls per issue, 0.30 of HBM)
//   * one lane issues 3 MMAs per 8-pixel k-step; stages are recycled through tcgen05.commit -> mbarrier
#pragma once
#include "backend.h"
#include "tc_common.cuh"
#include "tc_kpipe.cuh"

namespace uno {
namespace tc {

struct WgradParams {
    const float* G; long sGb;     // [batch][M][npix]
    const float* X; long sXb;     // [batch][N][npix]
    float* dW; long ldw;          // [M][N] (+=, pre-zeroed by the caller)
    long npix;
    int M, N, N_t, batch, stages;
    int chunks_per_b;             // ceil(npix / 32)
    long total_chunks;
    int tmem_cols;
    int debug;                    // only read by the <true> instantiation (UNO_B200_WGRAD_DEBUG, results are garbage): 1 no MMA,
                                  // 2 no operand stores, 8 no global loads, 16 no proxy fence, 32 no final flush
};

__host__ __device__ constexpr int wg_threads(int LW) { return (LW + 1) * 32; }
constexpr uint32_t kWgStage = 4 * kKpAHalf;   // A_hi, A_lo, B_hi, B_lo (each 128 rows x 32 k, K-major interleave)

__host__ __device__ inline size_t wgrad_smem_bytes(int stages) { return (size_t)stages * kWgStage + 32 * 8 + 16; }

// DBG = timing-probe instantiation (tools/wgrad_probe.py); the default <false> carries none of the probe branches
// LW = loader warps (8 or 16): a thread owns 256 / (4 * LW) rows of every chunk
template <int LW, bool DBG = false>
__global__ void __launch_bounds__(wg_threads(LW), 1) wgrad_tc_kernel(const WgradParams p) {
    constexpr int kWgLoadWarps = LW, kWgThreads = wg_threads(LW);
    const int dbg = DBG ? p.debug : 0;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * kWgStage);
    uint64_t* full = bars;          // [S]
    uint64_t* empty = bars + 8;     // [S]
    uint64_t* done = bars + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    constexpr int kMmaWarp = kWgLoadWarps;

    // this CTA's contiguous range of chunks
    const long per = (p.total_chunks + gridDim.x - 1) / gridDim.x;
    const long c_begin = (long)blockIdx.x * per;
    const long c_end = min(p.total_chunks, c_begin + per);
    const long total = c_end > c_begin ? c_end - c_begin : 0;

    // zero all stages once: padded rows (>= M resp. >= N) stay zero for the whole kernel
    for (uint32_t i = threadIdx.x; i < (uint32_t)S * kWgStage / 16; i += kWgThreads)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], kWgLoadWarps * 32);
            mbar_init(&empty[s], 1);
        }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kMmaWarp) {
        // whole warp loops, one elected lane issues; descriptors are advanced, never rebuilt (tc_common.cuh)
        if (total > 0) {
            const uint32_t idesc = make_idesc_tf32(128, p.N_t, 0, 0);
            const uint64_t a_hi0 = make_smem_desc(smem_u32(smem), kLboA, 128);
            constexpr uint32_t half = kKpAHalf >> 4, step = (2 * kLboA) >> 4, stage_step = kWgStage >> 4;
            int s = 0;
            uint32_t ph = 0;
            uint64_t a_st = a_hi0;
            uint32_t acc = 0;
            for (long g = 0; g < total; ++g) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    uint64_t da = a_st;                  // stage layout: [A hi | A lo | B hi | B lo], each kKpAHalf bytes
#pragma unroll
                    for (int ks = 0; ks < kKC / 8; ++ks) {
                        if (!(DBG && (dbg & 1))) {
                            mma_tf32(tmem_base, da, da + 2 * half, idesc, ks ? 1u : acc);
                            mma_tf32(tmem_base, da, da + 3 * half, idesc, 1u);
                            mma_tf32(tmem_base, da + half, da + 2 * half, idesc, 1u);
                        }
                        da += step;
                    }
                    tc_commit(&empty[s]);
                }
                __syncwarp();
                acc = 1u;
                a_st += stage_step;
                if (++s == S) { s = 0; ph ^= 1u; a_st = a_hi0; }
            }
            if (elect_one()) tc_commit(done);
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ loaders
        const int ltid = threadIdx.x;
        const int kq = ltid & 7, rbase = ltid >> 3;          // 4 pixels at kq*4 ; rows rbase + 32*i of the combined [G rows | X rows]
        const int rows = p.M + p.N;
        const int cpb = p.chunks_per_b;
        // issue cursor
        long i_chunk = c_begin;
        int i_b = (int)(c_begin / cpb);
        int i_c = (int)(c_begin - (long)i_b * cpb);
        int p_s = 0;
        uint32_t p_ph = 0;
        constexpr int RPT = 64 / LW;          // rows per thread: (M + N) <= 256
        constexpr int RSTEP = 4 * LW;         // row distance between a thread's rows
        float4 ring[kKpDepth][RPT];
        auto issue = [&](float4 (&v)[RPT]) {
            const long px = (long)i_c * kKC + kq * 4;
            const bool pxok = px < p.npix;       // npix % 4 == 0: a float4 is either fully inside or fully outside
            const float* gsrc = p.G + (long)i_b * p.sGb + px;
            const float* xsrc = p.X + (long)i_b * p.sXb + px;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int r = rbase + RSTEP * i;
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pxok && r < rows && !(DBG && (dbg & 8))) {
                    const float* q = (r < p.M) ? gsrc + (long)r * p.npix : xsrc + (long)(r - p.M) * p.npix;
                    v[i] = ldg_f4<true>(reinterpret_cast<const float4*>(q));
                }
            }
            ++i_chunk;
            if (++i_c == cpb) { i_c = 0; ++i_b; }
        };
        auto process = [&](const float4 (&v)[RPT]) {
            mbar_wait(&empty[p_s], p_ph ^ 1u);
            uint8_t* st = smem + (size_t)p_s * kWgStage;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int r = rbase + RSTEP * i;
                if (r < rows && !(DBG && (dbg & 2))) {
                    float4 hi, lo;
                    split_tf32(v[i].x, hi.x, lo.x);
                    split_tf32(v[i].y, hi.y, lo.y);
                    split_tf32(v[i].z, hi.z, lo.z);
                    split_tf32(v[i].w, hi.w, lo.w);
                    const bool isA = r < p.M;
                    const int rr = isA ? r : r - p.M;
                    uint8_t* d = st + (isA ? 0u : 2 * kKpAHalf) + (uint32_t)kq * kLboA + (uint32_t)rr * 16;
                    *reinterpret_cast<float4*>(d) = hi;
                    *reinterpret_cast<float4*>(d + kKpAHalf) = lo;
                }
            }
            if (!(DBG && (dbg & 16))) fence_proxy_async();
            mbar_arrive(&full[p_s]);
            if (++p_s == S) { p_s = 0; p_ph ^= 1u; }
        };
        {
            // Loads are issued for two consecutive 32-pixel chunks at once (256 contiguous bytes of every channel row per request
            // burst, L2::256B hint): the rows of a chunk are whole channel planes apart, so every 128-byte piece would otherwise
            // be a DRAM access of its own.  Measured on B200: 1.94 -> 1.66 ms per Darcy step, 1.30 -> 1.14 ms per NS-3D step.
            static_assert(kKpDepth == 4, "ring of four chunks");
            if (0 < total) issue(ring[0]);
            if (1 < total) issue(ring[1]);
            for (long g = 0; g < total; g += 4) {
                if (g + 2 < total) issue(ring[2]);
                if (g + 3 < total) issue(ring[3]);
                if (g + 0 < total) process(ring[0]);
                if (g + 1 < total) process(ring[1]);
                if (g + 4 < total) issue(ring[0]);
                if (g + 5 < total) issue(ring[1]);
                if (g + 2 < total) process(ring[2]);
                if (g + 3 < total) process(ring[3]);
            }
        }
        // ------------------------------------------------------------------ flush (loader warps 0-3 own the TMEM lane quarters)
        if (warp < 4 && total > 0) {
            mbar_wait_relaxed(done, 0);
            tc_fence_after();
            const int m = warp * 32 + lane;
            for (int c0 = 0; c0 < p.N_t; c0 += 16) {
                uint32_t r[16];
                tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
                tmem_ld_wait();
                if (m < p.M && !(DBG && (dbg & 32))) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < p.N) atomicAdd(p.dW + (long)m * p.ldw + c0 + j, __uint_as_float(r[j]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace tc
}  // namespace uno
