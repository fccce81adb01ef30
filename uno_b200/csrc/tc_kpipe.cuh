// tcgen05 GEMM with a long contraction dimension streamed from HBM:  C[R, N] = A[R, K] * B[K, N],
// A row-major (K contiguous), N <= 256, K arbitrary.
//
// This is the analysis stage of the truncated DFT along the last axis (real samples -> kept modes:
// K = grid width, N = 2*modes): x is read exactly once, 128 rows x 32 columns at a time.
//
//   * persistent CTAs (one per SM), static round-robin over 128-row tiles
//   * a ring of S shared-memory stages; per stage the A chunk (tf32 hi and lo images, UMMA K-major
//     "interleave" layout) written by 8 loader warps and the matching chunk of the constant B image
//     (pre-split hi/lo, chunk-major in global memory, L2 resident) fetched by one bulk async copy
//   * loader warps keep up to FOUR chunks of global loads in flight per thread (a register ring), issued two
//     consecutive chunks at a time (see "DRAM locality" below); rows that are only 4-byte aligned (odd grid
//     widths such as 481 rule out TMA tensor maps) take the row-class path, which still loads 16 bytes at a time
//   * one lane issues 3 MMAs (hi*hi, hi*lo, lo*hi) per 8-wide k-step into one of two TMEM accumulators
//   * 4 epilogue warps drain the other accumulator (tcgen05.ld 16x256b -> float2 stores)
#pragma once
#include "backend.h"
#include "tc_common.cuh"
#include "tc_rowgemm.cuh"

namespace uno {
namespace tc {

struct KPipeParams {
    const float* A; long lda; long R;
    const float* Bimg;     // [n_chunks][hi | lo], each half (KC/4) x N_t x 16 bytes
    float* C; long ldc;
    int N, K, N_t, n_chunks, stages;
    long m_tiles;
    int tmem_cols;
    int a_vec_ok;          // rows 16-byte aligned (lda % 4 == 0, base aligned)
    int rclass;            // opt-in (UNO_B200_KPIPE_ALIGN=1) for rows that are NOT 16-byte aligned: see "row classes" below
    int k_valid;           // row-class mode: the real contraction length (K is then k_valid + 3, the longest shifted row)
    int debug;             // timing probes (UNO_B200_KPIPE_DEBUG, results are garbage): 1 no MMA, 2 no operand stores, 4 no B copy, 8 no global loads (not in the row-class path), 32 plain arrival instead of tcgen05.commit, 128 one arrival per loader warp,
                           // 16 no proxy fence
};

// Row classes (rclass = 1).  With an odd row pitch (the 481-wide Darcy grid, the 83-long NS-3D time axis) only every fourth
// row starts on a 16-byte boundary and the loaders fall to 4-byte loads, whose instruction count is what bounds this kernel
// (tools/kpipe_probe.py: 0.69 us per chunk against 0.39 us with 16-byte loads).  In this mode a tile holds 128 rows of ONE
// residue class mod 4 out of a block of 512 (row = 512*(tile/4) + 4*i + tile%4) and a CTA only ever sees one class (the grid
// is a multiple of 4).  All rows of a class start the same number of floats sh = (class*lda) % 4 past a 16-byte boundary, so
// the loaders read each row from that boundary with 16-byte loads -- the contraction index is then shifted by sh -- and the
// CTA streams a B image whose rows are shifted by the same sh (rows k' < sh are zero; the loaders also zero those elements,
// which belong to the previous row).  The same idea as the column-shifted twiddle image of tc_rowgemm.cuh's parity mode.
constexpr int kKC = 32;                       // K elements per stage
constexpr int kKpLoadWarps = 8;
constexpr int kKpDepth = 4;                   // chunks of global loads kept in flight per loader thread
constexpr int kKpEpiWarps = 4;
#ifndef KPIPE_LW16_DEPTH
#define KPIPE_LW16_DEPTH 4
#endif
constexpr int kKpThreads = (kKpLoadWarps + kKpEpiWarps + 1) * 32;
constexpr uint32_t kKpAHalf = (kKC / 4) * kLboA;   // bytes of one A image (hi or lo) per stage

__host__ __device__ inline size_t kpipe_stage_bytes(int N_t) { return (size_t)2 * kKpAHalf + (size_t)2 * N_t * kKC * 4; }
__host__ __device__ inline size_t kpipe_smem_bytes(int N_t, int stages) { return stages * kpipe_stage_bytes(N_t) + 32 * 8 + 16; }

// RC = row-class mode (see above), compiled separately so that the default instantiation carries none of it.
// LW = loader warps: 8 (default) or 16 (opt-in UNO_B200_KPIPE_LW16=1: the loader warps' serial instruction stream per chunk is
// what bounds this kernel, tools/kpipe_probe.py -- twice the warps, half the rows per thread).
// DRAM locality: a chunk is 128 bytes of each of 128 rows that lie a row pitch apart, i.e. 128 scattered 128-byte requests.  The
// loaders therefore issue their global loads for TWO consecutive chunks at a time (256 contiguous bytes of every row requested
// together) and the 16-byte loads carry the L2::256B prefetch hint.  Measured on B200 (Darcy step, all analysis launches):
// 2.14 ms one chunk at a time, 2.05 with the hint, 1.83 paired, 1.79 both.
// DBG = timing-probe instantiation (tools/kpipe_probe.py, switch kpipe_debug); the default <false> carries none of the probe branches
template <int LW = kKpLoadWarps, bool RC = false, bool DBG = false>
__global__ void __launch_bounds__((LW + kKpEpiWarps + 2) * 32, 1) kpipe_kernel(const KPipeParams p) {
    const int dbg = DBG ? p.debug : 0;
    static_assert(LW == 8 || LW == 16, "loader warps");
    constexpr int RB = LW * 4;            // 16-byte paths: row slots per pass (thread -> row ltid/8 + RB*i)
    constexpr int RP = 128 / RB;          // 16-byte paths: rows per thread
    constexpr int RP4 = 128 / LW;         // 4-byte path: rows per thread (warp w -> rows w + LW*i)
    constexpr int DEPTH = LW == 8 ? kKpDepth : KPIPE_LW16_DEPTH;   // chunks of global loads in flight per loader thread
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    const uint32_t b_half = (uint32_t)p.N_t * kKC * 4;
    const uint32_t stage_bytes = 2 * kKpAHalf + 2 * b_half;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;            // [S] loaders (+ bulk copy bytes) -> mma
    uint64_t* empty = bars + 8;       // [S] mma -> loaders
    uint64_t* d_full = bars + 16;     // [2]
    uint64_t* d_empty = bars + 18;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    constexpr int kMmaWarp = LW + kKpEpiWarps, kCopyWarp = kMmaWarp + 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], ((dbg & 128) ? LW : LW * 32) + 1);   // +1: the arrive.expect_tx of the B copy (probe bit 128: one arrival per warp)
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
        // ------------------------------------------------------------------ MMA issuer
        // One elected lane; every descriptor is the stage-0 descriptor advanced by constants (no per-MMA descriptor
        // construction, no multiplies): this thread's instruction stream bounds the kernel when K is long.
        {
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
            uint64_t a_st = a_hi0, b_st = b_hi0;                                 // descriptors of the current stage
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
                            if (ks < nks && !(dbg & 1)) {
                                mma_tf32(d_tmem, da, db, idesc, ks ? 1u : acc);
                                mma_tf32(d_tmem, da, db + b_lo_off, idesc, 1u);
                                mma_tf32(d_tmem, da + a_lo_off, db, idesc, 1u);
                            }
                            da += a_step;
                            db += b_step;
                        }
                        if (dbg & 32) mbar_arrive(&empty[s]);      // probe: a plain arrival instead of tcgen05.commit (with bit 1)
                        else tc_commit(&empty[s]);
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
        }
    } else if (warp < LW) {
        // ------------------------------------------------------------------ loaders: 256 threads, chunk = 128 rows x 32 k
        // All indexing is incremental (no divisions, one 64-bit multiply per chunk): the loaders are the
        // instruction-issue critical path of this kernel.
        const int ltid = threadIdx.x;
        long n_my_tiles = 0;
        for (long tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) ++n_my_tiles;
        const long total = n_my_tiles * NKC;   // chunks this CTA processes, in order
        // issue-side cursor (runs DEPTH-1 chunks ahead) and process-side cursor
        long i_tile = blockIdx.x;
        int i_kc = 0;
        int p_s = 0;
        uint32_t p_ph = 0;
        auto stage_prologue = [&]() -> uint8_t* {
            mbar_wait(&empty[p_s], p_ph ^ 1u);
            return smem + (size_t)p_s * stage_bytes;
        };
        auto stage_epilogue = [&]() {
            if (!(dbg & 16)) fence_proxy_async();
            if (dbg & 128) {                                  // probe: one arrival per warp
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[p_s]);
            } else mbar_arrive(&full[p_s]);
            if (++p_s == S) { p_s = 0; p_ph ^= 1u; }
        };
        if constexpr (RC) {
            // row-class path: 16-byte loads from the 16-byte boundary at or before the row start (see "row classes" above)
            const int cls = (int)(blockIdx.x & 3);
            const int sh = (int)(((long)cls * p.lda) & 3);
            const int kq = ltid & 7, rbase = ltid >> 3;
            const long stride32 = 4 * RB * p.lda;                // rows RB apart in the tile are 4*RB apart in the tensor
            const int k_end = p.k_valid + sh;                 // shifted index k' is real iff sh <= k' < k_end
            const uint32_t so = (uint32_t)kq * kLboA + (uint32_t)rbase * 16;
            float4 ring[DEPTH][RP];
            // issue cursor kept incrementally: the loaders' own instruction stream is what bounds this kernel (tools/kpipe_probe.py)
            const float* src = nullptr;                       // this thread's first row at its 16 bytes of the current chunk
            int rows_left = 0, k0 = kq * 4;                   // row (4*RB*i) valid iff 4*RB*i < rows_left
            auto seek = [&]() {
                const long row0 = (i_tile >> 2) * 512 + 4 * rbase + cls;
                src = p.A + row0 * p.lda - sh + kq * 4;
                rows_left = (int)max(min(p.R - row0, 1L << 20), 0L);
                k0 = kq * 4;
            };
            seek();
            auto issue = [&](float4 (&v)[RP]) {
                if (k0 >= sh && k0 + 4 <= k_end) {
#pragma unroll
                    for (int i = 0; i < RP; ++i)
                        v[i] = (4 * RB * i < rows_left) ? ldg_f4<true>(reinterpret_cast<const float4*>(src + i * stride32)) : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
#pragma unroll
                    for (int i = 0; i < RP; ++i) {
                        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (4 * RB * i < rows_left) {
                            const float* q = src + i * stride32;
                            if (k0 + 0 >= sh && k0 + 0 < k_end) v[i].x = __ldg(q + 0);
                            if (k0 + 1 >= sh && k0 + 1 < k_end) v[i].y = __ldg(q + 1);
                            if (k0 + 2 >= sh && k0 + 2 < k_end) v[i].z = __ldg(q + 2);
                            if (k0 + 3 >= sh && k0 + 3 < k_end) v[i].w = __ldg(q + 3);
                        }
                    }
                }
                src += kKC;
                k0 += kKC;
                if (++i_kc == NKC) { i_kc = 0; i_tile += gridDim.x; seek(); }
            };
            auto process = [&](const float4 (&v)[RP]) {
                uint8_t* st = stage_prologue() + so;
#pragma unroll
                for (int i = 0; i < RP; ++i) {
                    float4 hi, lo;
                    split_tf32(v[i].x, hi.x, lo.x);
                    split_tf32(v[i].y, hi.y, lo.y);
                    split_tf32(v[i].z, hi.z, lo.z);
                    split_tf32(v[i].w, hi.w, lo.w);
                    *reinterpret_cast<float4*>(st + i * (RB * 16)) = hi;
                    *reinterpret_cast<float4*>(st + kKpAHalf + i * (RB * 16)) = lo;
                }
                stage_epilogue();
            };
            {
                // chunks go to ring slot (chunk % 4); loads are issued for two consecutive chunks at once
                static_assert(DEPTH == 4, "ring of four chunks");
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
        } else if (p.a_vec_ok) {
            // 16-byte path: thread -> (row = ltid/8 + 32*i, 4 k at (ltid%8)*4)
            const int kq = ltid & 7, rbase = ltid >> 3;
            const long stride32 = RB * p.lda;
            const uint32_t so = (uint32_t)kq * kLboA + (uint32_t)rbase * 16;
            float4 ring[DEPTH][RP];
            // issue cursor kept incrementally (see the row-class path)
            const float* src = nullptr;
            int rows_left = 0, k0 = kq * 4;                   // row (RB*i) valid iff RB*i < rows_left
            auto seek = [&]() {
                const long row0 = i_tile * 128 + rbase;
                src = p.A + row0 * p.lda + kq * 4;
                rows_left = (int)max(min(p.R - row0, 1L << 20), 0L);
                k0 = kq * 4;
            };
            seek();
            auto issue = [&](float4 (&v)[RP]) {
                if (k0 + 4 <= p.K) {
                    // every chunk but a ragged last one: straight-line predicated 16-byte loads (the loader warps' serial
                    // instruction stream per chunk is what bounds this kernel, tools/kpipe_probe.py)
                    const int lim = (dbg & 8) ? 0 : rows_left;
#pragma unroll
                    for (int i = 0; i < RP; ++i)
                        v[i] = (RB * i < lim) ? ldg_f4<true>(reinterpret_cast<const float4*>(src + i * stride32)) : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
#pragma unroll
                    for (int i = 0; i < RP; ++i) {
                        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (RB * i < rows_left && !(dbg & 8)) {
                            const float* q = src + i * stride32;
                            if (k0 + 0 < p.K) v[i].x = __ldg(q + 0);
                            if (k0 + 1 < p.K) v[i].y = __ldg(q + 1);
                            if (k0 + 2 < p.K) v[i].z = __ldg(q + 2);
                        }
                    }
                }
                src += kKC;
                k0 += kKC;
                if (++i_kc == NKC) { i_kc = 0; i_tile += gridDim.x; seek(); }
            };
            auto process = [&](const float4 (&v)[RP]) {
                uint8_t* st = stage_prologue() + so;
#pragma unroll
                for (int i = 0; i < RP; ++i) {
                    if (dbg & 2) break;
                    float4 hi, lo;
                    split_tf32(v[i].x, hi.x, lo.x);
                    split_tf32(v[i].y, hi.y, lo.y);
                    split_tf32(v[i].z, hi.z, lo.z);
                    split_tf32(v[i].w, hi.w, lo.w);
                    *reinterpret_cast<float4*>(st + i * (RB * 16)) = hi;
                    *reinterpret_cast<float4*>(st + kKpAHalf + i * (RB * 16)) = lo;
                }
                stage_epilogue();
            };
            {
                // chunks go to ring slot (chunk % 4); loads are issued for two consecutive chunks at once
                static_assert(DEPTH == 4, "ring of four chunks");
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
        } else {
            // 4-byte path: lane = k within the chunk, warp w -> rows w + 8*i
            const long stride8 = LW * p.lda;
            const uint32_t so = (uint32_t)(lane >> 2) * kLboA + (uint32_t)(lane & 3) * 4 + (uint32_t)warp * 16;
            float ring[DEPTH][RP4];
            auto issue = [&](float (&v)[RP4]) {
                const long row0 = i_tile * 128 + warp;
                const int k = i_kc * kKC + lane;
                const float* src = p.A + row0 * p.lda + k;
                const long rows_left = (k < p.K) ? (p.R - row0) : 0;   // row (LW*i) valid iff LW*i < rows_left
#pragma unroll
                for (int i = 0; i < RP4; ++i) v[i] = (LW * i < rows_left && !(dbg & 8)) ? __ldg(src + i * stride8) : 0.f;
                if (++i_kc == NKC) { i_kc = 0; i_tile += gridDim.x; }
            };
            auto process = [&](const float (&v)[RP4]) {
                uint8_t* st = stage_prologue() + so;
#pragma unroll
                for (int i = 0; i < RP4; ++i) {
                    if (dbg & 2) break;
                    float hi, lo;
                    split_tf32(v[i], hi, lo);
                    *reinterpret_cast<float*>(st + i * (LW * 16)) = hi;
                    *reinterpret_cast<float*>(st + kKpAHalf + i * (LW * 16)) = lo;
                }
                stage_epilogue();
            };
            {
                // chunks go to ring slot (chunk % 4); loads are issued for two consecutive chunks at once
                static_assert(DEPTH == 4, "ring of four chunks");
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
        }
    } else if (warp == kCopyWarp) {
        // ------------------------------------------------------------------ B-chunk copies: one lane issues, per chunk, the bulk copy
        // of the twiddle image into the stage the MMAs have released.  (It used to be thread 0 of the loaders, inside their
        // per-chunk loop: the predicate and the reconvergence code around it cost every loader thread ~15 instructions per chunk,
        // and that instruction stream is what bounds this kernel.)
        if (lane == 0) {
            long n_my_tiles = 0;
            for (long tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) ++n_my_tiles;
            const long total = n_my_tiles * NKC;
            const uint32_t img_chunk_floats = 2 * b_half / 4;
            const float* bimg = p.Bimg + (RC ? (size_t)(blockIdx.x & 3) * NKC * img_chunk_floats : (size_t)0);
            int s = 0, kc = 0;
            uint32_t ph = 0;
            for (long g = 0; g < total; ++g) {
                mbar_wait(&empty[s], ph ^ 1u);
                if (dbg & 4) mbar_arrive(&full[s]);
                else {
                    mbar_arrive_expect_tx(&full[s], 2 * b_half);
                    bulk_g2s(smem + (size_t)s * stage_bytes + 2 * kKpAHalf, bimg + (size_t)kc * img_chunk_floats, 2 * b_half, &full[s]);
                }
                if (++s == S) { s = 0; ph ^= 1u; }
                if (++kc == NKC) kc = 0;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: one warp per TMEM lane quarter
        const int q = warp - LW;
        RowGemmParams ep;
        ep.C = p.C; ep.C2 = nullptr; ep.ldc = p.ldc; ep.N = p.N; ep.N_t = p.N_t; ep.R = p.R; ep.nclass = 1; ep.cls_log2 = 0; ep.shift_mul = 0;
        const bool vec2 = (p.ldc % 2 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 7) == 0);
        int it = 0;
        for (long tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait_relaxed(&d_full[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_base = tmem_base + (uint32_t)buf * buf_cols + ((uint32_t)(q * 32) << 16);
            if constexpr (RC) {
                // row-class mode: thread = one row of the class, float2 stores (N even, ldc even, C 8-byte aligned: host-checked)
                const long grow = (tile >> 2) * 512 + 4 * (q * 32 + lane) + (tile & 3);
                float* crow = p.C + (grow < p.R ? grow : 0) * p.ldc;
                for (int c0 = 0; c0 < p.N; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(t_base + (uint32_t)c0, v);
                    tmem_ld_wait();
                    if (grow < p.R) {
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (c0 + 2 * u < p.N)
                                *reinterpret_cast<float2*>(crow + c0 + 2 * u) = make_float2(__uint_as_float(v[2 * u]), __uint_as_float(v[2 * u + 1]));
                    }
                }
                tc_fence_before();
                mbar_arrive(&d_empty[buf]);
                continue;
            }
            // both column "halves" handled by this warp
            constexpr int EJ = LW == 8 ? 2 : 1;      // 16 loader warps leave 80 registers per thread: one chunk per round
            if (vec2) {
                rowgemm_epilogue_tile<EPI_STORE, true, 2, EJ>(ep, t_base, tile, q, 0, 0, lane);
                rowgemm_epilogue_tile<EPI_STORE, true, 2, EJ>(ep, t_base, tile, q, 0, 1, lane);
            } else {
                rowgemm_epilogue_tile<EPI_STORE, false, 2, EJ>(ep, t_base, tile, q, 0, 0, lane);
                rowgemm_epilogue_tile<EPI_STORE, false, 2, EJ>(ep, t_base, tile, q, 0, 1, lane);
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
