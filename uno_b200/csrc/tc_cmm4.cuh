// tcgen05 per-mode complex channel contraction, FOUR consecutive modes per work item (SpectralConv*.compl_mul*,
// integral_operators.py:45, :179, :383):
//     C[m, n, q] = sum_k opA(A[m, k, q]) * opB(B[k, n, q])          one small complex GEMM per kept mode q
//
// Same real expansion as tc_cmm.cuh (the weight rows (n, re|im) on the UMMA M side, the samples on the N side, k' = (k, re|im)),
// but the kept-mode index q is the contiguous one of all three tensors, so a kernel that handles ONE mode per item touches
// 8 bytes of every 32-byte sector it pulls in -- loads of A and B and the stores of C alike.  Here an item is a group of four
// consecutive modes: every global access is a whole sector (two neighbouring lanes take 16 bytes each), and the four GEMMs run
// side by side:
//   * a stage = one UMMA k-step (4 complex k = 8 real k') of all four modes: 4 x (128 x 8) weight images and 4 x (N_t x 8)
//     sample images, tf32 hi / lo, K-major "interleave" layout, written by the loader warps from 16-byte loads;
//   * 12 MMAs per stage (3xTF32 for each mode) into four TMEM accumulators (N_t columns each), double buffered across items;
//   * the epilogue pairs the (re, im) rows of neighbouring lanes and the four accumulators with two shuffles per sample and
//     stores 16 bytes per lane: 32 contiguous bytes per (sample, channel).
// Requires 16-byte friendly operands (all strides and corner offsets even in complex elements); tc_cmm.cuh takes the rest.
//
// An item streams its weights through ONE SM: 128 input channels of four modes are 1 MB and take 30 - 40 us (latency-bound at
// ~30 GB/s per CTA; neither the loader warps' instruction count nor the MMA count is the limit -- halving either measured no
// change).  Levels with few modes have fewer items than SMs (32 - 64 at the inner NS-2D levels), so the launcher splits the
// reduction of every output tile over `ksplit` items that add into a zeroed C (cmm_zero_kernel): the same stream on two to
// eight times as many SMs.
#pragma once
#include "backend.h"
#include "tc_common.cuh"
#include "tc_kpipe.cuh"

namespace uno {
namespace tc {

struct Cmm4Params {
    CmmArgs a;
    int N_t;          // sample columns per tile, multiple of 16, <= 64
    int ns_tiles;     // column tiles over the M samples
    int ms_tiles;     // 128-row tiles over the 2*N rows (n, re|im)
    int n_chunks;     // k-steps (4 complex k) PER ITEM: ceil(K / 4) / ksplit
    int ksplit;       // the reduction of one output tile is split over this many items, which then ADD into a pre-zeroed C
    int qg;           // ceil(q_inner / 4) mode groups
    int stages, tmem_cols;
    long items;       // ms_tiles * ns_tiles * ncorner * q_outer * qg * ksplit
};

constexpr int kC4Modes = 4;
constexpr int kC4LoadWarps = 16;                             // warps 0-7 stage the weight images, 8-15 the sample images
constexpr int kC4Threads = (kC4LoadWarps + kKpEpiWarps + 1) * 32;
constexpr int kC4Depth = 4;                                  // chunks of global loads in flight per loader thread (six spill at 80 registers: 0.99 -> 1.19 ms per NS-2D call)
constexpr uint32_t kC4AHalf = 2 * kLboA;                     // one weight image (hi or lo): two 4-wide k' groups of 128 rows
constexpr uint32_t kC4AMode = 2 * kC4AHalf;                  // hi + lo

// sample image of one mode: hi + lo, two k' groups of N_t 16-byte rows each, + 32 bytes so that the two lanes of a pair (which
// write the same row of images two modes apart) fall in different bank groups
__host__ __device__ inline uint32_t cmm4_b_mode(int N_t) { return (uint32_t)64 * N_t + 32; }
__host__ __device__ inline size_t cmm4_stage_bytes(int N_t) { return (size_t)kC4Modes * (kC4AMode + cmm4_b_mode(N_t)); }
__host__ __device__ inline size_t cmm4_smem_bytes(int N_t, int stages) { return stages * cmm4_stage_bytes(N_t) + 32 * 8 + 16; }

struct Cmm4Item {
    int ms, ns, corner, qo, qi0, kc0;     // kc0: first k-step of this item's share of the reduction
};
__device__ __forceinline__ Cmm4Item cmm4_item(const Cmm4Params& p, long w) {
    Cmm4Item it;
    it.kc0 = (int)(w % p.ksplit) * p.n_chunks; w /= p.ksplit;
    it.qi0 = 4 * (int)(w % p.qg); w /= p.qg;
    it.qo = (int)(w % p.a.q_outer); w /= p.a.q_outer;
    it.corner = (int)(w % p.a.ncorner); w /= p.a.ncorner;
    it.ns = (int)(w % p.ns_tiles);
    it.ms = (int)(w / p.ns_tiles);
    return it;
}

__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}

// C = 0 over exactly the elements the contraction writes (the corners are sub-blocks of a larger spectrum tensor)
__global__ void cmm_zero_kernel(const CmmArgs a) {
    const long per = (long)a.M * a.N * a.q_outer * a.q_inner;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < per * a.ncorner; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i / per);
        long r = i - (long)c * per;
        const int qi = (int)(r % a.q_inner); r /= a.q_inner;
        const int qo = (int)(r % a.q_outer); r /= a.q_outer;
        const int n = (int)(r % a.N);
        const int m = (int)(r / a.N);
        reinterpret_cast<float2*>(a.C[c])[(long)m * a.c_sm + (long)n * a.c_sn + (long)qo * a.c_sqo + qi] = make_float2(0.f, 0.f);
    }
}

__global__ void __launch_bounds__(kC4Threads, 1) cmm_tc4_kernel(const Cmm4Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    const uint32_t b_half = (uint32_t)p.N_t * 32;            // one sample image (hi or lo): two k' groups of N_t rows
    const uint32_t b_mode = cmm4_b_mode(p.N_t);
    const uint32_t stage_bytes = kC4Modes * (kC4AMode + b_mode);
    const uint32_t b_base = kC4Modes * kC4AMode;             // sample images follow the four weight images
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* full = bars;            // [S] loaders -> mma
    uint64_t* empty = bars + 8;       // [S] mma -> loaders
    uint64_t* d_full = bars + 16;     // [2]
    uint64_t* d_empty = bars + 18;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    constexpr int kMmaWarp = kC4LoadWarps + kKpEpiWarps;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], kC4LoadWarps * 32);
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
    const uint32_t buf_cols = (uint32_t)(kC4Modes * p.N_t);
    const int NKC = p.n_chunks;

    if (warp == kMmaWarp) {
        // ------------------------------------------------------------------ MMA issuer: 4 modes x 3 products per stage
        const uint32_t idesc = make_idesc_tf32(128, p.N_t, 0, 0);
        const uint32_t smem_base = smem_u32(smem);
        const uint64_t a0 = make_smem_desc(smem_base, kLboA, 128);
        const uint64_t b0 = make_smem_desc(smem_base + b_base, (uint32_t)p.N_t * 16, 128);
        const uint32_t a_lo = kC4AHalf >> 4, b_lo = b_half >> 4, a_md = kC4AMode >> 4, b_md = b_mode >> 4, st_step = stage_bytes >> 4;
        int s = 0;
        uint32_t ph = 0;
        uint64_t a_st = a0, b_st = b0;
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
                if (elect_one()) {
                    uint64_t da = a_st, db = b_st;
                    uint32_t dt = d_tmem;
#pragma unroll
                    for (int j = 0; j < kC4Modes; ++j) {
                        mma_tf32(dt, da, db, idesc, acc);
                        mma_tf32(dt, da, db + b_lo, idesc, 1u);
                        mma_tf32(dt, da + a_lo, db, idesc, 1u);
                        da += a_md;
                        db += b_md;
                        dt += (uint32_t)p.N_t;
                    }
                    tc_commit(&empty[s]);
                }
                acc = 1u;
                __syncwarp();
                a_st += st_step;
                b_st += st_step;
                if (++s == S) { s = 0; ph ^= 1u; a_st = a0; b_st = b0; }
            }
            if (elect_one()) tc_commit(&d_full[buf]);
            __syncwarp();
        }
    } else if (warp < kC4LoadWarps) {
        // ------------------------------------------------------------------ loaders
        // Warps 0-7 stage the weights, warps 8-15 the samples (one group doing both was bound by its own instruction stream:
        // ~950 cycles per stage for 24 KB; measured on B200 per NS-2D call: 1.01 -> 0.99 ms -- the kernel is bound by the
        // latency of one CTA's weight stream, 30 - 40 us per 128-channel item, not by these instructions nor by its MMA count: stacking
        // the hi / lo sample images along N, 8 instead of 12 MMAs per stage, measured no change).  Thread pair
        // (2t, 2t+1) shares a task; lane parity `hf` picks the mode half (modes 2hf, 2hf+1): 16 bytes per lane, one 32-byte
        // sector per pair.  Weight task t: channel row pair wn = t % 64 of the tile, k pair wkp = t / 64.  Sample task
        // t < 2*N_t: sample xm = t % N_t, k pair xkp = t / N_t.  A task = the elements k, k+1 of its k pair.
        const bool is_w = warp < 8;                          // warp-uniform role
        const int ltid = threadIdx.x & 255;
        const int pair = ltid >> 1, hf = ltid & 1;
        const float sa = p.a.conjA ? -1.f : 1.f, sb = p.a.conjB ? -1.f : 1.f;
        long n_my = 0;
        for (long w = blockIdx.x; w < p.items; w += gridDim.x) ++n_my;
        const long total = n_my * NKC;
        const int wn = pair & 63, wkp = pair >> 6;
        const bool x_task = !is_w && pair < 2 * p.N_t;
        const int xkp = x_task ? pair / p.N_t : 0;
        const int xm = pair - xkp * p.N_t;
        const int kp = is_w ? wkp : xkp;                     // this thread's k pair inside a k-step
        // image offsets of this thread's 16-byte rows inside a stage, for its first mode (the second is one mode image further)
        const uint32_t w_so = (uint32_t)(2 * hf) * kC4AMode + (uint32_t)wkp * kLboA + (uint32_t)(2 * wn) * 16;
        const uint32_t x_so = b_base + (uint32_t)(2 * hf) * b_mode + (uint32_t)xkp * (uint32_t)(p.N_t * 16) + (uint32_t)xm * 16;
        long i_w = blockIdx.x;
        int i_kc = 0;
        const float2* g0 = nullptr;       // weights: &B[k = 0, n of this thread, first mode]; samples: &A[m of this thread, k = 0, first mode]
        long g_sk = 0;                    // stride of k in that tensor
        int nmodes = 0;                   // how many of this thread's two modes exist (ragged last group)
        int i_kc0 = 0;                    // first k-step of the item (split reductions)
        auto seek = [&]() {
            g0 = nullptr;
            nmodes = 0;
            if (i_w >= p.items) return;
            const Cmm4Item it = cmm4_item(p, i_w);
            i_kc0 = it.kc0;
            const int q = it.qi0 + 2 * hf;
            nmodes = min(2, max(0, p.a.q_inner - q));
            if (nmodes == 0) return;
            if (is_w) {
                const int n = it.ms * 64 + wn;
                if (n < p.a.N) g0 = reinterpret_cast<const float2*>(p.a.B[it.corner]) + (long)it.qo * p.a.b_sqo + q + (long)n * p.a.b_sn;
            } else {
                const int m = it.ns * p.N_t + xm;
                if (x_task && m < p.a.M) g0 = reinterpret_cast<const float2*>(p.a.A[it.corner]) + (long)it.qo * p.a.a_sqo + q + (long)m * p.a.a_sm;
            }
        };
        g_sk = is_w ? p.a.b_sk : p.a.a_sk;
        seek();
        auto ld2 = [&](const float2* q) -> float4 {          // modes (q, q+1) of one element; the second may not exist
            if (nmodes == 2) return __ldg(reinterpret_cast<const float4*>(q));
            const float2 e = __ldg(q);
            return make_float4(e.x, e.y, 0.f, 0.f);
        };
        struct Slot { float4 e[2]; };                        // [k, k+1] of the k pair, each (re, im) of the thread's two modes
        Slot ring[kC4Depth];
        auto issue = [&](Slot& v) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const int k = (i_kc0 + i_kc) * 4 + 2 * kp;
            v.e[0] = (g0 && k < p.a.K) ? ld2(g0 + (long)k * g_sk) : z;
            v.e[1] = (g0 && k + 1 < p.a.K) ? ld2(g0 + (long)(k + 1) * g_sk) : z;
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
        int p_s = 0;
        uint32_t p_ph = 0;
        auto process = [&](const Slot& v) {
            mbar_wait(&empty[p_s], p_ph ^ 1u);
            uint8_t* st = smem + (size_t)p_s * stage_bytes;
            const float4 e0 = v.e[0], e1 = v.e[1];
            if (is_w) {
                // weight rows (n, re) and (n, im) of both modes: columns (k re, k im, k+1 re, k+1 im).  The odd lane of a pair
                // writes its (n, im) row while the even lane writes its (n, re) row (images two modes apart alias in the banks)
                uint8_t* d = st + w_so;
                const float4 re_a = make_float4(e0.x, -sb * e0.y, e1.x, -sb * e1.y), im_a = make_float4(sb * e0.y, e0.x, sb * e1.y, e1.x);
                const float4 re_b = make_float4(e0.z, -sb * e0.w, e1.z, -sb * e1.w), im_b = make_float4(sb * e0.w, e0.z, sb * e1.w, e1.z);
                const uint32_t o1 = hf ? 16u : 0u, o2 = hf ? 0u : 16u;
                put(d + o1, kC4AHalf, hf ? im_a : re_a);
                put(d + o2, kC4AHalf, hf ? re_a : im_a);
                put(d + kC4AMode + o1, kC4AHalf, hf ? im_b : re_b);
                put(d + kC4AMode + o2, kC4AHalf, hf ? re_b : im_b);
            } else if (x_task) {
                uint8_t* d = st + x_so;
                put(d, b_half, make_float4(e0.x, sa * e0.y, e1.x, sa * e1.y));
                put(d + b_mode, b_half, make_float4(e0.z, sa * e0.w, e1.z, sa * e1.w));
            }
            fence_proxy_async();
            mbar_arrive(&full[p_s]);
            if (++p_s == S) { p_s = 0; p_ph ^= 1u; }
        };
#pragma unroll
        for (int d = 0; d < kC4Depth - 1; ++d)
            if (d < total) issue(ring[d]);
        for (long g = 0; g < total; g += kC4Depth) {
#pragma unroll
            for (int d = 0; d < kC4Depth; ++d) {
                if (g + d + kC4Depth - 1 < total) issue(ring[(d + kC4Depth - 1) % kC4Depth]);
                if (g + d < total) process(ring[d]);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: one warp per TMEM lane quarter
        // lane pair (2t, 2t+1) holds the (re, im) rows of one channel n in each of the four accumulators; after two exchanges
        // the even lane owns modes 0, 1 and the odd lane modes 2, 3 of (sample, n) as (re, im, re, im): one 16-byte store each
        const int q = warp - kC4LoadWarps;
        const int odd = lane & 1;
        int it = 0;
        for (long w = blockIdx.x; w < p.items; w += gridDim.x, ++it) {
            const int buf = it & 1;
            const Cmm4Item im = cmm4_item(p, w);
            mbar_wait_relaxed(&d_full[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_base = tmem_base + (uint32_t)buf * buf_cols + ((uint32_t)(q * 32) << 16);
            const int n = im.ms * 64 + ((q * 32 + lane) >> 1);
            const bool nok = n < p.a.N;
            const int m0 = im.ns * p.N_t;
            const int q0 = im.qi0 + 2 * odd;                               // first of the two modes this lane stores
            const int nst = min(2, max(0, p.a.q_inner - q0));              // how many of them exist
            float2* crow = reinterpret_cast<float2*>(p.a.C[im.corner]) + (long)im.qo * p.a.c_sqo + q0 + (long)(nok ? n : 0) * p.a.c_sn;
            const int mcols = min(p.N_t, p.a.M - m0);
            for (int c0 = 0; c0 < mcols; c0 += 8) {
                uint32_t v[kC4Modes][8];
#pragma unroll
                for (int j = 0; j < kC4Modes; ++j) tmem_ld_32x32b_x8(t_base + (uint32_t)(j * p.N_t + c0), v[j]);
                tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    // even lane: re of modes 0..3; odd lane: im of modes 0..3
                    const float s1 = __uint_as_float(odd ? v[0][u] : v[2][u]);      // odd sends im0, even sends re2
                    const float s2 = __uint_as_float(odd ? v[1][u] : v[3][u]);      // odd sends im1, even sends re3
                    const float r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
                    const float r2 = __shfl_xor_sync(0xffffffffu, s2, 1);
                    const float4 out = odd ? make_float4(r1, __uint_as_float(v[2][u]), r2, __uint_as_float(v[3][u]))
                                           : make_float4(__uint_as_float(v[0][u]), r1, __uint_as_float(v[1][u]), r2);
                    const int c = c0 + u;
                    if (nok && c < mcols && nst > 0) {
                        float2* dst = crow + (long)(m0 + c) * p.a.c_sm;
                        if (p.ksplit > 1) {               // partial sum of a split reduction: C was zeroed by the launcher
                            float* d = reinterpret_cast<float*>(dst);
                            atomicAdd(d, out.x);
                            atomicAdd(d + 1, out.y);
                            if (nst == 2) { atomicAdd(d + 2, out.z); atomicAdd(d + 3, out.w); }
                        } else if (nst == 2) *reinterpret_cast<float4*>(dst) = out;
                        else *dst = make_float2(out.x, out.y);
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
