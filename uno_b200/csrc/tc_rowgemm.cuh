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
    int nclass;                   // 1, or 2 / 4 / 8 row classes: a tile holds rows of ONE class (row % nclass) and its B image is shifted (see below)
    int cls_log2;                 // log2(nclass)
    int shift_mul;                // column shift of class k = (k * shift_mul) & 7
};

// Odd leading dimension (the 481-wide Darcy grid): every other row of C starts 4 bytes off an 8-byte boundary, so
// the accumulator layout (a thread owns columns 2j, 2j+1) cannot be stored as float2 and the epilogue falls to
// half-filled 32-byte sectors.  In parity mode a tile is 128 rows of the SAME parity out of a block of 256
// (row = 256*(tile/2) + 2*i + tile%2), a CTA only ever sees one parity (even grid), and the CTAs of the odd rows
// load a twiddle image whose columns are shifted by one (accumulator column c = output column c-1): the pair a
// thread owns is then 8-byte aligned in every row and both parities take the float2 path.
//
// Row classes generalise this to 32-byte SECTORS: a lane quad owns eight consecutive columns = 32 bytes, which is one whole sector
// only if the row starts on a 32-byte boundary -- with a pitch of 481 floats the start moves by 4 bytes per row, so seven rows
// out of eight had every quad straddle two sectors (the 240^2 -> 481^2 forward launch of the Darcy model, which reads one and
// writes two such tensors, ran at 0.44 of HBM where the aligned 240-wide launches reach 0.65).  With nclass = 8 / gcd(ldc, 8)
// classes a tile is 128 rows of the same residue (row = 128*nclass*(tile/nclass) + nclass*i + tile%nclass), a CTA only sees one
// class (grid a multiple of nclass), and class k loads the twiddle image shifted by s_k = (k * ldc) mod 8 columns, so that
// accumulator column c is output column c - s_k and every quad of every row is exactly one sector.  When the column tiles
// have no room for a shift of seven (or C is not 32-byte aligned) odd pitches fall back to two classes with shift k.
__device__ __forceinline__ long rowgemm_row(const RowGemmParams& p, long tile, int i) {
    const int cs = p.cls_log2;                           // nclass = 1 << cls_log2
    return ((tile >> cs) << (7 + cs)) + ((long)i << cs) + (tile & (p.nclass - 1));
}
__device__ __forceinline__ int rowgemm_shift(const RowGemmParams& p, long tile) {
    return p.nclass > 1 ? (int)(((tile & (p.nclass - 1)) * p.shift_mul) & 7) : 0;
}

// Epilogue warps come in groups of four (one warp per TMEM lane quarter); the G groups split the 16-column chunks of a tile
// between them and a warp handles J chunks per round.  Default G = 2, J = 2 (8 warps, 32 x 32 elements per warp and round).
// Opt-in (UNO_B200_ROWGEMM_EPI16=1): G = 4, J = 1 -- 16 warps of 32 x 16 elements per round, i.e. the same bytes in flight
// spread over twice the warps, so that four warps per scheduler instead of two cover each other's load latency (the
// accumulate launches of this kernel sit at 17 % warp occupancy with 4-6 long-scoreboard stalls per issue, profiles/
// r01_ncu_full_v43_synthesis_resample.txt) within the 107 registers per thread that 608 threads leave.
constexpr int kRowGemmLoadWarps = 2;
__host__ __device__ constexpr int rowgemm_epi_warps(int G) { return 4 * G; }
__host__ __device__ constexpr int rowgemm_threads(int G) { return (4 * G + kRowGemmLoadWarps + 1) * 32; }
constexpr int kRowGemmEpiWarps = rowgemm_epi_warps(2);
constexpr int kRowGemmThreads = rowgemm_threads(2);
constexpr uint32_t kLboA = 128 * 16 + 16;   // +16 B: the loaders' 16-byte stores of one warp fall in distinct bank groups

__host__ __device__ inline size_t rowgemm_smem_bytes(int K_pad, int N_t) {
    const size_t b = (size_t)2 * K_pad * N_t * 4;
    const size_t a = (size_t)4 * (K_pad / 4) * kLboA;
    return b + a + 16 * 8 + 16;
}

// libm erff here on purpose.  Measured on B200 per Darcy step (this kernel's launches): erff 3.67 ms; the one-ex2 form the SIMT
// kernels use (tc_common.cuh gelu_fwd_fast, a third of the instructions) 3.74 ms; the single-exponential Abramowitz-Stegun form of
// the backward kernels (two MUFU operations per element) 3.88 ms.  The epilogue warps are bound by the latency of their addend
// loads, not by issue slots, and every special-function operation lengthens the dependent chain of a round, while erff is almost
// all FMA-pipe work.
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Epilogue of one 128-row tile for one warp: TMEM lane quarter `q`, 16-column chunks c0 = 16*(2*i + half).
// All loads of a 32-column group are issued before any store so that they are in flight together; the four
// row pointers of a thread are formed once per tile (no 64-bit multiplies in the column loop).
template <int EPI, bool VEC2, int G = 2, int J = 2>
__device__ __forceinline__ void rowgemm_epilogue_tile(const RowGemmParams& p, uint32_t t_base, long tile, int q, int n_base, int half, int lane,
                                                      bool skip_full = false) {
    static_assert(J == 1 || J == 2, "one or two 16-column chunks per round");
    constexpr bool kAccum = (EPI != EPI_STORE);
    const long c2_delta = (EPI == EPI_ACCUM_GELU) ? (p.C2 - p.C) : 0;
    float* rowp[4];          // index hh*2 + rr
    bool rok[4];
    n_base -= rowgemm_shift(p, tile);                    // accumulator column c of this tile = output column n_base + c
#pragma unroll
    for (int hr = 0; hr < 4; ++hr) {
        const long grow = rowgemm_row(p, tile, q * 32 + (hr >> 1) * 16 + (hr & 1) * 8 + (lane >> 2));
        rok[hr] = grow < p.R;
        rowp[hr] = p.C + (rok[hr] ? grow : 0) * p.ldc + n_base + 2 * (lane & 3);
    }
    const int ncol = p.N - n_base - 2 * (lane & 3);     // column c (relative) valid iff cmin <= c < ncol
    const int cmin = -(n_base + 2 * (lane & 3));        // > 0 for the first threads of a column-shifted first tile (shift up to 7), else <= 0
    for (int ci = half; ci * 16 < p.N_t; ci += J * G) {
        const int c0a = ci * 16, c0b = (ci + G) * 16;
        const bool has_b = J == 2 && c0b < p.N_t && n_base + c0b < p.N;
        if (n_base + c0a >= p.N) break;
        if (skip_full && n_base + c0a >= 0 && n_base + c0a + 16 <= p.N) continue;   // done by rowgemm_epilogue_tile_fast
        uint32_t r[2 * J][8];
        tmem_ld_16x256b_x2(t_base + (uint32_t)c0a, r[0]);
        tmem_ld_16x256b_x2(t_base + (16u << 16) + (uint32_t)c0a, r[1]);
        if (J == 2 && has_b) {
            tmem_ld_16x256b_x2(t_base + (uint32_t)c0b, r[2 * J - 2]);
            tmem_ld_16x256b_x2(t_base + (16u << 16) + (uint32_t)c0b, r[2 * J - 1]);
        }
        // element pair e = (chunk j, half hh, repeat rep, row-pair rr)
        float2 cz[8 * J];
        if (kAccum) {
#pragma unroll
            for (int e = 0; e < 8 * J; ++e) {
                const int j = e >> 3, hh = (e >> 2) & 1, rep = (e >> 1) & 1, rr = e & 1;
                const int c = (j ? c0b : c0a) + rep * 8;
                const bool live = (j == 0 || has_b) && rok[hh * 2 + rr];
                const float* q = rowp[hh * 2 + rr] + c;
                cz[e] = make_float2(0.f, 0.f);
                const bool in0 = live && c >= cmin && c < ncol, in1 = live && c + 1 >= cmin && c + 1 < ncol;
                // (an L2::256B prefetch hint on these addend loads measured 3.66 against 3.70 ms per Darcy step: not kept)
                if (VEC2 && in0 && in1) cz[e] = *reinterpret_cast<const float2*>(q);
                else {
                    if (in0) cz[e].x = q[0];
                    if (in1) cz[e].y = q[1];
                }
            }
        }
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8 * J; ++e) {
            const int j = e >> 3, hh = (e >> 2) & 1, rep = (e >> 1) & 1, rr = e & 1;
            const int c = (j ? c0b : c0a) + rep * 8;
            const bool live = (j == 0 || has_b) && rok[hh * 2 + rr];
            const bool ok0 = live && c >= cmin && c < ncol, ok1 = live && c + 1 >= cmin && c + 1 < ncol;
            float* q = rowp[hh * 2 + rr] + c;
            float2 acc = make_float2(__uint_as_float(r[j * 2 + hh][rep * 4 + rr * 2 + 0]), __uint_as_float(r[j * 2 + hh][rep * 4 + rr * 2 + 1]));
            if (kAccum) { acc.x += cz[e].x; acc.y += cz[e].y; }
            float2 act = acc;
            if (EPI == EPI_ACCUM_GELU || EPI == EPI_ACCUM_GELU_INPLACE) { act.x = gelu_erf(acc.x); act.y = gelu_erf(acc.y); }
            const float2 out1 = (EPI == EPI_ACCUM_GELU_INPLACE) ? act : acc;
            if (VEC2 && ok0 && ok1) {
                *reinterpret_cast<float2*>(q) = out1;
                if (EPI == EPI_ACCUM_GELU) *reinterpret_cast<float2*>(q + c2_delta) = act;
            } else {
                if (ok0) { q[0] = out1.x; if (EPI == EPI_ACCUM_GELU) q[c2_delta] = act.x; }
                if (ok1) { q[1] = out1.y; if (EPI == EPI_ACCUM_GELU) q[c2_delta + 1] = act.y; }
            }
        }
    }
}

// Fast path of the epilogue (J = 1, 8-byte aligned pairs, all 128 rows of the tile inside the matrix): the 16-column chunks that
// lie completely inside the output run without a single bounds predicate -- straight-line 8-byte loads and stores off four row
// pointers -- and the addend of the warp's NEXT chunk is requested before the current one is combined, so that its latency
// overlaps the GELU arithmetic and the stores.  Why: ncu (profiles/r02_ncu_full_v51_darcy.txt) has 43 % of this kernel's stall
// samples on the first use of the addend, and the predicated form spends ~20 control instructions (BSSY / BRA / BSYNC) per load,
// ~450 per thread and round, before and between those loads; a microbenchmark of the SAME access pattern without them streams
// 5.65 TB/s from one 512-thread CTA per SM (tools/microbench/epilogue_patterns.cu) against 3.4 - 3.8 TB/s here.  Measured on B200:
// 3.76 -> 3.58 ms per Darcy step for this kernel's launches, 0.64 of the HBM peak under ncu (0.58 before).  Chunks cut by the
// matrix edge (and the first chunk of a column-shifted parity tile) are left to the general form above.
// Returns false (nothing done, accumulator not yet awaited) when the tile does not qualify.
template <int EPI, int G>
__device__ __forceinline__ bool rowgemm_epilogue_tile_fast(const RowGemmParams& p, uint32_t t_base, long tile, int q, int n_base, int half,
                                                           int lane, uint64_t* d_full, uint32_t ph) {
    constexpr bool kAccum = (EPI != EPI_STORE);
    if (rowgemm_row(p, tile, 127) >= p.R) return false;                  // warp-uniform: a ragged last tile
    const int nb = n_base - rowgemm_shift(p, tile);                      // accumulator column c = output column nb + c
    const long c2_delta = (EPI == EPI_ACCUM_GELU) ? (p.C2 - p.C) : 0;
    float* rowp[4];                                                      // index hh*2 + rr, as in the general form
#pragma unroll
    for (int hr = 0; hr < 4; ++hr)
        rowp[hr] = p.C + rowgemm_row(p, tile, q * 32 + (hr >> 1) * 16 + (hr & 1) * 8 + (lane >> 2)) * p.ldc + nb + 2 * (lane & 3);
    auto full = [&](int ci) { return ci * 16 < p.N_t && nb + ci * 16 >= 0 && nb + ci * 16 + 16 <= p.N; };
    auto load = [&](int ci, float2 (&cz)[8]) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int hh = e >> 2, rep = (e >> 1) & 1, rr = e & 1;
            cz[e] = *reinterpret_cast<const float2*>(rowp[hh * 2 + rr] + ci * 16 + rep * 8);
        }
    };
    // the warp's chunks are ci = half, half + G, ...; `cur` walks the full ones
    int cur = half;
    while (cur * 16 < p.N_t && !full(cur)) cur += G;
    float2 cz[8], nz[8];
    const bool any = cur * 16 < p.N_t;
    if (kAccum && any) load(cur, cz);                                    // in flight while the MMA of this tile finishes
    mbar_wait_relaxed(d_full, ph);
    tc_fence_after();
    while (cur * 16 < p.N_t) {
        int nxt = cur + G;
        while (nxt * 16 < p.N_t && !full(nxt)) nxt += G;
        uint32_t r[2][8];
        tmem_ld_16x256b_x2(t_base + (uint32_t)(cur * 16), r[0]);
        tmem_ld_16x256b_x2(t_base + (16u << 16) + (uint32_t)(cur * 16), r[1]);
        if (kAccum && nxt * 16 < p.N_t) load(nxt, nz);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int hh = e >> 2, rep = (e >> 1) & 1, rr = e & 1;
            float2 acc = make_float2(__uint_as_float(r[hh][rep * 4 + rr * 2 + 0]), __uint_as_float(r[hh][rep * 4 + rr * 2 + 1]));
            if (kAccum) { acc.x += cz[e].x; acc.y += cz[e].y; }
            float2 act = acc;
            if (EPI == EPI_ACCUM_GELU || EPI == EPI_ACCUM_GELU_INPLACE) { act.x = gelu_erf(acc.x); act.y = gelu_erf(acc.y); }
            float* dst = rowp[hh * 2 + rr] + cur * 16 + rep * 8;
            *reinterpret_cast<float2*>(dst) = (EPI == EPI_ACCUM_GELU_INPLACE) ? act : acc;
            if (EPI == EPI_ACCUM_GELU) *reinterpret_cast<float2*>(dst + c2_delta) = act;
        }
        if (kAccum) {
#pragma unroll
            for (int e = 0; e < 8; ++e) cz[e] = nz[e];
        }
        cur = nxt;
    }
    return true;
}

// (A software-pipelined form of this epilogue that loaded the next round's addend before storing the current round was measured
// on B200 and was no faster -- 3.68 vs 3.62 ms per Darcy step -- so it was removed.)
template <int EPI, int G = 2, int J = 2>
__global__ void __launch_bounds__(rowgemm_threads(G), 1) rowgemm_smallk_kernel(const RowGemmParams p) {
    constexpr int kRowGemmEpiWarps = rowgemm_epi_warps(G);   // shadows the default-configuration constant
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
            const int shifted = rowgemm_shift(p, blockIdx.x);               // gridDim.x is a multiple of nclass: one class per CTA
            bulk_g2s(sB, p.Bimg + (size_t)(shifted * p.n_tiles + nt) * (2 * b_half / 4), 2 * b_half, b_full);
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
        // ------------------------------------------------------------------ A loaders: thread -> rows t and t+64, loop over the k chunks
        constexpr int NL = kRowGemmLoadWarps * 32;
        constexpr int BATCH = 6;
        const int ltid = threadIdx.x - kRowGemmEpiWarps * 32;
        int it = 0;
        for (long tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
            const int s = it & 1;
            const uint32_t ph = (uint32_t)(it >> 1) & 1u;
            mbar_wait(&a_empty[s], ph ^ 1u);
            uint8_t* dst_hi = sA + (size_t)(s * 2 + 0) * a_bytes;
            uint8_t* dst_lo = sA + (size_t)(s * 2 + 1) * a_bytes;
#pragma unroll
            for (int rr = 0; rr < 128 / NL; ++rr) {
                const int row = ltid + rr * NL;
                const long grow = rowgemm_row(p, tile, row);
                const bool rvalid = grow < p.R;
                const float* src = p.A + (rvalid ? grow : 0) * p.lda;
                const uint32_t ro = (uint32_t)row * 16;
                for (int kc0 = 0; kc0 < CH; kc0 += BATCH) {
                    float4 v[BATCH];
#pragma unroll
                    for (int u = 0; u < BATCH; ++u) {
                        const int kc = kc0 + u;
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (kc < CH && rvalid) {
                            const float* q = src + kc * 4;
                            if (p.a_vec_ok && kc * 4 + 4 <= p.K) {
                                v[u] = __ldg(reinterpret_cast<const float4*>(q));
                            } else {
                                if (kc * 4 + 0 < p.K) v[u].x = __ldg(q + 0);
                                if (kc * 4 + 1 < p.K) v[u].y = __ldg(q + 1);
                                if (kc * 4 + 2 < p.K) v[u].z = __ldg(q + 2);
                                if (kc * 4 + 3 < p.K) v[u].w = __ldg(q + 3);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < BATCH; ++u) {
                        const int kc = kc0 + u;
                        if (kc >= CH) break;
                        float4 hi, lo;
                        split_tf32(v[u].x, hi.x, lo.x);
                        split_tf32(v[u].y, hi.y, lo.y);
                        split_tf32(v[u].z, hi.z, lo.z);
                        split_tf32(v[u].w, hi.w, lo.w);
                        const uint32_t o = (uint32_t)kc * kLboA + ro;
                        *reinterpret_cast<float4*>(dst_hi + o) = hi;
                        *reinterpret_cast<float4*>(dst_lo + o) = lo;
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(&a_full[s]);
        }
    } else {
        // ------------------------------------------------------------------ epilogue: warp e -> TMEM lane quarter e%4, column half e/4
        const bool vec2 = (p.ldc % 2 == 0 || p.nclass > 1) && ((reinterpret_cast<uintptr_t>(p.C) & 7) == 0) &&
                          (EPI != EPI_ACCUM_GELU || (reinterpret_cast<uintptr_t>(p.C2) & 7) == 0);
        const int q = warp & 3, half = warp >> 2;      // half = column group 0 .. G-1
        const int n_base = nt * p.N_t;
        int it = 0;
        for (long tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++it) {
            const int s = it & 1;
            const uint32_t ph = (uint32_t)(it >> 1) & 1u;
            const uint32_t t_base = tmem_base + (uint32_t)s * buf_cols + ((uint32_t)(q * 32) << 16);
            bool fast = false;
            if (J == 1 && vec2) fast = rowgemm_epilogue_tile_fast<EPI, G>(p, t_base, tile, q, n_base, half, lane, &d_full[s], ph);
            if (!fast) {
                mbar_wait_relaxed(&d_full[s], ph);
                tc_fence_after();
            }
            // the general form: the whole tile, or only the chunks the fast path left (those cut by the matrix edge)
            if (vec2) rowgemm_epilogue_tile<EPI, true, G, J>(p, t_base, tile, q, n_base, half, lane, fast);
            else rowgemm_epilogue_tile<EPI, false, G, J>(p, t_base, tile, q, n_base, half, lane, fast);
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
