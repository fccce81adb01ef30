// Hand-written sm_100a kernels behind uno_b200/csrc/backend.h.
//
// Kernel inventory (DESIGN.md has the roofline each one is bound by):
//   gemm_kernel        C = A*B (+epilogue)   truncated-DFT analysis/synthesis along the last axis,
//                                             1x1 channel mix, weight-gradient reduction (split-K, atomics)
//   mid2_kernel        complex [J x H] transform along a leading axis, batched, register-tiled
//   cmm_kernel         per-mode complex channel contraction (mode index on the lanes)
//   banded_kernel      anti-aliased bicubic resample bands (and their transposes)
//   elementwise / plane reductions: GELU fwd/bwd, InstanceNorm stats / apply / backward, bias sums
//
// Everything here is fp32 FMA on the SIMT pipes with fp32 accumulation; the tensor-core variants of
// the GEMM-shaped stages live in gemm_tc.cuh and are selected by be_gemm when the shape qualifies.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cooperative_groups.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "backend.h"
#include "tc_rowgemm.cuh"
#include "tc_kpipe.cuh"
#include "tc_mid.cuh"
#include "tc_cmm.cuh"
#include "tc_cmm4.cuh"
#include "tc_conv.cuh"
#include "tc_wgrad.cuh"

namespace uno {

namespace {

inline cudaStream_t S(stream_t s) { return (cudaStream_t)s; }

// ---- per-launch profiling (off by default) -----------------------------------------------------------
struct ProfRec { const char* tag; cudaEvent_t e0, e1; double bytes, flops; int scope; };
struct ProfScopeInfo { std::string label; long calls = 0; double bytes = 0, flops = 0; };
std::atomic<bool> g_prof_on{false};
std::atomic<long> g_launches{0};   // forward and autograd's backward thread both launch
std::vector<ProfRec> g_prof;
std::vector<ProfScopeInfo> g_prof_scopes;
thread_local int t_prof_scope = -1;
std::mutex g_prof_mu;

struct ProfScope {
    bool on;
    cudaStream_t st;
    size_t idx = 0;
    ProfScope(const char* tag, double bytes, double flops, cudaStream_t s) : on(g_prof_on), st(s) {
        ++g_launches;
        if (!on) return;
        ProfRec r{tag, nullptr, nullptr, bytes, flops, t_prof_scope};
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, st);
        std::lock_guard<std::mutex> lk(g_prof_mu);
        idx = g_prof.size();
        g_prof.push_back(r);
    }
    ~ProfScope() {
        if (!on) return;
        std::lock_guard<std::mutex> lk(g_prof_mu);
        cudaEventRecord(g_prof[idx].e1, st);
    }
};

#define CU_LAUNCH_CHECK()                          \
    do {                                           \
        cudaError_t _e = cudaGetLastError();       \
        if (_e != cudaSuccess) return (int)_e;     \
    } while (0)

// Host -> device copy of a constant table (plan matrices, operand images), complete on return for EVERY stream.  A plain
// cudaMemcpy from pageable memory returns once the data sits in the driver's staging buffer; the DMA that follows is ordered
// only against blocking streams, and the library's side streams (and a caller's capture stream) are non-blocking: the first
// kernel that read a freshly built image on such a stream could see it half written (found on B200 as an intermittent wrong
// gradient in the first backward of a new shape when its kernels were already resident).  Tables are built once per shape.
static cudaError_t upload_sync(void* dst, const void* src, size_t bytes) {
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    return cudaDeviceSynchronize();
}

// Exact-erf GELU and its derivative from ONE exponential: with E = exp(-x^2/2) and
// t = 1/(1 + p|x|/sqrt2), 1 - erf(|x|/sqrt2) = E*t*poly(t) (Abramowitz-Stegun 7.1.26, |eps| <= 1.5e-7); measured in
// fp32 against fp64: |d gelu| <= 4.3e-7, |d gelu'| <= 3.2e-7, two orders below the parity tolerance, for ~16
// instructions instead of ~60 (erff + expf).  -DUNO_GELU_LIBM restores the libm forms.
__device__ __forceinline__ void gelu_both(float x, float& act, float& grad) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
    const float E = __expf(-0.5f * x * x);
    float poly = fmaf(t, 1.061405429f, -1.453152027f);
    poly = fmaf(t, poly, 1.421413741f);
    poly = fmaf(t, poly, -0.284496736f);
    poly = fmaf(t, poly, 0.254829592f);
    const float h = 0.5f * poly * t * E;
    const float cdf = x < 0.f ? h : 1.0f - h;
    act = x * cdf;
    grad = fmaf(x * 0.39894228040143267794f, E, cdf);
}
#ifdef UNO_GELU_LIBM
__device__ __forceinline__ float gelu_f(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
    return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) +
           x * 0.39894228040143267794f * expf(-0.5f * x * x);
}
#else
__device__ __forceinline__ float gelu_f(float x) { return tc::gelu_fwd_fast(x); }   // forward only: one ex2, no division
__device__ __forceinline__ float gelu_grad_f(float x) {
    float a, g;
    gelu_both(x, a, g);
    return g;
}
#endif

// =====================================================================================================
// GEMM: C[M,N] = A[M,K] * B[K,N], arbitrary element strides on A and B, row-major C.
// 256 threads arranged TX (n) x 256/TX (m); thread tile TM x TN with columns interleaved by TX so
// that a warp's store / B-fragment read covers consecutive columns.
// =====================================================================================================
struct GemmK {
    const float* A; long a_rs, a_cs, sA;
    const float* B; long b_rs, b_cs, sB;
    float* C; long ldc, sC;
    float* C2;
    const float* bias;
    int M, N, K;
    int ksplit;      // number of K partitions (atomic epilogue when > 1 or epi == EPI_ATOMIC)
    int kchunk;      // K elements per partition
    int epi;
};
enum { EPI_ATOMIC = 100 };

template <int BM, int BN, int BK, int TX>
__global__ void __launch_bounds__(256) gemm_kernel(const GemmK a) {
    constexpr int TY = 256 / TX;
    constexpr int TM = BM / TY;
    constexpr int TN = BN / TX;
    static_assert(TM >= 1 && TN >= 1, "tile");
    constexpr int LDA_S = BM + 4;
    __shared__ __align__(16) float As[2][BK][LDA_S];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int zb = blockIdx.z / a.ksplit;          // batch index
    const int zk = blockIdx.z - zb * a.ksplit;     // K partition
    const int ntiles = (a.N + BN - 1) / BN;       // tiles linearised on grid.x, n fastest
    const long m0 = (long)(blockIdx.x / ntiles) * BM;
    const int n0 = (int)(blockIdx.x % ntiles) * BN;
    const float* __restrict__ A = a.A + zb * a.sA;
    const float* __restrict__ B = a.B + zb * a.sB;
    const int k_begin = zk * a.kchunk;
    const int k_end = min(a.K, k_begin + a.kchunk);

    constexpr int A_PER = (BM * BK + 255) / 256;
    constexpr int B_PER = (BK * BN + 255) / 256;
    float ra[A_PER], rb[B_PER];
    const bool a_kfast = (a.a_cs == 1);
    const bool b_nfast = (a.b_cs == 1);

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int idx = tid + i * 256;
            int m, k;
            if (a_kfast) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
            float v = 0.f;
            if (idx < BM * BK && m0 + m < a.M && k0 + k < k_end)
                v = __ldg(A + (m0 + m) * a.a_rs + (long)(k0 + k) * a.a_cs);
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            const int idx = tid + i * 256;
            int n, k;
            if (b_nfast) { n = idx % BN; k = idx / BN; } else { k = idx % BK; n = idx / BK; }
            float v = 0.f;
            if (idx < BK * BN && n0 + n < a.N && k0 + k < k_end)
                v = __ldg(B + (long)(k0 + k) * a.b_rs + (long)(n0 + n) * a.b_cs);
            rb[i] = v;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int idx = tid + i * 256;
            int m, k;
            if (a_kfast) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
            if (idx < BM * BK) As[buf][k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            const int idx = tid + i * 256;
            int n, k;
            if (b_nfast) { n = idx % BN; k = idx / BN; } else { k = idx % BK; n = idx / BK; }
            if (idx < BK * BN) Bs[buf][k][n] = rb[i];
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    int buf = 0;
    load_tiles(k_begin);
    store_tiles(0);
    __syncthreads();
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool more = k0 + BK < k_end;
        if (more) load_tiles(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[TM], bv[TN];
            if constexpr (TM % 4 == 0) {
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
                    av[i] = t.x; av[i + 1] = t.y; av[i + 2] = t.z; av[i + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < TM; ++i) av[i] = As[buf][k][ty * TM + i];
            }
#pragma unroll
            for (int j = 0; j < TN; ++j) bv[j] = Bs[buf][k][tx + j * TX];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            store_tiles(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    float* __restrict__ C = a.C + zb * a.sC;
    float* __restrict__ C2 = a.C2 ? a.C2 + zb * a.sC : nullptr;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long m = m0 + ty * TM + i;
        if (m >= a.M) continue;
        const float bm = a.bias ? __ldg(a.bias + m) : 0.f;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx + j * TX;
            if (n >= a.N) continue;
            const long off = m * a.ldc + n;
            const float v = acc[i][j];
            switch (a.epi) {
                case EPI_STORE: C[off] = v + bm; break;
                case EPI_ACCUM: C[off] += v; break;
                case EPI_ACCUM_GELU: { const float s = C[off] + v; C[off] = s; C2[off] = gelu_f(s); } break;
                case EPI_ACCUM_GELU_INPLACE: C[off] = gelu_f(C[off] + v); break;
                case EPI_ATOMIC: atomicAdd(C + off, v); break;
            }
        }
    }
}

template <int BM, int BN, int BK, int TX>
int launch_gemm(const GemmK& k, int batch, cudaStream_t st) {
    const long tiles = (long)((k.M + BM - 1) / BM) * ((k.N + BN - 1) / BN);
    dim3 grid((unsigned)tiles, 1, (unsigned)(batch * k.ksplit));
    gemm_kernel<BM, BN, BK, TX><<<grid, 256, 0, st>>>(k);
    CU_LAUNCH_CHECK();
    return 0;
}

int dispatch_gemm(const GemmK& k, int batch, cudaStream_t st) {
    if (k.M <= 0 || k.N <= 0 || batch <= 0) return 0;
    const bool smallK = (k.kchunk <= 48);
    if (k.N <= 56) {   // skinny output: narrow thread layout, BN tailored to N
        const int N = k.N;
#define UNO_NARROW(BN)                                                        \
    return smallK ? launch_gemm<128, BN, 8, 16>(k, batch, st) : launch_gemm<128, BN, 16, 16>(k, batch, st)
        if (N <= 16) { UNO_NARROW(16); }
        if (N <= 32) { UNO_NARROW(32); }
        UNO_NARROW(48);
#undef UNO_NARROW
    }
    if (k.M <= 32) return smallK ? launch_gemm<32, 64, 8, 32>(k, batch, st) : launch_gemm<32, 64, 16, 32>(k, batch, st);
    if (k.M <= 64 || (k.M <= 192 && k.M % 128 != 0 && k.M % 64 == 0))
        return smallK ? launch_gemm<64, 64, 8, 32>(k, batch, st) : launch_gemm<64, 64, 16, 32>(k, batch, st);
    return smallK ? launch_gemm<128, 64, 8, 32>(k, batch, st) : launch_gemm<128, 64, 16, 32>(k, batch, st);
}

// =====================================================================================================
// complex transform along a leading axis, batched over planes:  Y[o,j,i] = sum_h Mat[j,h] * X[o,h,i]
// Register-tiled: a thread owns TJ x TI complex outputs of one plane; a CTA holds ppc planes x (nj x ni) threads
// and walks h in chunks staged in shared memory (the Mat chunk is shared by the CTA's planes).  Per h a thread
// issues TJ 8-byte + one (TI=2: 16-byte) shared loads for 4*TJ*TI FMAs.
// =====================================================================================================
struct Mid2K {
    const float2* X; const float2* Mat; float2* Y;
    int O, H, J, I;
    int nj, ni, ppc, HK, tilesJ, tilesI;
};

template <int TJ, int TI>
__global__ void __launch_bounds__(256) mid2_kernel(const Mid2K k) {
    extern __shared__ __align__(16) float2 msm[];
    const int JT = k.nj * TJ, IT = k.ni * TI, JTP = JT | 1;
    // shared memory: 2 x { Mat chunk [HK][JTP] (padded to an even count) ; X chunk [ppc][HK][IT] (16-byte aligned) }
    const int tid = threadIdx.x;
    const int tpp = k.nj * k.ni;
    const int grp = tid / tpp, t = tid - grp * tpp;
    const int tj = t / k.ni, ti = t - tj * k.ni;
    const bool active = grp < k.ppc;
    long bid = blockIdx.x;
    const int it = (int)(bid % k.tilesI); bid /= k.tilesI;
    const int jt = (int)(bid % k.tilesJ); bid /= k.tilesJ;
    const long o0 = bid * k.ppc;
    const int j0 = jt * JT, i0 = it * IT;
    float2 acc[TJ][TI];
#pragma unroll
    for (int a = 0; a < TJ; ++a)
#pragma unroll
        for (int b = 0; b < TI; ++b) acc[a][b] = make_float2(0.f, 0.f);
    // chunks of HK rows of h are staged with 8-byte LDGSTS into a double buffer: the copy of chunk c+1 overlaps the
    // arithmetic of chunk c (out-of-range elements are zero-filled through src-size 0)
    const size_t ms_elems = (size_t)k.HK * JTP + (((size_t)k.HK * JTP) & 1);
    const size_t buf_elems = ms_elems + (size_t)k.ppc * k.HK * IT;
    // Staging cursors without divisions: element idx = tid + 256*n of a chunk maps to (j, h) resp. (plane g, h, i); the
    // mapping is the same for every chunk, so the first one is divided out once and the others follow by carry
    // propagation (the index arithmetic used to cost a quarter of the kernel's instructions).
    const int m_h0 = tid % k.HK, m_j0 = tid / k.HK, m_dh = 256 % k.HK, m_dj = 256 / k.HK;
    const int x_i0 = tid % IT, x_r0 = tid / IT, x_di = 256 % IT, x_dr = 256 / IT;
    auto stage = [&](int h0, int buf) {
        float2* Mb = msm + (size_t)buf * buf_elems;
        float2* Xb = Mb + ms_elems;
        const int hk = min(k.HK, k.H - h0);
        {
            int h = m_h0, j = m_j0;
            for (int idx = tid; idx < JT * k.HK; idx += 256) {
                const bool on = j0 + j < k.J && h < hk;
                const float2* src = on ? k.Mat + (long)(j0 + j) * k.H + h0 + h : k.Mat;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(Mb + h * JTP + j);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(on ? 8 : 0) : "memory");
                h += m_dh; j += m_dj;
                if (h >= k.HK) { h -= k.HK; ++j; }
            }
        }
        {
            int i = x_i0, r = x_r0;                      // r = g * HK + h
            int h = r % k.HK, g = r / k.HK;
            const int x_dh = x_dr % k.HK, x_dg = x_dr / k.HK;
            for (int idx = tid; idx < k.ppc * k.HK * IT; idx += 256) {
                const bool on = o0 + g < k.O && h < hk && i0 + i < k.I;
                const float2* src = on ? k.X + ((o0 + g) * k.H + h0 + h) * (long)k.I + i0 + i : k.X;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(Xb + idx);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(on ? 8 : 0) : "memory");
                i += x_di; h += x_dh; g += x_dg;
                if (i >= IT) { i -= IT; ++h; }
                if (h >= k.HK) { h -= k.HK; ++g; }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0, 0);
    int cur = 0;
    for (int h0 = 0; h0 < k.H; h0 += k.HK, cur ^= 1) {
        const int hk = min(k.HK, k.H - h0);
        if (h0 + k.HK < k.H) {
            stage(h0 + k.HK, cur ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        if (active) {
            const float2* mp = msm + (size_t)cur * buf_elems + tj * TJ;
            const float2* xp = msm + (size_t)cur * buf_elems + ms_elems + (size_t)grp * k.HK * IT + ti * TI;
#pragma unroll 4
            for (int h = 0; h < hk; ++h) {
                float2 m[TJ], x[TI];
#pragma unroll
                for (int a = 0; a < TJ; ++a) m[a] = mp[h * JTP + a];
                if (TI == 2) {
                    const float4 v = *reinterpret_cast<const float4*>(xp + (size_t)h * IT);
                    x[0] = make_float2(v.x, v.y);
                    x[TI - 1] = make_float2(v.z, v.w);
                } else {
                    x[0] = xp[(size_t)h * IT];
                }
#pragma unroll
                for (int a = 0; a < TJ; ++a)
#pragma unroll
                    for (int b = 0; b < TI; ++b) {
                        acc[a][b].x = fmaf(m[a].x, x[b].x, acc[a][b].x); acc[a][b].x = fmaf(-m[a].y, x[b].y, acc[a][b].x);
                        acc[a][b].y = fmaf(m[a].x, x[b].y, acc[a][b].y); acc[a][b].y = fmaf(m[a].y, x[b].x, acc[a][b].y);
                    }
            }
        }
        __syncthreads();
    }
    if (active && o0 + grp < k.O) {
        float2* Yo = k.Y + (o0 + grp) * (long)k.J * k.I;
#pragma unroll
        for (int a = 0; a < TJ; ++a) {
            const int j = j0 + tj * TJ + a;
            if (j >= k.J) continue;
#pragma unroll
            for (int b = 0; b < TI; ++b) {
                const int i = i0 + ti * TI + b;
                if (i < k.I) Yo[(long)j * k.I + i] = acc[a][b];
            }
        }
    }
}

// =====================================================================================================
// per-mode complex contraction, mode index on the lanes: C[m,n,q] = sum_k A[m,k,q] B[k,n,q]
// block (32 q, 4 n-tiles); thread tile 4 m x 4 n.
// =====================================================================================================
__global__ void __launch_bounds__(128, 6) cmm_kernel(const CmmArgs a, int chunks_per_row) {
    const int per_corner = chunks_per_row * a.q_outer;
    const int corner = blockIdx.x / per_corner;
    const int bx = blockIdx.x - corner * per_corner;
    const int qo = bx / chunks_per_row;
    const int qi = (bx - qo * chunks_per_row) * 32 + threadIdx.x;
    const int m0 = blockIdx.y * 4;
    const int n0 = (blockIdx.z * 4 + threadIdx.y) * 4;
    if (qi >= a.q_inner || n0 >= a.N) return;
    const float2* A = reinterpret_cast<const float2*>(a.A[corner]) + (long)qo * a.a_sqo + qi;
    const float2* B = reinterpret_cast<const float2*>(a.B[corner]) + (long)qo * a.b_sqo + qi;
    float2* C = reinterpret_cast<float2*>(a.C[corner]) + (long)qo * a.c_sqo + qi;
    const float sa = a.conjA ? -1.f : 1.f, sb = a.conjB ? -1.f : 1.f;
    float2 acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
    int mi[4], nj[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { mi[i] = min(m0 + i, a.M - 1); nj[i] = min(n0 + i, a.N - 1); }
#pragma unroll 4
    for (int k = 0; k < a.K; ++k) {
        float2 av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            av[i] = __ldg(A + (long)mi[i] * a.a_sm + (long)k * a.a_sk);
            av[i].y *= sa;
            bv[i] = __ldg(B + (long)k * a.b_sk + (long)nj[i] * a.b_sn);
            bv[i].y *= sb;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][j].x = fmaf(av[i].x, bv[j].x, acc[i][j].x);
                acc[i][j].x = fmaf(-av[i].y, bv[j].y, acc[i][j].x);
                acc[i][j].y = fmaf(av[i].x, bv[j].y, acc[i][j].y);
                acc[i][j].y = fmaf(av[i].y, bv[j].x, acc[i][j].y);
            }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (m0 + i >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (n0 + j < a.N) C[(long)(m0 + i) * a.c_sm + (long)(n0 + j) * a.c_sn] = acc[i][j];
    }
}

// =====================================================================================================
// per-mode complex contraction, tiled: the regime of many channels (NS-2D inner levels: 64 x 192 x 192 per mode, 36 modes).
// A CTA owns 4 consecutive modes x a 32 (m) x 32 (n) output tile; a thread one mode and a 4 x 4 complex register tile.
// k runs in chunks of 16 staged with 8-byte LDGSTS into a double buffer (lanes run over the 4 modes first, so every
// 32-byte sector fetched from L2 is used whole); per k a thread issues 4 16-byte shared loads for 64 FMAs.
// =====================================================================================================
constexpr int kC2Q = 4, kC2M = 32, kC2N = 32, kC2K = 16;
constexpr int kC2QStrideA = kC2K * kC2M + 8, kC2QStrideB = kC2K * kC2N + 8;   // float2 elements; +8 keeps the 4 modes' stores apart
constexpr int kC2BufElems = kC2Q * (kC2QStrideA + kC2QStrideB);

__global__ void __launch_bounds__(256) cmm2_kernel(const CmmArgs a, int qchunks) {
    extern __shared__ __align__(16) float2 csm[];
    const int tid = threadIdx.x;
    const int per_corner = qchunks * a.q_outer;
    const int corner = blockIdx.x / per_corner;
    const int bx = blockIdx.x - corner * per_corner;
    const int qo = bx / qchunks;
    const int q0 = (bx - qo * qchunks) * kC2Q;
    const int m0 = blockIdx.y * kC2M, n0 = blockIdx.z * kC2N;
    const float2* A = reinterpret_cast<const float2*>(a.A[corner]) + (long)qo * a.a_sqo;
    const float2* B = reinterpret_cast<const float2*>(a.B[corner]) + (long)qo * a.b_sqo;
    float2* C = reinterpret_cast<float2*>(a.C[corner]) + (long)qo * a.c_sqo;
    // staging map: lane bits 0-1 = mode, the rest = (row, k) pair index
    const int sq = tid & 3, sp = tid >> 2;
    const bool q_ok = q0 + sq < a.q_inner;
    auto stage = [&](int k0, int buf) {
        float2* As = csm + (size_t)buf * kC2BufElems + sq * kC2QStrideA;
        float2* Bs = csm + (size_t)buf * kC2BufElems + kC2Q * kC2QStrideA + sq * kC2QStrideB;
#pragma unroll
        for (int it = 0; it < kC2K * kC2M / 64; ++it) {
            const int pr = sp + 64 * it;
            const int m = pr % kC2M, kk = pr / kC2M;
            const bool on = q_ok && m0 + m < a.M && k0 + kk < a.K;
            const float2* src = on ? A + (long)(m0 + m) * a.a_sm + (long)(k0 + kk) * a.a_sk + q0 + sq : A;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(As + kk * kC2M + m);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(on ? 8 : 0) : "memory");
        }
#pragma unroll
        for (int it = 0; it < kC2K * kC2N / 64; ++it) {
            const int pr = sp + 64 * it;
            const int n = pr % kC2N, kk = pr / kC2N;
            const bool on = q_ok && n0 + n < a.N && k0 + kk < a.K;
            const float2* src = on ? B + (long)(k0 + kk) * a.b_sk + (long)(n0 + n) * a.b_sn + q0 + sq : B;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(Bs + kk * kC2N + n);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(on ? 8 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // compute map: 64 threads per mode, 8 x 8 of them over the tile
    const int cq = tid >> 6, ty = (tid >> 3) & 7, tx = tid & 7;
    const float sa = a.conjA ? -1.f : 1.f, sb = a.conjB ? -1.f : 1.f;
    float2 acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
    stage(0, 0);
    int cur = 0;
    for (int k0 = 0; k0 < a.K; k0 += kC2K, cur ^= 1) {
        if (k0 + kC2K < a.K) {
            stage(k0 + kC2K, cur ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const float2* As = csm + (size_t)cur * kC2BufElems + cq * kC2QStrideA + ty * 4;
        const float2* Bs = csm + (size_t)cur * kC2BufElems + kC2Q * kC2QStrideA + cq * kC2QStrideB + tx * 4;
        const int kn = min(kC2K, a.K - k0);
#pragma unroll 4
        for (int kk = 0; kk < kn; ++kk) {
            float2 av[4], bv[4];
            const float4 a01 = *reinterpret_cast<const float4*>(As + kk * kC2M), a23 = *reinterpret_cast<const float4*>(As + kk * kC2M + 2);
            const float4 b01 = *reinterpret_cast<const float4*>(Bs + kk * kC2N), b23 = *reinterpret_cast<const float4*>(Bs + kk * kC2N + 2);
            av[0] = make_float2(a01.x, a01.y * sa); av[1] = make_float2(a01.z, a01.w * sa);
            av[2] = make_float2(a23.x, a23.y * sa); av[3] = make_float2(a23.z, a23.w * sa);
            bv[0] = make_float2(b01.x, b01.y * sb); bv[1] = make_float2(b01.z, b01.w * sb);
            bv[2] = make_float2(b23.x, b23.y * sb); bv[3] = make_float2(b23.z, b23.w * sb);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j].x = fmaf(av[i].x, bv[j].x, acc[i][j].x);
                    acc[i][j].x = fmaf(-av[i].y, bv[j].y, acc[i][j].x);
                    acc[i][j].y = fmaf(av[i].x, bv[j].y, acc[i][j].y);
                    acc[i][j].y = fmaf(av[i].y, bv[j].x, acc[i][j].y);
                }
        }
        __syncthreads();
    }
    if (q0 + cq < a.q_inner) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty * 4 + i;
            if (m >= a.M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx * 4 + j;
                if (n < a.N) C[(long)m * a.c_sm + (long)n * a.c_sn + q0 + cq] = acc[i][j];
            }
        }
    }
}

// =====================================================================================================
// banded resample
// =====================================================================================================
__global__ void __launch_bounds__(256) banded_kernel(const BandedArgs a, long total) {
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % a.inner);
        long r = idx / a.inner;
        const int j = (int)(r % a.n_out);
        const long o = r / a.n_out;
        const float* w = a.w + (long)j * a.taps;
        const float* x = a.x + (o * a.n_in + __ldg(a.start + j)) * a.inner + i;
        float acc = 0.f;
        for (int t = 0; t < a.taps; ++t) acc = fmaf(__ldg(w + t), __ldg(x + (long)t * a.inner), acc);
        a.y[idx] = acc;
    }
}


// fused separable 2-D resample: one CTA = one plane x (32 x 64) output tile.  The input window is staged
// in shared memory, the column (last-axis) bands are applied into a second shared buffer, then the row
// bands; x is read once and y written once.
constexpr int kB2TH = 32, kB2TW = 64, kB2MidLd = 80;   // mid row pitch 80: rows r and r+1 of a warp land in disjoint banks
__global__ void __launch_bounds__(256) banded2d_kernel(const Banded2DArgs a, int tiles_h, int tiles_w, int RIN, int ldin) {
    extern __shared__ float sm[];
    float* in_s = sm;                                  // [RIN][ldin], ldin odd
    float* mid_s = in_s + (size_t)RIN * ldin;          // [RIN][kB2MidLd]
    float* w0s = mid_s + (size_t)RIN * kB2MidLd;       // [TH][taps0]
    float* w1t = w0s + kB2TH * a.taps0;                // [taps1][TW]  (transposed: conflict-free across columns)
    int* st0s = reinterpret_cast<int*>(w1t + kB2TW * a.taps1);
    int* st1s = st0s + kB2TH;
    long bid = blockIdx.x;
    const int tw = (int)(bid % tiles_w); bid /= tiles_w;
    const int th = (int)(bid % tiles_h); bid /= tiles_h;
    const long p = bid;
    const int tid = threadIdx.x;
    const int i0 = th * kB2TH, j0 = tw * kB2TW;
    const int nh = min(kB2TH, a.n_out0 - i0), nw = min(kB2TW, a.n_out1 - j0);
    const int r0 = __ldg(a.start0 + i0), c0 = __ldg(a.start1 + j0);
    const int rin = min(__ldg(a.start0 + i0 + nh - 1) + a.taps0, a.n_in0) - r0;
    const int cin = min(__ldg(a.start1 + j0 + nw - 1) + a.taps1, a.n_in1) - c0;
    const float* xp = a.x + (p * a.n_in0 + r0) * (long)a.n_in1 + c0;
    const int tx = tid & 63, ty = tid >> 6;
    // asynchronous 4-byte copies (LDGSTS): the whole input window is in flight at once
    for (int r = ty; r < rin; r += 4) {
        const float* src = xp + (long)r * a.n_in1;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(in_s + r * ldin);
        for (int c = tx; c < cin; c += 64)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * c), "l"(src + c) : "memory");
    }
    for (int i = tid; i < nh * a.taps0; i += 256) w0s[i] = __ldg(a.w0 + (long)i0 * a.taps0 + i);
    if (tx < nw)
        for (int t = ty; t < a.taps1; t += 4) w1t[t * kB2TW + tx] = __ldg(a.w1 + (long)(j0 + tx) * a.taps1 + t);
    for (int i = tid; i < nh; i += 256) st0s[i] = __ldg(a.start0 + i0 + i) - r0;
    for (int i = tid; i < nw; i += 256) st1s[i] = __ldg(a.start1 + j0 + i) - c0;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    {   // column bands: a warp covers 2 rows x 16 columns, so stride-2 (down-sampling) reads are conflict-free
        const int hx = tid & 15, hy = tid >> 4;
        for (int r = hy; r < rin; r += 16) {
            const float* row = in_s + r * ldin;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = hx + 16 * jj;
                if (j < nw) {
                    const float* src = row + st1s[j];
                    const float* w = w1t + j;
                    float acc = 0.f;
#pragma unroll 4
                    for (int t = 0; t < a.taps1; ++t) acc = fmaf(w[t * kB2TW], src[t], acc);
                    mid_s[r * kB2MidLd + j] = acc;
                }
            }
        }
    }
    __syncthreads();
    if (tx < nw) {
        float* yp = a.y + (p * a.n_out0 + i0) * (long)a.n_out1 + j0 + tx;
        for (int i = ty; i < nh; i += 4) {
            const float* w = w0s + i * a.taps0;
            const float* src = mid_s + st0s[i] * kB2MidLd + tx;
            float acc = 0.f;
#pragma unroll 4
            for (int t = 0; t < a.taps0; ++t) acc = fmaf(w[t], src[t * kB2MidLd], acc);
            yp[(long)i * a.n_out1] = acc;
        }
    }
}

// =====================================================================================================
// elementwise and per-plane reductions
// =====================================================================================================
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const float* __restrict__ pre, float* __restrict__ y, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = gelu_f(pre[i]);
}
__device__ __forceinline__ double block_sum(double v, double* sh);

// The (up to two) strided sources of an upstream gradient as the kernels see them (backend.h UpGrad): plane p = (b, c) of
// source i starts at p_i + b * bs_i + c * L.
struct UpGradK {
    const float* p0; long bs0; const float* p1; long bs1;
};
__device__ __forceinline__ const float* upgrad_plane(const float* base, long bs, long p, int C, long L) {
    const long b = p / C;
    return base + b * bs + (p - b * C) * L;
}

// plane-structured activation backward with the conv-bias gradient folded in: grid (planes, chunks).
//   g = (gy0 + gy1) * gelu'(pre)      ACT = false: g = gy0 + gy1 (a block without activation: only sums / gathers its sources)
// Planes of an odd grid (481 x 481) start at any 4-byte offset, so each plane is cut into a scalar head (up to the next
// 16-byte boundary), a float4 body and a scalar tail; all tensors share the cut when `vec_ok` (host-checked: bases 16-byte
// aligned and either every tensor contiguous -- plane p starts at p*L everywhere -- or L and the batch strides multiples of 4).
template <bool ACT, bool TWO>
__global__ void __launch_bounds__(256) gelu_bwd_bias_kernel(const UpGradK gy, const float* __restrict__ pre,
                                                            float* __restrict__ g, int C, long L, float* __restrict__ gbias,
                                                            float alpha, int vec_ok) {
    __shared__ double sh[32];
    const long p = blockIdx.x;
    const float* gp = upgrad_plane(gy.p0, gy.bs0, p, C, L);
    const float* hp = TWO ? upgrad_plane(gy.p1, gy.bs1, p, C, L) : nullptr;
    const float* pp = ACT ? pre + p * L : nullptr;
    float* op = g + p * L;
    const long t = (long)blockIdx.y * blockDim.x + threadIdx.x, stride = (long)gridDim.y * blockDim.x;
    const long head = vec_ok ? min(L, (4 - ((p * L) & 3)) & 3) : L;
    const long nv = (L - head) >> 2;
    float s = 0.f;
    {
        const float4* g4 = reinterpret_cast<const float4*>(gp + head);
        const float4* h4 = reinterpret_cast<const float4*>(TWO ? hp + head : gp);
        const float4* p4 = reinterpret_cast<const float4*>(ACT ? pp + head : gp);
        float4* o4 = reinterpret_cast<float4*>(op + head);
#pragma unroll 2
        for (long j = t; j < nv; j += stride) {
            float4 a = __ldcs(g4 + j);
            if (TWO) {
                const float4 c = __ldcs(h4 + j);
                a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
            }
            float4 v = a;
            if (ACT) {
                const float4 b = __ldcs(p4 + j);
                v.x = a.x * gelu_grad_f(b.x);
                v.y = a.y * gelu_grad_f(b.y);
                v.z = a.z * gelu_grad_f(b.z);
                v.w = a.w * gelu_grad_f(b.w);
            }
            o4[j] = v;
            s += (v.x + v.y) + (v.z + v.w);
        }
    }
    // scalar remainder: [0, head) and [head + 4*nv, L)
    const long rem = L - 4 * nv;
    for (long r = t; r < rem; r += stride) {
        const long i = r < head ? r : r + 4 * nv;
        float u = gp[i];
        if (TWO) u += hp[i];
        const float v = ACT ? u * gelu_grad_f(pp[i]) : u;
        op[i] = v;
        s += v;
    }
    if (gbias) {
        const double tot = block_sum((double)s, sh);
        if (threadIdx.x == 0) atomicAdd(gbias + (p % C), alpha * (float)tot);
    }
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    if (w == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (l == 0) sh[0] = t;
    }
    __syncthreads();
    return sh[0];
}

__global__ void __launch_bounds__(512) plane_stats_kernel(const float* __restrict__ x, float* __restrict__ stats, long L, float eps) {
    __shared__ double sh[32];
    const float* p = x + (long)blockIdx.x * L;
    float s = 0.f;
    double sd = 0.0;
    int cnt = 0;
    for (long i = threadIdx.x; i < L; i += blockDim.x) {
        s += p[i];
        if (++cnt == 64) { sd += s; s = 0.f; cnt = 0; }
    }
    sd += s;
    const double mu = block_sum(sd, sh) / (double)L;
    const float muf = (float)mu;
    s = 0.f; sd = 0.0; cnt = 0;
    for (long i = threadIdx.x; i < L; i += blockDim.x) {
        const float d = p[i] - muf;
        s = fmaf(d, d, s);
        if (++cnt == 64) { sd += s; s = 0.f; cnt = 0; }
    }
    sd += s;
    const double var = block_sum(sd, sh) / (double)L;
    if (threadIdx.x == 0) {
        stats[2 * blockIdx.x] = muf;
        stats[2 * blockIdx.x + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

__global__ void __launch_bounds__(256) norm_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float* __restrict__ y, int C, long L, int non_lin) {
    const long p = blockIdx.x;
    const int c = (int)(p % C);
    const float mu = stats[2 * p], rstd = stats[2 * p + 1];
    const float g = gamma[c] * rstd, b = beta[c] - mu * rstd * gamma[c];
    const float* xp = x + p * L;
    float* yp = y + p * L;
    for (long i = (long)blockIdx.y * blockDim.x + threadIdx.x; i < L; i += (long)gridDim.y * blockDim.x) {
        const float n = fmaf(xp[i], g, b);
        yp[i] = non_lin ? gelu_f(n) : n;
    }
}

__global__ void __launch_bounds__(512) norm_act_bwd_kernel(const UpGradK gy, const float* __restrict__ x,
                                                           const float* __restrict__ stats, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float* __restrict__ g,
                                                           float* __restrict__ ggamma, float* __restrict__ gbeta, int C,
                                                           long L, int non_lin) {
    __shared__ double sh[32];
    const long p = blockIdx.x;
    const int c = (int)(p % C);
    const float mu = stats[2 * p], rstd = stats[2 * p + 1];
    const float ga = gamma[c], be = beta[c];
    const float* xp = x + p * L;
    const float* gp = upgrad_plane(gy.p0, gy.bs0, p, C, L);
    const float* hp = gy.p1 ? upgrad_plane(gy.p1, gy.bs1, p, C, L) : nullptr;
    float s1 = 0.f, s2 = 0.f;
    double d1 = 0.0, d2 = 0.0;
    int cnt = 0;
    for (long i = threadIdx.x; i < L; i += blockDim.x) {
        const float xh = (xp[i] - mu) * rstd;
        const float up = hp ? gp[i] + hp[i] : gp[i];
        const float gn = non_lin ? up * gelu_grad_f(fmaf(xh, ga, be)) : up;
        s1 += gn;
        s2 = fmaf(gn, xh, s2);
        if (++cnt == 64) { d1 += s1; d2 += s2; s1 = s2 = 0.f; cnt = 0; }
    }
    d1 += s1; d2 += s2;
    const double t1 = block_sum(d1, sh);
    const double t2 = block_sum(d2, sh);
    if (threadIdx.x == 0) {
        atomicAdd(ggamma + c, (float)t2);
        atomicAdd(gbeta + c, (float)t1);
    }
    const float m1 = (float)(t1 / (double)L), m2 = (float)(t2 / (double)L);
    const float k = ga * rstd;
    float* op = g + p * L;
    for (long i = threadIdx.x; i < L; i += blockDim.x) {
        const float xh = (xp[i] - mu) * rstd;
        const float up = hp ? gp[i] + hp[i] : gp[i];
        const float gn = non_lin ? up * gelu_grad_f(fmaf(xh, ga, be)) : up;
        op[i] = k * (gn - m1 - xh * m2);
    }
}

__global__ void __launch_bounds__(512) channel_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int C, long L, float alpha) {
    __shared__ double sh[32];
    const long p = blockIdx.x;
    const float* xp = x + p * L;
    float s = 0.f;
    double sd = 0.0;
    int cnt = 0;
    for (long i = threadIdx.x; i < L; i += blockDim.x) {
        s += xp[i];
        if (++cnt == 64) { sd += s; s = 0.f; cnt = 0; }
    }
    sd += s;
    const double t = block_sum(sd, sh);
    if (threadIdx.x == 0) atomicAdd(out + (p % C), alpha * (float)t);
}

__global__ void __launch_bounds__(256) add_channel_const_kernel(float* __restrict__ y, const float* __restrict__ v, float alpha, int C, long L) {
    const long p = blockIdx.x;
    const float c = v[p % C] * alpha;
    float* yp = y + p * L;
    for (long i = (long)blockIdx.y * blockDim.x + threadIdx.x; i < L; i += (long)gridDim.y * blockDim.x) yp[i] += c;
}

inline unsigned grid_for(size_t n, int block, size_t cap = 148 * 16) {
    size_t g = (n + block - 1) / block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

// Function attributes, SM counts and scratch buffers are PER DEVICE: a process may drive several GPUs (a model on cuda:1).
inline int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
}
struct DeviceOnce {   // "has this one-time setup run on the current device yet?"
    std::atomic<unsigned long long> mask{0};
    bool done() const { return (mask.load(std::memory_order_acquire) >> (current_device() & 63)) & 1ull; }
    void mark() { mask.fetch_or(1ull << (current_device() & 63), std::memory_order_release); }
};

template <typename K>
int ensure_smem(K kernel, size_t bytes) {
    // raise the dynamic shared-memory limit of this instantiation (idempotent, cheap)
    if (bytes <= 48 * 1024) return 0;
    return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

#include "pixel_mlp.cuh"
#include "pixel_mlp_tc.cuh"
#include "resample2d.cuh"
#include "norm_cluster.cuh"
#include "train_ops.cuh"


// =====================================================================================================
// tensor-core path: host-side operand images and dispatch
// =====================================================================================================
inline float tf32_rn(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0xFFFu + ((u >> 13) & 1u);
    u &= 0xFFFFE000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

struct TcImage {
    float* dev = nullptr;
    int n_tiles = 0, N_t = 0, K_pad = 0;
};
struct TcKey {
    const void* p; int K, N; long ldb; int variant = 0;
    bool operator<(const TcKey& o) const {
        if (p != o.p) return p < o.p;
        if (K != o.K) return K < o.K;
        if (N != o.N) return N < o.N;
        if (variant != o.variant) return variant < o.variant;
        return ldb < o.ldb;
    }
};
std::map<TcKey, TcImage> g_tc_images;
std::mutex g_tc_mu;

bool tc_enabled() { return cfg(CFG_TC) != 0; }

// B [K x N] (device, row-major, ldb) -> per n-tile [hi | lo] images in the UMMA K-major interleave layout
// of the transposed operand (N_t rows x K_pad): element (n, k) at float offset (k/4)*N_t*4 + n*4 + k%4.
// With `shifts` == 2 a second set of tiles follows in which tile column n holds B column t*N_t + n - 1 (the images
// the odd-row CTAs of the parity mode load, tc_rowgemm.cuh).
int tc_get_rowgemm_image(const float* B, long ldb, int K, int N, int shifts, TcImage* out) {
    std::lock_guard<std::mutex> lk(g_tc_mu);
    TcKey key{B, K, N, ldb, shifts};
    auto it = g_tc_images.find(key);
    if (it != g_tc_images.end()) { *out = it->second; return 0; }
    TcImage img;
    img.K_pad = ((K + 7) / 8) * 8;
    img.n_tiles = (N + 255) / 256;
    const int per = (N + img.n_tiles - 1) / img.n_tiles;
    img.N_t = ((per + 15) / 16) * 16;
    std::vector<float> hB((size_t)K * N);
    cudaError_t e = cudaMemcpy2D(hB.data(), (size_t)N * 4, B, (size_t)ldb * 4, (size_t)N * 4, K, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return (int)e;
    const size_t half = (size_t)img.K_pad * img.N_t;
    std::vector<float> h((size_t)shifts * img.n_tiles * 2 * half, 0.0f);
    for (int sh = 0; sh < shifts; ++sh)
        for (int t = 0; t < img.n_tiles; ++t)
            for (int n = 0; n < img.N_t; ++n) {
                const int gn = t * img.N_t + n - sh;
                if (gn < 0 || gn >= N) continue;
                const size_t base = (size_t)(sh * img.n_tiles + t) * 2 * half;
                for (int k = 0; k < K; ++k) {
                    const float b = hB[(size_t)k * N + gn];
                    const float hi = tf32_rn(b);
                    const float lo = tf32_rn(b - hi);
                    const size_t o = (size_t)(k / 4) * img.N_t * 4 + (size_t)n * 4 + (k % 4);
                    h[base + o] = hi;
                    h[base + half + o] = lo;
                }
            }
    e = cudaMalloc(&img.dev, h.size() * 4);
    if (e != cudaSuccess) return (int)e;
    e = upload_sync(img.dev, h.data(), h.size() * 4);
    if (e != cudaSuccess) return (int)e;
    g_tc_images[key] = img;
    *out = img;
    return 0;
}

struct TcKpImage { float* dev = nullptr; int N_t = 0, n_chunks = 0; };
std::map<TcKey, TcKpImage> g_tc_kp_images;

struct TcMidImage { float* dev = nullptr; int N_t = 0, n_tiles = 0, n_chunks = 0; };
std::map<TcKey, TcMidImage> g_tc_mid_images;

void tc_forget(const void* p) {
    std::lock_guard<std::mutex> lk(g_tc_mu);
    for (auto it = g_tc_images.begin(); it != g_tc_images.end();) {
        if (it->first.p == p) { cudaFree(it->second.dev); it = g_tc_images.erase(it); } else ++it;
    }
    for (auto it = g_tc_kp_images.begin(); it != g_tc_kp_images.end();) {
        if (it->first.p == p) { cudaFree(it->second.dev); it = g_tc_kp_images.erase(it); } else ++it;
    }
    for (auto it = g_tc_mid_images.begin(); it != g_tc_mid_images.end();) {
        if (it->first.p == p) { cudaFree(it->second.dev); it = g_tc_mid_images.erase(it); } else ++it;
    }
}

int g_num_sms[64] = {0};
int num_sms() {
    const int dev = current_device() & 63;
    if (!g_num_sms[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        g_num_sms[dev] = n > 0 ? n : 148;
    }
    return g_num_sms[dev];
}


template <int EPI, int G, int J>
int launch_rowgemm_variant(const tc::RowGemmParams& p, size_t smem, cudaStream_t st) {
    static DeviceOnce configured;
    if (!configured.done()) {
        cudaError_t e = cudaFuncSetAttribute(tc::rowgemm_smallk_kernel<EPI, G, J>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured.mark();
    }
    int gx = num_sms() / p.n_tiles;
    gx &= ~(p.nclass - 1);                  // a CTA must only ever see tiles of one row class (m_tiles is a multiple of nclass too)
    if (gx < p.nclass) gx = p.nclass;
    if ((long)gx > p.m_tiles) gx = (int)p.m_tiles;
    tc::rowgemm_smallk_kernel<EPI, G, J><<<dim3(gx, p.n_tiles), tc::rowgemm_threads(G), smem, st>>>(p);
    CU_LAUNCH_CHECK();
    return 0;
}

template <int EPI>
int launch_rowgemm(const tc::RowGemmParams& p, size_t smem, cudaStream_t st) {
    // 16 epilogue warps of 32x16 elements per round (default: measured 3.66 -> 3.62 ms per Darcy step, 0.57 -> 0.51 ms per
    // NS-2D call) or 8 warps of 32x32; same arithmetic in the same order, bit-identical results
    if (cfg(CFG_ROWGEMM_EPI16)) return launch_rowgemm_variant<EPI, 4, 1>(p, smem, st);
    return launch_rowgemm_variant<EPI, 2, 2>(p, smem, st);
}

// chunk-major image for the K-pipelined kernel: [chunk][hi | lo], half = (KC/4) x N_t x 16 bytes,
// element (n, k) of chunk c at float offset ((k%32)/4)*N_t*4 + n*4 + k%4

// `lda_mod4` < 0: one image.  Otherwise the four row-class images of tc_kpipe.cuh follow each other ([class][chunk][hi | lo]),
// all with ceil((K + 3) / 32) chunks: class c holds B shifted down by sh = (c * lda) % 4 rows (rows k' < sh are zero).
int tc_get_kpipe_image(const float* B, long ldb, int K, int N, TcKpImage* out, int lda_mod4 = -1) {
    std::lock_guard<std::mutex> lk(g_tc_mu);
    TcKey key{B, K, N, ldb, lda_mod4 < 0 ? 0 : 4 + lda_mod4};
    auto it = g_tc_kp_images.find(key);
    if (it != g_tc_kp_images.end()) { *out = it->second; return 0; }
    TcKpImage img;
    const int classes = lda_mod4 < 0 ? 1 : 4;
    img.N_t = ((N + 15) / 16) * 16;
    img.n_chunks = ((lda_mod4 < 0 ? K : K + 3) + tc::kKC - 1) / tc::kKC;
    std::vector<float> hB((size_t)K * N);
    cudaError_t e = cudaMemcpy2D(hB.data(), (size_t)N * 4, B, (size_t)ldb * 4, (size_t)N * 4, K, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return (int)e;
    const size_t half = (size_t)img.N_t * tc::kKC;
    std::vector<float> h((size_t)classes * img.n_chunks * 2 * half, 0.0f);
    for (int cls = 0; cls < classes; ++cls) {
        const int sh = lda_mod4 < 0 ? 0 : (cls * lda_mod4) % 4;
        const size_t cbase = (size_t)cls * img.n_chunks * 2 * half;
        for (int k = 0; k < K; ++k) {
            const int c = (k + sh) / tc::kKC, kk = (k + sh) % tc::kKC;
            for (int n = 0; n < N; ++n) {
                const float b = hB[(size_t)k * N + n];
                const float hi = tf32_rn(b);
                const float lo = tf32_rn(b - hi);
                const size_t o = (size_t)(kk / 4) * img.N_t * 4 + (size_t)n * 4 + (kk % 4);
                h[cbase + (size_t)c * 2 * half + o] = hi;
                h[cbase + (size_t)c * 2 * half + half + o] = lo;
            }
        }
    }
    e = cudaMalloc(&img.dev, h.size() * 4);
    if (e != cudaSuccess) return (int)e;
    e = upload_sync(img.dev, h.data(), h.size() * 4);
    if (e != cudaSuccess) return (int)e;
    g_tc_kp_images[key] = img;
    *out = img;
    return 0;
}

template <int LW, bool RC, bool DBG = false>
int launch_kpipe(const tc::KPipeParams& p, int gx, size_t smem, cudaStream_t st) {
    static DeviceOnce configured;
    if (!configured.done()) {
        cudaError_t e = cudaFuncSetAttribute(tc::kpipe_kernel<LW, RC, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured.mark();
    }
    tc::kpipe_kernel<LW, RC, DBG><<<gx, (LW + tc::kKpEpiWarps + 2) * 32, smem, st>>>(p);
    CU_LAUNCH_CHECK();
    return 0;
}

int try_tc_kpipe(const GemmArgs& a, cudaStream_t st) {
    if (!tc_enabled() || !a.b_const || a.batch != 1 || a.a_cs != 1 || a.bias || a.epi != EPI_STORE || a.K <= 64 || a.N < 8 || a.N > 256 || a.M < 1)
        return -1;
    const int N_t = ((a.N + 15) / 16) * 16;
    int stages = (int)((200 * 1024) / tc::kpipe_stage_bytes(N_t));
    if (stages > 4) stages = 4;
    if (stages < 2) return -1;
    const bool a_vec_ok = (a.a_rs % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0);
    // row-class mode (tc_kpipe.cuh) for rows that are not 16-byte aligned (the 481-wide Darcy grid, the 83-long NS-3D axis)
    const bool rclass = !a_vec_ok && a.a_rs % 4 != 0 && a.a_rs >= 8 && (reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && a.N % 2 == 0 &&
                        a.ldc % 2 == 0 && (reinterpret_cast<uintptr_t>(a.C) & 7) == 0 && a.M >= 512 && cfg(CFG_KPIPE_ALIGN) != 0;
    TcKpImage img;
    int rc = tc_get_kpipe_image(a.B, a.ldb, a.K, a.N, &img, rclass ? (int)(a.a_rs % 4) : -1);
    if (rc) return rc;
    tc::KPipeParams p;
    p.A = a.A; p.lda = a.a_rs; p.R = a.M;
    p.Bimg = img.dev; p.C = a.C; p.ldc = a.ldc;
    p.N = a.N; p.K = rclass ? a.K + 3 : a.K; p.N_t = img.N_t; p.n_chunks = img.n_chunks; p.stages = stages;
    p.rclass = rclass ? 1 : 0; p.k_valid = a.K;
    p.m_tiles = rclass ? 4 * (((long)a.M + 511) / 512) : ((long)a.M + 127) / 128;
    int cols = 32;
    while (cols < 2 * img.N_t) cols *= 2;
    p.tmem_cols = cols;
    p.a_vec_ok = a_vec_ok;
    p.debug = cfg(CFG_KPIPE_DEBUG);
    int gx = num_sms();
    if (rclass) gx &= ~3;                   // a CTA must only ever see tiles of one row class (m_tiles is a multiple of 4 too)
    if ((long)gx > p.m_tiles) gx = (int)p.m_tiles;
    const size_t smem = tc::kpipe_smem_bytes(img.N_t, stages);
    // 16 loader warps (default: analysis 2.49 -> 2.19 ms per Darcy step, 2.13 with the row classes) or 8
    if (p.debug) return rclass ? launch_kpipe<16, true, true>(p, gx, smem, st) : launch_kpipe<16, false, true>(p, gx, smem, st);   // timing probes
    if (cfg(CFG_KPIPE_LW16)) return rclass ? launch_kpipe<16, true>(p, gx, smem, st) : launch_kpipe<16, false>(p, gx, smem, st);
    return rclass ? launch_kpipe<tc::kKpLoadWarps, true>(p, gx, smem, st) : launch_kpipe<tc::kKpLoadWarps, false>(p, gx, smem, st);
}

// ---- leading-axis complex transform on tcgen05 (tc_mid.cuh); switch mid_tc, default on --------------------------------
// Image of the real expansion of the complex matrix Mat[J, H] (device, interleaved):
//   B[(h,re), (j,re)] = Re   B[(h,im), (j,re)] = -Im   B[(h,re), (j,im)] = Im   B[(h,im), (j,im)] = Re
// per column tile: [chunk][hi | lo], half = (KC/4) x N_t x 16 bytes, element (n, k) at ((k%32)/4)*N_t*4 + n*4 + k%4.
int tc_get_mid_image(const float* Mat, int J, int H, TcMidImage* out) {
    std::lock_guard<std::mutex> lk(g_tc_mu);
    TcKey key{Mat, H, J, 0};
    auto it = g_tc_mid_images.find(key);
    if (it != g_tc_mid_images.end()) { *out = it->second; return 0; }
    TcMidImage img;
    const int N = 2 * J, K = 2 * H;
    img.n_tiles = (N + 255) / 256;
    const int per = (N + img.n_tiles - 1) / img.n_tiles;
    img.N_t = ((per + 15) / 16) * 16;
    img.n_chunks = (K + tc::kKC - 1) / tc::kKC;
    std::vector<float> hM((size_t)J * H * 2);
    cudaError_t e = cudaMemcpy(hM.data(), Mat, hM.size() * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return (int)e;
    const size_t half = (size_t)img.N_t * tc::kKC;
    std::vector<float> h((size_t)img.n_tiles * img.n_chunks * 2 * half, 0.0f);
    for (int j = 0; j < J; ++j)
        for (int hh = 0; hh < H; ++hh) {
            const float re = hM[((size_t)j * H + hh) * 2], im = hM[((size_t)j * H + hh) * 2 + 1];
            for (int ci = 0; ci < 2; ++ci)
                for (int co = 0; co < 2; ++co) {
                    const float b = (ci == co) ? re : (ci == 1 ? -im : im);
                    const int k = 2 * hh + ci, n = 2 * j + co;
                    const int t = n / img.N_t, nl = n % img.N_t;
                    const int c = k / tc::kKC, kk = k % tc::kKC;
                    const float hi = tf32_rn(b);
                    const float lo = tf32_rn(b - hi);
                    const size_t base = ((size_t)t * img.n_chunks + c) * 2 * half;
                    const size_t o = (size_t)(kk / 4) * img.N_t * 4 + (size_t)nl * 4 + (kk % 4);
                    h[base + o] = hi;
                    h[base + half + o] = lo;
                }
        }
    e = cudaMalloc(&img.dev, h.size() * 4);
    if (e != cudaSuccess) return (int)e;
    e = upload_sync(img.dev, h.data(), h.size() * 4);
    if (e != cudaSuccess) return (int)e;
    g_tc_mid_images[key] = img;
    *out = img;
    return 0;
}

bool mid_tc_enabled() { return cfg(CFG_MID_TC) != 0; }

// returns -1 when the shape is not taken (the caller runs the SIMT kernel)
int try_tc_mid(const MidArgs& a, cudaStream_t st) {
    if (!mid_tc_enabled() || !tc_enabled()) return -1;
    const long R = (long)a.O * a.I;
    if (a.H < 4 || a.J < 4 || R < 128) return -1;
    // 3xTF32 drops the lo*lo products (~2^-22 per term): over a contraction of 2*H real terms the error grows like sqrt(H) and
    // reached the forward tolerance (2e-5) at H = 1200 on B200; every grid of the reference's models is <= 481 (sweep: 512)
    if (a.H > 768) return -1;
    if ((reinterpret_cast<uintptr_t>(a.X) & 7) || (reinterpret_cast<uintptr_t>(a.Y) & 7)) return -1;
    bool capturing = false;
    {   // the first call for a matrix builds its image with synchronous copies: not inside a stream capture
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) capturing = true;
    }
    if (capturing) {
        std::lock_guard<std::mutex> lk(g_tc_mu);
        if (g_tc_mid_images.find(TcKey{a.Mat, a.H, a.J, 0}) == g_tc_mid_images.end()) return -1;
    }
    TcMidImage img;
    int rc = tc_get_mid_image(a.Mat, a.J, a.H, &img);
    if (rc) return rc;
    int stages = (int)((200 * 1024) / tc::kpipe_stage_bytes(img.N_t));
    if (stages > 4) stages = 4;
    if (stages < 2) return -1;
    tc::MidTcParams p;
    p.X = a.X; p.Y = a.Y; p.Bimg = img.dev;
    p.O = a.O; p.H = a.H; p.J = a.J; p.I = a.I;
    p.R = R; p.K = 2 * a.H;
    p.N_t = img.N_t; p.n_tiles = img.n_tiles; p.n_chunks = img.n_chunks; p.stages = stages;
    p.m_tiles = (R + 127) / 128;
    int cols = 32;
    while (cols < 2 * img.N_t) cols *= 2;
    p.tmem_cols = cols;
    static DeviceOnce configured;
    if (!configured.done()) {
        cudaError_t e = cudaFuncSetAttribute(tc::mid_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured.mark();
    }
    int gx = std::max(1, num_sms() / img.n_tiles);
    if ((long)gx > p.m_tiles) gx = (int)p.m_tiles;
    tc::mid_tc_kernel<<<dim3(gx, img.n_tiles), tc::kMidThreads, tc::kpipe_smem_bytes(img.N_t, stages), st>>>(p);
    CU_LAUNCH_CHECK();
    return 0;
}

// ---- per-mode complex contraction on tcgen05 (tc_cmm.cuh); switch cmm_tc, default on for shapes that fill a tile -----
bool cmm_tc_enabled() { return cfg(CFG_CMM_TC) != 0; }

int launch_tc_cmm4(const CmmArgs& a, cudaStream_t st) {
    tc::Cmm4Params p;
    p.a = a;
    p.ns_tiles = (a.M + 63) / 64;
    const int per = (a.M + p.ns_tiles - 1) / p.ns_tiles;
    p.N_t = ((per + 15) / 16) * 16;
    p.ms_tiles = (2 * a.N + 127) / 128;
    p.n_chunks = (a.K + 3) / 4;
    p.qg = (a.q_inner + 3) / 4;
    // split the reduction while the items leave more than half of the SMs idle (tc_cmm4.cuh): a divisor of the k-step count,
    // at least four k-steps (16 channels) per item.  Measured on B200: 1.08 -> 1.01 ms per NS-2D call
    p.ksplit = 1;
    {
        const long base = (long)p.ms_tiles * p.ns_tiles * a.ncorner * a.q_outer * p.qg;
        for (int ks = a.deterministic ? 0 : 8; ks >= 2; --ks)      // (forward outputs keep a fixed summation order)
                if (p.n_chunks % ks == 0 && p.n_chunks / ks >= 4 && base * ks <= (long)num_sms()) { p.ksplit = ks; break; }
    }
    p.n_chunks /= p.ksplit;
    int stages = (int)((200 * 1024) / tc::cmm4_stage_bytes(p.N_t));
    if (stages > 4) stages = 4;
    if (stages < 2) return -1;
    p.stages = stages;
    int cols = 32;
    while (cols < 2 * tc::kC4Modes * p.N_t) cols *= 2;
    if (cols > 512) return -1;
    p.tmem_cols = cols;
    p.items = (long)p.ms_tiles * p.ns_tiles * a.ncorner * a.q_outer * p.qg * p.ksplit;
    static DeviceOnce configured;
    if (!configured.done()) {
        cudaError_t e = cudaFuncSetAttribute(tc::cmm_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured.mark();
    }
    long gx = num_sms();
    if (gx > p.items) gx = p.items;
    if (p.ksplit > 1) {
        const long n = (long)a.M * a.N * a.q_outer * a.q_inner * a.ncorner;
        tc::cmm_zero_kernel<<<(unsigned)std::min<long>((n + 255) / 256, 4L * num_sms()), 256, 0, st>>>(a);
        CU_LAUNCH_CHECK();
    }
    tc::cmm_tc4_kernel<<<(unsigned)gx, tc::kC4Threads, tc::cmm4_smem_bytes(p.N_t, stages), st>>>(p);
    CU_LAUNCH_CHECK();
    return 0;
}

// returns -1 when the shape is not taken (the caller runs the SIMT kernels)
int try_tc_cmm(const CmmArgs& a, cudaStream_t st) {
    if (!cmm_tc_enabled() || !tc_enabled()) return -1;
    if (a.K < 4 || a.M < 4 || a.N < 8) return -1;
    if (cfg(CFG_CMM_TC) != 2) {
        // One UMMA per mode is 128 weight rows (n, re|im) x N_t samples: taken only when the operands fill at least half of
        // that tile and the reduction is long enough to amortise the operand images.  Measured on B200: NS-2D (batch 64,
        // 32..192 channels) 1.62 -> 1.28 ms per call; NS-3D (batch 8) got SLOWER (0.78 -> 1.47 ms) and stays on the SIMT kernel.
        const int ns_t = (a.M + 63) / 64, nt = ((((a.M + ns_t - 1) / ns_t) + 15) / 16) * 16;
        const double fill = (2.0 * a.N / (128.0 * ((2 * a.N + 127) / 128))) * ((double)a.M / ((double)nt * ns_t));
        if (fill < 0.5 || a.K < 16) return -1;
    }
    for (int c = 0; c < a.ncorner; ++c)
        if ((reinterpret_cast<uintptr_t>(a.A[c]) & 7) || (reinterpret_cast<uintptr_t>(a.B[c]) & 7) || (reinterpret_cast<uintptr_t>(a.C[c]) & 7))
            return -1;
    {   // four modes per item (tc_cmm4.cuh) when every operand is 16-byte friendly: whole-sector loads and stores
        bool even = ((a.a_sm | a.a_sk | a.a_sqo | a.b_sk | a.b_sn | a.b_sqo | a.c_sm | a.c_sn | a.c_sqo) & 1) == 0;
        for (int c = 0; c < a.ncorner; ++c)
            if ((reinterpret_cast<uintptr_t>(a.A[c]) | reinterpret_cast<uintptr_t>(a.B[c]) | reinterpret_cast<uintptr_t>(a.C[c])) & 15) even = false;
        // (falling back to one mode per item where four leave part of the machine idle -- the inner U-levels have 9 groups per
        // corner -- measured no difference; exp0 = 4 forces the one-mode kernel for the variant tests)
        if (even && !(cfg(CFG_EXP0) & 4)) {
            const int rc = launch_tc_cmm4(a, st);
            if (rc >= 0) return rc;
        }
    }
    tc::CmmTcParams p;
    p.a = a;
    p.ns_tiles = (a.M + 63) / 64;
    const int per = (a.M + p.ns_tiles - 1) / p.ns_tiles;
    p.N_t = ((per + 15) / 16) * 16;
    p.ms_tiles = (2 * a.N + 127) / 128;
    p.n_chunks = (a.K + tc::kCmKC - 1) / tc::kCmKC;
    int stages = (int)((200 * 1024) / tc::kpipe_stage_bytes(p.N_t));
    if (stages > 4) stages = 4;
    if (stages < 2) return -1;
    p.stages = stages;
    int cols = 32;
    while (cols < 2 * p.N_t) cols *= 2;
    p.tmem_cols = cols;
    p.items = (long)p.ms_tiles * p.ns_tiles * a.ncorner * a.q_outer * a.q_inner;
    static DeviceOnce configured;
    if (!configured.done()) {
        cudaError_t e = cudaFuncSetAttribute(tc::cmm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured.mark();
    }
    long gx = num_sms();
    if (gx > p.items) gx = p.items;
    tc::cmm_tc_kernel<<<(unsigned)gx, tc::kKpThreads, tc::kpipe_smem_bytes(p.N_t, stages), st>>>(p);
    CU_LAUNCH_CHECK();
    return 0;
}

// per-stream scratch for the weight image of the tensor-core channel mix (stream order makes reuse safe)
std::map<std::pair<int, cudaStream_t>, float*> g_conv_scratch;   // (device, stream): the legacy default stream is shared by all devices
constexpr size_t kConvScratchBytes = 160 * 1024;

int try_tc_conv(const GemmArgs& a, cudaStream_t st) {
    // GemmArgs view: C_b[M, N] = A[M, K] * B_b[K, N]  with M = out channels, N = pixels, K = in channels
    if (!tc_enabled() || !a.channel_mix || a.sA != 0 || a.epi != EPI_STORE || a.N % 4 != 0 || a.N < 128 || a.M < 8 || a.M > 256 || a.K < 8)
        return -1;
    if (a.ldb != a.N || a.ldc != a.N || a.sB != (long)a.K * a.N || a.sC != (long)a.M * a.N) return -1;
    if ((reinterpret_cast<uintptr_t>(a.B) & 15) || (reinterpret_cast<uintptr_t>(a.C) & 15)) return -1;
    const int N_t = ((a.M + 15) / 16) * 16;
    const int n_chunks = (a.K + tc::kKC - 1) / tc::kKC;
    const size_t b_bytes = (size_t)2 * N_t * n_chunks * tc::kKC * 4;
    if (b_bytes > kConvScratchBytes) return -1;
    int stages = (int)((220 * 1024 - b_bytes - 512) / (2 * tc::kCvAHalf));
    if (stages > 4) stages = 4;
    if (stages < 2) return -1;
    float* img = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_tc_mu);
        const std::pair<int, cudaStream_t> skey(current_device(), st);
        auto it = g_conv_scratch.find(skey);
        if (it == g_conv_scratch.end()) {
            cudaError_t e = cudaMalloc(&img, kConvScratchBytes);
            if (e != cudaSuccess) return (int)e;
            g_conv_scratch[skey] = img;
        } else img = it->second;
    }
    const int K_pad = n_chunks * tc::kKC;
    tc::conv_weight_image_kernel<<<(N_t * K_pad + 255) / 256, 256, 0, st>>>(a.A, a.a_rs, a.a_cs, a.M, a.K, N_t, K_pad, img);
    CU_LAUNCH_CHECK();
    tc::ConvTcParams p;
    p.X = a.B; p.sXb = a.sB; p.npix = a.N;
    p.Bimg = img; p.bias = a.bias; p.Y = a.C; p.sYb = a.sC;
    p.K = a.K; p.N = a.M; p.N_t = N_t; p.n_chunks = n_chunks; p.stages = stages; p.batch = a.batch;
    p.tiles_per_b = ((long)a.N + 127) / 128;
    p.n_tiles = p.tiles_per_b * a.batch;
    int cols = 32;
    while (cols < 2 * N_t) cols *= 2;
    p.tmem_cols = cols;
    int gx = num_sms();
    if ((long)gx > p.n_tiles) gx = (int)p.n_tiles;
    const tc::ConvTcParams& pc = p;
    const size_t smem = tc::conv_tc_smem_bytes(N_t, n_chunks, stages);
    auto launch = [&](auto lw) -> int {
        constexpr int LW = decltype(lw)::value;
        static DeviceOnce configured;
        if (!configured.done()) {
            cudaError_t e = cudaFuncSetAttribute(tc::conv1x1_tc_kernel<LW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return (int)e;
            configured.mark();
        }
        tc::conv1x1_tc_kernel<LW><<<gx, tc::cv_threads(LW), smem, st>>>(pc);
        CU_LAUNCH_CHECK();
        return 0;
    };
    // 16 loader warps: measured on B200 against 8: 1.63 -> 1.48 ms per Darcy step, 0.79 -> 0.75 ms per NS-3D step
    return launch(std::integral_constant<int, 16>());
}

template <int LW, bool DBG>
int launch_wgrad(const tc::WgradParams& p, unsigned gx, cudaStream_t st) {
    static DeviceOnce configured;
    if (!configured.done()) {
        cudaError_t e = cudaFuncSetAttribute(tc::wgrad_tc_kernel<LW, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured.mark();
    }
    tc::wgrad_tc_kernel<LW, DBG><<<gx, tc::wg_threads(LW), tc::wgrad_smem_bytes(p.stages), st>>>(p);
    CU_LAUNCH_CHECK();
    return 0;
}

int try_tc_wgrad(const GemmNtArgs& a, cudaStream_t st) {
    if (!tc_enabled() || a.M < 1 || a.N < 1 || a.M > 128 || a.N > 128 || a.K % 4 != 0 || a.K < 256) return -1;
    if (a.lda != a.K || a.ldb != a.K || a.sA != (long)a.M * a.K || a.sB != (long)a.N * a.K) return -1;
    if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.B) & 15)) return -1;
    tc::WgradParams p;
    p.G = a.A; p.sGb = a.sA; p.X = a.B; p.sXb = a.sB; p.dW = a.C; p.ldw = a.ldc; p.npix = a.K;
    p.M = a.M; p.N = a.N; p.N_t = ((a.N + 15) / 16) * 16; p.batch = a.batch;
    p.stages = 3;
    p.chunks_per_b = (int)((a.K + tc::kKC - 1) / tc::kKC);
    p.total_chunks = (long)p.chunks_per_b * a.batch;
    int cols = 32;
    while (cols < p.N_t) cols *= 2;
    p.tmem_cols = cols;
    long gx = num_sms();
    if (gx > (p.total_chunks + 3) / 4) gx = (p.total_chunks + 3) / 4;
    if (gx < 1) gx = 1;
    p.debug = cfg(CFG_WGRAD_DEBUG);   // timing probes (tools/wgrad_probe.py)
    // 16 loader warps: measured on B200 against 8 (a thread then owns 8 rows of a chunk, 168 registers): 1.66 -> 1.27 ms per
    // Darcy step, 1.14 -> 0.94 ms per NS-3D step
    if (p.debug) return launch_wgrad<16, true>(p, (unsigned)gx, st);
    return launch_wgrad<16, false>(p, (unsigned)gx, st);
}

// returns -1 when the shape does not qualify (caller falls back to the SIMT kernel)
int try_tc_rowgemm(const GemmArgs& a, cudaStream_t st) {
    if (!tc_enabled() || !a.b_const || a.batch != 1 || a.a_cs != 1 || a.bias || a.K > 64 || a.N < 8 || a.M < 1) return -1;   // N < 16 pads to one 16-column tile
    TcImage img;
    const int K_pad = ((a.K + 7) / 8) * 8;
    const int n_tiles = (a.N + 255) / 256;
    const int N_t = ((((a.N + n_tiles - 1) / n_tiles) + 15) / 16) * 16;
    const size_t smem = tc::rowgemm_smem_bytes(K_pad, N_t);
    if (smem > 220 * 1024) return -1;
    // odd row pitch: split the rows by parity so that both halves store aligned float2 (tc_rowgemm.cuh)
    const bool no_parity = cfg(CFG_ROWGEMM_PARITY) == 0;
    const bool c2_ok = a.epi != EPI_ACCUM_GELU || (reinterpret_cast<uintptr_t>(a.C2) & 7) == 0;
    const int parity = (!no_parity && (a.ldc & 1) && (reinterpret_cast<uintptr_t>(a.C) & 7) == 0 && c2_ok && n_tiles * N_t > a.N) ? 1 : 0;
    // row classes (tc_rowgemm.cuh): every lane quad on one whole 32-byte sector when the pitch is not a multiple of 8 floats
    int nclass = parity ? 2 : 1, shift_mul = 1;
    {
        const int g8 = (a.ldc % 8 == 0) ? 8 : (a.ldc % 4 == 0) ? 4 : (a.ldc % 2 == 0) ? 2 : 1;          // gcd(ldc, 8)
        const bool c32 = (reinterpret_cast<uintptr_t>(a.C) & 31) == 0 && (a.epi != EPI_ACCUM_GELU || (reinterpret_cast<uintptr_t>(a.C2) & 31) == 0);
        if (!no_parity && cfg(CFG_ROWGEMM_PARITY) != 2 && g8 < 8 && c32 && n_tiles * N_t >= a.N + 7 && a.M >= 128L * (8 / g8)) { nclass = 8 / g8; shift_mul = (int)(a.ldc & 7); }
    }
    int rc = tc_get_rowgemm_image(a.B, a.ldb, a.K, a.N, nclass == 1 ? 1 : (shift_mul == 1 && nclass == 2 ? 2 : 8), &img);
    if (rc) return rc;
    tc::RowGemmParams p;
    p.A = a.A; p.lda = a.a_rs; p.R = a.M;
    p.Bimg = img.dev;
    p.C = a.C; p.C2 = a.C2; p.ldc = a.ldc;
    p.N = a.N; p.K = a.K; p.K_pad = img.K_pad; p.n_tiles = img.n_tiles; p.N_t = img.N_t;
    p.nclass = nclass; p.cls_log2 = nclass == 8 ? 3 : nclass == 4 ? 2 : nclass == 2 ? 1 : 0; p.shift_mul = shift_mul;
    p.m_tiles = (long)nclass * (((long)a.M + 128L * nclass - 1) / (128L * nclass));
    p.epi = a.epi;
    int cols = 32;
    while (cols < 2 * img.N_t) cols *= 2;
    p.tmem_cols = cols;
    p.a_vec_ok = (a.a_rs % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0) && (a.K % 4 == 0);
    switch (a.epi) {
        case EPI_STORE: return launch_rowgemm<EPI_STORE>(p, smem, st);
        case EPI_ACCUM: return launch_rowgemm<EPI_ACCUM>(p, smem, st);
        case EPI_ACCUM_GELU: return launch_rowgemm<EPI_ACCUM_GELU>(p, smem, st);
        case EPI_ACCUM_GELU_INPLACE: return launch_rowgemm<EPI_ACCUM_GELU_INPLACE>(p, smem, st);
    }
    return -1;
}

}  // namespace

// =====================================================================================================
// backend.h implementation
// =====================================================================================================
const char* be_name() { return "cuda-sm100a"; }
int be_current_device() { return current_device(); }
const char* be_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }

void be_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (on && !g_prof_on) {
        for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        g_prof.clear();
        g_prof_scopes.clear();
    }
    g_prof_on = on != 0;
}

int be_profile_enabled() { return g_prof_on ? 1 : 0; }

void be_profile_scope_begin(const char* label, double bytes, double flops) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    int idx = -1;
    for (size_t i = 0; i < g_prof_scopes.size(); ++i)
        if (g_prof_scopes[i].label == label) { idx = (int)i; break; }
    if (idx < 0) {
        idx = (int)g_prof_scopes.size();
        g_prof_scopes.push_back(ProfScopeInfo{label});
    }
    g_prof_scopes[idx].calls += 1;
    g_prof_scopes[idx].bytes += bytes;
    g_prof_scopes[idx].flops += flops;
    t_prof_scope = idx;
}

void be_profile_scope_end() { t_prof_scope = -1; }

size_t be_profile_report_scopes(char* buf, size_t cap) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::vector<double> ms(g_prof_scopes.size(), 0.0);
    std::vector<long> n(g_prof_scopes.size(), 0);
    for (auto& r : g_prof) {
        if (r.scope < 0 || r.scope >= (int)g_prof_scopes.size()) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) continue;
        ms[r.scope] += t;
        n[r.scope] += 1;
    }
    std::string out = "{";
    for (size_t i = 0; i < g_prof_scopes.size(); ++i) {
        char line[384];
        snprintf(line, sizeof line, "%s\"%s\": {\"calls\": %ld, \"launches\": %ld, \"ms\": %.6f, \"bytes\": %.0f, \"flops\": %.0f}",
                 i ? ", " : "", g_prof_scopes[i].label.c_str(), g_prof_scopes[i].calls, n[i], ms[i], g_prof_scopes[i].bytes,
                 g_prof_scopes[i].flops);
        out += line;
    }
    out += "}";
    if (buf && cap) {
        size_t len = out.size() < cap - 1 ? out.size() : cap - 1;
        memcpy(buf, out.data(), len);
        buf[len] = 0;
    }
    return out.size();
}

size_t be_profile_report(char* buf, size_t cap) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    struct Agg { long n = 0; double ms = 0, bytes = 0, flops = 0; };
    std::map<std::string, Agg> agg;
    for (auto& r : g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) continue;
        Agg& a = agg[r.tag];
        a.n += 1; a.ms += ms; a.bytes += r.bytes; a.flops += r.flops;
    }
    std::string out = "{";
    bool first = true;
    for (auto& kv : agg) {
        char line[256];
        snprintf(line, sizeof line, "%s\"%s\": {\"launches\": %ld, \"ms\": %.6f, \"bytes\": %.0f, \"flops\": %.0f}",
                 first ? "" : ", ", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.bytes, kv.second.flops);
        out += line;
        first = false;
    }
    out += "}";
    if (buf && cap) {
        size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return out.size();
}

long be_launch_count() { return g_launches.load(); }

void be_range_push(const char* label) {
    if (cfg(CFG_NVTX)) nvtxRangePushA(label);
}
void be_range_pop() {
    if (cfg(CFG_NVTX)) nvtxRangePop();
}

// ---- side stream + event ring for fork / join -----------------------------------------------------------
namespace {
struct SideStreams {
    std::mutex mu;
    std::map<int, cudaStream_t> streams;       // one per device
    std::map<int, std::vector<cudaEvent_t>> events;
    std::map<int, unsigned> next;
    bool disabled = false;
} g_side;

cudaEvent_t side_event(int dev) {
    auto& ring = g_side.events[dev];
    if (ring.empty()) {
        ring.resize(32);
        for (auto& e : ring) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    }
    return ring[g_side.next[dev]++ % ring.size()];
}
}  // namespace

stream_t be_side_stream(int which) {
    std::lock_guard<std::mutex> lk(g_side.mu);
    if (g_side.disabled) return nullptr;
    // Round 1 measured no gain from the overlap (every kernel filled the machine: 29.8 vs 30.1 ms per Darcy step).  The kernels
    // have since become latency-bound persistent ones with 20-30 % of the warp slots occupied, and two of them side by side hide
    // each other's stalls: 20.50 -> 20.27 ms (Darcy), 10.82 -> 10.43 (NS-3D), 4.76 -> 4.52 (NS-2D call).  Default on; bench.py
    // switches it off for its per-kernel profiling steps (concurrent kernels blur the per-kernel event timings).
    if (!cfg(CFG_OVERLAP)) return nullptr;
    if ((which & 1) && cfg(CFG_OVERLAP) == 2) return nullptr;   // 2: the branch overlap only, no third stream
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    const int key = dev * 2 + (which & 1);
    auto it = g_side.streams.find(key);
    if (it == g_side.streams.end()) {
        cudaStream_t st = nullptr;
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { g_side.disabled = true; return nullptr; }
        it = g_side.streams.emplace(key, st).first;
    }
    return (stream_t)it->second;
}

static int link_streams(cudaStream_t from, cudaStream_t to) {
    // everything enqueued on `from` so far happens before anything enqueued on `to` from now on
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    cudaEvent_t ev;
    {
        std::lock_guard<std::mutex> lk(g_side.mu);
        ev = side_event(dev);
    }
    e = cudaEventRecord(ev, from);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaStreamWaitEvent(to, ev, 0);
}
int be_fork(stream_t main_stream, stream_t side) { return link_streams(S(main_stream), S(side)); }
int be_join(stream_t main_stream, stream_t side) { return link_streams(S(side), S(main_stream)); }

int be_upload(void** dptr, const void* host, size_t bytes) {
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 4);
    if (e != cudaSuccess) return (int)e;
    e = upload_sync(*dptr, host, bytes);
    return (int)e;
}
void be_free(void* d) {
    tc_forget(d);
    cudaFree(d);
}
int be_memset(void* d, int v, size_t bytes, stream_t s) { return (int)cudaMemsetAsync(d, v, bytes, S(s)); }

int be_gemm(const GemmArgs& a, stream_t s) {
    GemmK k;
    k.A = a.A; k.a_rs = a.a_rs; k.a_cs = a.a_cs; k.sA = a.sA;
    k.B = a.B; k.b_rs = a.ldb; k.b_cs = 1; k.sB = a.sB;
    k.C = a.C; k.ldc = a.ldc; k.sC = a.sC; k.C2 = a.C2; k.bias = a.bias;
    k.M = a.M; k.N = a.N; k.K = a.K; k.ksplit = 1; k.kchunk = a.K; k.epi = a.epi;
    const double mn = (double)a.M * a.N * a.batch;
    // algorithmic bytes: A and B read once (a batch-shared operand once), C written once (+ read when accumulating,
    // + second output for the fused GELU)
    double bytes = 4.0 * ((double)a.M * a.K * (a.sA ? a.batch : 1) + (double)a.K * a.N * (a.sB ? a.batch : 1) + mn);
    if (a.epi != EPI_STORE) bytes += 4.0 * mn;
    if (a.epi == EPI_ACCUM_GELU) bytes += 4.0 * mn;
    ProfScope ps(a.tag, bytes, 2.0 * mn * a.K, S(s));
    int rc = try_tc_rowgemm(a, S(s));
    if (rc >= 0) return rc;
    rc = try_tc_kpipe(a, S(s));
    if (rc >= 0) return rc;
    rc = try_tc_conv(a, S(s));
    if (rc >= 0) return rc;
    return dispatch_gemm(k, a.batch, S(s));
}

int be_gemm_nt_atomic(const GemmNtArgs& a, stream_t s) {
    GemmK k;
    k.A = a.A; k.a_rs = a.lda; k.a_cs = 1; k.sA = a.sA;
    k.B = a.B; k.b_rs = 1; k.b_cs = a.ldb; k.sB = a.sB;
    k.C = a.C; k.ldc = a.ldc; k.sC = 0; k.C2 = nullptr; k.bias = nullptr;
    k.M = a.M; k.N = a.N; k.K = a.K; k.epi = EPI_ATOMIC;
    // split K so that batch * ksplit CTAs per output tile fill the machine
    const int tiles = ((a.M + 63) / 64) * ((a.N + 63) / 64);
    int want = (148 * 4 + tiles * a.batch - 1) / (tiles * a.batch);
    int kchunk = (a.K + want - 1) / want;
    kchunk = ((kchunk + 63) / 64) * 64;
    if (kchunk < 256) kchunk = 256;
    k.kchunk = kchunk;
    k.ksplit = (a.K + kchunk - 1) / kchunk;
    ProfScope ps("conv1x1_wgrad", 4.0 * ((double)a.M * a.K + (double)a.N * a.K) * a.batch + 4.0 * a.M * a.N,
                 2.0 * a.M * a.N * (double)a.K * a.batch, S(s));
    const int rc = try_tc_wgrad(a, S(s));
    if (rc >= 0) return rc;
    // (tiling operands wider than 128 channels -- the 192-channel NS-2D levels -- into four tensor-core launches measured no
    // faster than the SIMT kernel below: 0.48 against 0.45 ms per NS-2D call, the operands are a few MB)
    return dispatch_gemm(k, a.batch, S(s));
}

namespace {
template <int TJ, int TI>
int launch_mid2(const Mid2K& k, size_t smem, long blocks, int threads, cudaStream_t st) {
    int rc = ensure_smem(mid2_kernel<TJ, TI>, smem);
    if (rc) return rc;
    mid2_kernel<TJ, TI><<<(unsigned)blocks, threads, smem, st>>>(k);
    CU_LAUNCH_CHECK();
    return 0;
}
}  // namespace

int be_mid(const MidArgs& a, stream_t s) {
    if (a.O <= 0 || a.J <= 0 || a.I <= 0) return 0;
    Mid2K k;
    k.X = reinterpret_cast<const float2*>(a.X);
    k.Mat = reinterpret_cast<const float2*>(a.Mat);
    k.Y = reinterpret_cast<float2*>(a.Y);
    k.O = a.O; k.H = a.H; k.J = a.J; k.I = a.I;
    // thread tile: TI = 2 columns when there are any to pair; TJ = the largest of 4, 3, 2 that pads J by <= 6 %
    const int TI = a.I >= 2 ? 2 : 1;
    int TJ = 2;
    for (int cand = 4; cand >= 2; --cand) {
        const int padded = (a.J + cand - 1) / cand * cand;
        if (padded * 100 <= a.J * 106 || cand == 2) { TJ = cand; break; }
    }
    const int ni_all = (a.I + TI - 1) / TI;
    k.ni = std::min(ni_all, 32);
    k.tilesI = (ni_all + k.ni - 1) / k.ni;
    const int nj_all = (a.J + TJ - 1) / TJ;
    const int nj_cap = std::max(1, 256 / k.ni);
    k.tilesJ = (nj_all + nj_cap - 1) / nj_cap;
    k.nj = (nj_all + k.tilesJ - 1) / k.tilesJ;            // balanced tiles
    k.ppc = std::max(1, 256 / (k.nj * k.ni));
    if ((long)k.ppc > a.O) k.ppc = (int)a.O;
    // (planes sharing a CTA share its Mat tile: splitting them up for more CTAs measured slower, 2.2 vs 1.8 ms per step)
    const int threads = 256;   // compile-time strides in the staging loops
    k.HK = a.H <= 48 ? a.H : 32;
    const int JT = k.nj * TJ, IT = k.ni * TI, JTP = JT | 1;
    size_t ms = (size_t)k.HK * JTP;
    ms += ms & 1;
    const size_t smem = 2 * (ms + (size_t)k.ppc * k.HK * IT) * sizeof(float2);   // double buffered
    const long blocks = (long)((a.O + k.ppc - 1) / k.ppc) * k.tilesJ * k.tilesI;
    ProfScope ps("dft_mid", 8.0 * ((double)a.O * a.I * (a.H + a.J) + (double)a.J * a.H), 8.0 * a.O * (double)a.J * a.H * a.I, S(s));
    {
        const int rc = try_tc_mid(a, S(s));   // opt-in tcgen05 path (UNO_B200_MID_TC=1)
        if (rc >= 0) return rc;
    }
    switch (TJ * 10 + TI) {
        case 42: return launch_mid2<4, 2>(k, smem, blocks, threads, S(s));
        case 32: return launch_mid2<3, 2>(k, smem, blocks, threads, S(s));
        case 22: return launch_mid2<2, 2>(k, smem, blocks, threads, S(s));
        case 41: return launch_mid2<4, 1>(k, smem, blocks, threads, S(s));
        case 31: return launch_mid2<3, 1>(k, smem, blocks, threads, S(s));
        default: return launch_mid2<2, 1>(k, smem, blocks, threads, S(s));
    }
}

int be_cmm(const CmmArgs& a, stream_t s) {
    if (a.M <= 0 || a.N <= 0 || a.q_inner <= 0 || a.q_outer <= 0) return 0;
    const int chunks = (a.q_inner + 31) / 32;
    const double q = (double)a.q_inner * a.q_outer * a.ncorner;
    ProfScope ps("mode_contraction", 8.0 * q * ((double)a.M * a.K + (double)a.K * a.N + (double)a.M * a.N), 8.0 * q * a.M * a.N * a.K, S(s));
    {
        const int rc = try_tc_cmm(a, S(s));   // opt-in tcgen05 path (UNO_B200_CMM_TC=1)
        if (rc >= 0) return rc;
    }
    if (a.M >= 16 && a.N >= 16) {
        // enough rows and columns to fill 32 x 32 tiles: the shared-memory tiled kernel (4 modes per CTA)
        const int qchunks = (a.q_inner + kC2Q - 1) / kC2Q;
        const size_t smem = (size_t)2 * kC2BufElems * sizeof(float2);
        int rc = ensure_smem(cmm2_kernel, smem);
        if (rc) return rc;
        dim3 grid2((unsigned)(qchunks * a.q_outer * a.ncorner), (unsigned)((a.M + kC2M - 1) / kC2M), (unsigned)((a.N + kC2N - 1) / kC2N));
        cmm2_kernel<<<grid2, 256, smem, S(s)>>>(a, qchunks);
        CU_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid((unsigned)(chunks * a.q_outer * a.ncorner), (unsigned)((a.M + 3) / 4), (unsigned)((a.N + 15) / 16));
    cmm_kernel<<<grid, dim3(32, 4), 0, S(s)>>>(a, chunks);
    CU_LAUNCH_CHECK();
    return 0;
}

int be_banded(const BandedArgs& a, stream_t s) {
    const long total = a.outer * a.n_out * a.inner;
    if (total <= 0) return 0;
    ProfScope ps("resample_banded", 4.0 * ((double)a.outer * a.inner * (a.n_in + a.n_out)), 2.0 * total * a.taps, S(s));
    banded_kernel<<<grid_for((size_t)total, 256, 148 * 32), 256, 0, S(s)>>>(a, total);
    CU_LAUNCH_CHECK();
    return 0;
}


namespace {
template <int G0, int W0, int G1, int W1>
int launch_resample2d(const Banded2DArgs& a, cudaStream_t st) {
    Resample2K k;
    k.x = a.x; k.y = a.y; k.planes = a.planes;
    k.n_in0 = a.n_in0; k.n_out0 = a.n_out0; k.n_in1 = a.n_in1; k.n_out1 = a.n_out1;
    k.gs0 = a.gs0; k.D0 = a.D0; k.ng0 = a.ng0;
    k.gs1 = a.gs1; k.D1 = a.D1; k.ng1 = a.ng1;
    k.TH = a.tile_groups0 * G0;
    k.RIN = (a.tile_span0 + 3) & ~3;          // multiple of 4 keeps the weight images 16-byte aligned
    k.ldin = a.tile_span1 | 1;                // odd pitch: lanes sweeping rows hit distinct banks
    k.tiles_h = (a.ng0 + a.tile_groups0 - 1) / a.tile_groups0;
    k.tiles_w = (a.n_out1 + rs_tile_w(G1) - 1) / rs_tile_w(G1);
    const size_t smem = resample2d_smem(k.RIN, k.ldin, k.TH, G0, W0, G1, W1);
    if (smem > 160 * 1024) return -1;
    int rc = ensure_smem(resample2d_kernel<G0, W0, G1, W1, 2>, smem);
    if (rc) return rc;
    if (a.planes > 65535) return -1;          // planes ride on grid.z
    ProfScope ps("resample_banded", 4.0 * a.planes * ((double)a.n_in0 * a.n_in1 + (double)a.n_out0 * a.n_out1),
                 2.0 * a.planes * ((double)a.n_in0 * a.n_out1 * W1 + (double)a.n_out0 * a.n_out1 * W0), st);
    const dim3 grid((unsigned)k.tiles_w, (unsigned)k.tiles_h, (unsigned)a.planes);
    resample2d_kernel<G0, W0, G1, W1, 2><<<grid, 256, smem, st>>>(k);
    CU_LAUNCH_CHECK();
    return 0;
}

// returns -1 when the register-blocked kernel does not take the shape
int try_resample2d(const Banded2DArgs& a, cudaStream_t st) {
    if (!a.D0 || !a.D1 || a.tile_groups0 < 1) return -1;
    const int c0 = a.G0 == 8 ? 0 : (a.W0 == 8 ? 1 : 2), c1 = a.G1 == 8 ? 0 : (a.W1 == 8 ? 1 : 2);
    switch (c0 * 3 + c1) {
        case 0: return launch_resample2d<8, 8, 8, 8>(a, st);
        case 1: return launch_resample2d<8, 8, 4, 8>(a, st);
        case 2: return launch_resample2d<8, 8, 4, 16>(a, st);
        case 3: return launch_resample2d<4, 8, 8, 8>(a, st);
        case 4: return launch_resample2d<4, 8, 4, 8>(a, st);
        case 5: return launch_resample2d<4, 8, 4, 16>(a, st);
        case 6: return launch_resample2d<4, 16, 8, 8>(a, st);
        case 7: return launch_resample2d<4, 16, 4, 8>(a, st);
        default: return launch_resample2d<4, 16, 4, 16>(a, st);
    }
}
}  // namespace

int be_banded2d(const Banded2DArgs& a, stream_t s) {
    if (a.planes <= 0) return 0;
    {
        const int rc = try_resample2d(a, S(s));
        if (rc >= 0) return rc;
    }
    const int RIN = a.span0, CIN = a.span1;
    const int ldin = CIN | 1;   // odd row pitch
    const size_t smem = ((size_t)RIN * ldin + (size_t)RIN * kB2MidLd + (size_t)kB2TH * a.taps0 + (size_t)kB2TW * a.taps1 + kB2TH + kB2TW) * 4;
    if (smem > 200 * 1024) {
        // very large scale factors: two passes through scratch (shrinking axis first)
        BandedArgs l, m;
        if (a.n_out1 <= a.n_in1) {
            l.x = a.x; l.y = a.tmp; l.start = a.start1; l.w = a.w1; l.n_in = a.n_in1; l.n_out = a.n_out1; l.taps = a.taps1; l.outer = a.planes * a.n_in0; l.inner = 1;
            m.x = a.tmp; m.y = a.y; m.start = a.start0; m.w = a.w0; m.n_in = a.n_in0; m.n_out = a.n_out0; m.taps = a.taps0; m.outer = a.planes; m.inner = a.n_out1;
            int rc = be_banded(l, s);
            return rc ? rc : be_banded(m, s);
        }
        m.x = a.x; m.y = a.tmp; m.start = a.start0; m.w = a.w0; m.n_in = a.n_in0; m.n_out = a.n_out0; m.taps = a.taps0; m.outer = a.planes; m.inner = a.n_in1;
        l.x = a.tmp; l.y = a.y; l.start = a.start1; l.w = a.w1; l.n_in = a.n_in1; l.n_out = a.n_out1; l.taps = a.taps1; l.outer = a.planes * a.n_out0; l.inner = 1;
        int rc = be_banded(m, s);
        return rc ? rc : be_banded(l, s);
    }
    static DeviceOnce configured;
    if (!configured.done()) {
        cudaError_t e = cudaFuncSetAttribute(banded2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured.mark();
    }
    const int tiles_h = (a.n_out0 + kB2TH - 1) / kB2TH, tiles_w = (a.n_out1 + kB2TW - 1) / kB2TW;
    const long blocks = a.planes * tiles_h * tiles_w;
    ProfScope ps("resample_banded", 4.0 * a.planes * ((double)a.n_in0 * a.n_in1 + (double)a.n_out0 * a.n_out1),
                 2.0 * a.planes * ((double)a.n_in0 * a.n_out1 * a.taps1 + (double)a.n_out0 * a.n_out1 * a.taps0), S(s));
    banded2d_kernel<<<(unsigned)blocks, 256, smem, S(s)>>>(a, tiles_h, tiles_w, RIN, ldin);
    CU_LAUNCH_CHECK();
    return 0;
}

int be_gelu_fwd(const float* pre, float* y, size_t n, stream_t s) {
    if (!n) return 0;
    ProfScope ps("gelu_fwd", 8.0 * n, 0, S(s));
    gelu_fwd_kernel<<<grid_for(n, 256), 256, 0, S(s)>>>(pre, y, n);
    CU_LAUNCH_CHECK();
    return 0;
}
namespace {
// every source either contiguous (plane p starts at p*L, like pre and g) or 16-byte friendly (L and batch strides multiples of 4)
inline bool upgrad_vec_ok(const UpGrad& gy, int C, long L, const void* a, const void* b) {
    uintptr_t bits = reinterpret_cast<uintptr_t>(gy.p0) | reinterpret_cast<uintptr_t>(gy.p1) | reinterpret_cast<uintptr_t>(a) |
                     reinterpret_cast<uintptr_t>(b);
    if (bits & 15) return false;
    const long cl = (long)C * L;
    const bool contiguous = gy.bs0 == cl && (!gy.p1 || gy.bs1 == cl);
    const bool quad = (L & 3) == 0 && (gy.bs0 & 3) == 0 && (!gy.p1 || (gy.bs1 & 3) == 0);
    return contiguous || quad;
}
inline UpGradK upgrad_k(const UpGrad& gy) { return UpGradK{gy.p0, gy.bs0, gy.p1, gy.bs1}; }
}  // namespace

int be_gelu_bwd_bias(const UpGrad& gy, const float* pre, float* g, long planes, int C, long L, float* gbias, float alpha,
                     stream_t s) {
    if (planes <= 0 || L <= 0) return 0;
    const int nsrc = gy.p1 ? 2 : 1;
    ProfScope ps("gelu_bwd", 4.0 * (nsrc + (pre ? 2 : 1)) * planes * L, 0, S(s));
    // enough CTAs per plane to fill the machine, few enough that the atomics stay negligible
    unsigned gy_ = (unsigned)std::max<long>(1, std::min<long>((L + 2047) / 2048, (148L * 32 + planes - 1) / planes));
    const int vec_ok = upgrad_vec_ok(gy, C, L, pre, g) ? 1 : 0;
    const dim3 grid((unsigned)planes, gy_);
    const UpGradK k = upgrad_k(gy);
    if (pre) {
        if (gy.p1) gelu_bwd_bias_kernel<true, true><<<grid, 256, 0, S(s)>>>(k, pre, g, C, L, gbias, alpha, vec_ok);
        else gelu_bwd_bias_kernel<true, false><<<grid, 256, 0, S(s)>>>(k, pre, g, C, L, gbias, alpha, vec_ok);
    } else {
        if (gy.p1) gelu_bwd_bias_kernel<false, true><<<grid, 256, 0, S(s)>>>(k, pre, g, C, L, gbias, alpha, vec_ok);
        else gelu_bwd_bias_kernel<false, false><<<grid, 256, 0, S(s)>>>(k, pre, g, C, L, gbias, alpha, vec_ok);
    }
    CU_LAUNCH_CHECK();
    return 0;
}
int be_plane_stats(const float* x, float* stats, long planes, long L, float eps, stream_t s) {
    if (planes <= 0) return 0;
    ProfScope ps("instnorm_stats", 4.0 * planes * L, 0, S(s));
    plane_stats_kernel<<<(unsigned)planes, 512, 0, S(s)>>>(x, stats, L, eps);
    CU_LAUNCH_CHECK();
    return 0;
}
// Shared-memory budget per CTA of the InstanceNorm cluster kernels: 72 KB (three CTAs per SM) is what the Darcy levels were
// tuned with.  Planes that do not fit 8 x 72 KB -- the NS-3D levels -- get a second try with 200 KB per CTA (one CTA per SM)
// before the two-kernel path (measured: NS-3D step 16.1 -> 14.9 ms, InstanceNorm backward 0.98 -> 0.41 ms).
inline int norm_cluster_pick(long L, int bytes_per_elem, int* slice) {
    int cs = norm_cluster_size(L, bytes_per_elem, 72 * 1024, slice);
    if (cs > 0) return cs;
    if (cfg(CFG_NORM_BIG_CLUSTER)) return norm_cluster_size(L, bytes_per_elem, 200 * 1024, slice);
    return 0;
}

int be_norm_fused_fwd(const float* x, float* stats, const float* gamma, const float* beta, float* y, long planes, int C,
                      long L, float eps, int non_lin, stream_t s) {
    if (planes <= 0) return 0;
    int slice = 0;
    const int cs = planes * 8 <= 0x7fffffffL ? norm_cluster_pick(L, 4, &slice) : 0;
    if (cs > 0) {
        // one sweep: the plane stays in the shared memory of a cluster of `cs` CTAs between statistics and normalisation
        ProfScope ps("instnorm_gelu_fwd", 8.0 * planes * L, 0, S(s));
        const int rc = launch_cluster(norm_fwd_cluster_kernel, planes, cs, (size_t)slice * 4, S(s), x, stats, gamma, beta, y, C, L, slice,
                                      eps, non_lin);
        if (rc) return rc;
        CU_LAUNCH_CHECK();
        return 0;
    }
    int rc = be_plane_stats(x, stats, planes, L, eps, s);
    return rc ? rc : be_norm_act_fwd(x, stats, gamma, beta, y, planes, C, L, non_lin, s);
}
int be_norm_act_fwd(const float* x, const float* stats, const float* gamma, const float* beta, float* y,
                    long planes, int C, long L, int non_lin, stream_t s) {
    if (planes <= 0) return 0;
    unsigned gx = grid_for((size_t)L, 256, 64);
    ProfScope ps("instnorm_gelu_fwd", 8.0 * planes * L, 0, S(s));
    norm_act_fwd_kernel<<<dim3((unsigned)planes, gx), 256, 0, S(s)>>>(x, stats, gamma, beta, y, C, L, non_lin);
    CU_LAUNCH_CHECK();
    return 0;
}
int be_norm_act_bwd(const UpGrad& gyu, const float* x, const float* stats, const float* gamma, const float* beta,
                    float* g, float* ggamma, float* gbeta, long planes, int C, long L, int non_lin, stream_t s) {
    if (planes <= 0) return 0;
    const UpGradK gy = upgrad_k(gyu);
    ProfScope ps("instnorm_gelu_bwd", 4.0 * (gyu.p1 ? 4 : 3) * planes * L, 0, S(s));
    int slice = 0;
    const int cs = planes * 8 <= 0x7fffffffL ? norm_cluster_pick(L, 8, &slice) : 0;
    if (cs > 0) {
        const int rc = launch_cluster(norm_bwd_cluster_kernel, planes, cs, (size_t)slice * 8, S(s), gy, x, stats, gamma, beta, g, ggamma,
                                      gbeta, C, L, slice, non_lin);
        if (rc) return rc;
        CU_LAUNCH_CHECK();
        return 0;
    }
    norm_act_bwd_kernel<<<(unsigned)planes, 512, 0, S(s)>>>(gy, x, stats, gamma, beta, g, ggamma, gbeta, C, L, non_lin);
    CU_LAUNCH_CHECK();
    return 0;
}
int be_channel_sum(const float* x, float* out, long planes, int C, long L, float alpha, stream_t s) {
    if (planes <= 0) return 0;
    ProfScope ps("bias_grad", 4.0 * planes * L, 0, S(s));
    channel_sum_kernel<<<(unsigned)planes, 512, 0, S(s)>>>(x, out, C, L, alpha);
    CU_LAUNCH_CHECK();
    return 0;
}
int be_add_channel_const(float* y, const float* v, float alpha, long planes, int C, long L, stream_t s) {
    if (planes <= 0) return 0;
    unsigned gx = grid_for((size_t)L, 256, 64);
    ProfScope ps("bias_add", 8.0 * planes * L, 0, S(s));
    add_channel_const_kernel<<<dim3((unsigned)planes, gx), 256, 0, S(s)>>>(y, v, alpha, C, L);
    CU_LAUNCH_CHECK();
    return 0;
}

// =====================================================================================================
// model glue: fused lift / projection MLPs (pixel_mlp.cuh)
// =====================================================================================================
namespace {

PixGeom pix_geom(const int* n, const int* N, const int* lo) {
    PixGeom g;
    g.n0 = n[0]; g.n1 = n[1]; g.n2 = n[2];
    g.N0 = N[0]; g.N1 = N[1]; g.N2 = N[2];
    g.lo0 = lo[0]; g.lo1 = lo[1]; g.lo2 = lo[2];
    g.nraw = (long)n[0] * n[1] * n[2];
    g.npad = (long)N[0] * N[1] * N[2];
    return g;
}

LiftK lift_k(const LiftArgs& a) {
    LiftK k;
    k.g = pix_geom(a.n, a.N, a.lo);
    k.batch = a.batch; k.raw_ch = a.raw_ch; k.grid_ch = a.grid_ch; k.cin = a.raw_ch + a.grid_ch; k.hid = a.hid; k.out_ch = a.out_ch;
    k.a = a.a; k.grid = a.grid; k.w_a = a.w_a; k.b_a = a.b_a; k.w_b = a.w_b; k.b_b = a.b_b;
    k.h = a.h; k.gh = a.gh; k.gh2 = a.gh2; k.ga = a.ga; k.gw_a = a.gw_a; k.gb_a = a.gb_a; k.gw_b = a.gw_b; k.gb_b = a.gb_b;
    return k;
}

template <int CIN, int HID>
int launch_lift(const LiftK& k, bool bwd, cudaStream_t st) {
    if (!bwd) {
        const size_t smem = lift_fwd_smem(CIN, HID, k.out_ch);
        int rc = ensure_smem(lift_fwd_kernel<CIN, HID>, smem);
        if (rc) return rc;
        const long total = (long)k.batch * k.g.npad;
        const unsigned grid = grid_for((size_t)total, kPixTP, 148 * 8);
        lift_fwd_kernel<CIN, HID><<<grid, kPixTP, smem, st>>>(k);
    } else {
        const size_t smem = lift_bwd_smem(CIN, HID, k.out_ch);
        int rc = ensure_smem(lift_bwd_kernel<CIN, HID>, smem);
        if (rc) return rc;
        const long ntiles = ((long)k.batch * k.g.nraw + kPixTP - 1) / kPixTP;
        int per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lift_bwd_kernel<CIN, HID>, kPixTP, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        const unsigned grid = (unsigned)std::min<long>(ntiles, 148L * per_sm);   // persistent: every CTA resident
        lift_bwd_kernel<CIN, HID><<<grid, kPixTP, smem, st>>>(k, ntiles);
    }
    CU_LAUNCH_CHECK();
    return 0;
}

int dispatch_lift(const LiftK& k, bool bwd, cudaStream_t st) {
    const int ci = k.cin <= 4 ? 0 : (k.cin <= 8 ? 1 : 2);
    const int hi = k.hid <= 16 ? 0 : 1;
    switch (ci * 2 + hi) {
        case 0: return launch_lift<4, 16>(k, bwd, st);
        case 1: return launch_lift<4, 32>(k, bwd, st);
        case 2: return launch_lift<8, 16>(k, bwd, st);
        case 3: return launch_lift<8, 32>(k, bwd, st);
        case 4: return launch_lift<16, 16>(k, bwd, st);
        default: return launch_lift<16, 32>(k, bwd, st);
    }
}

ProjK proj_k(const ProjArgs& a) {
    ProjK k;
    k.g = pix_geom(a.n, a.N, a.lo);
    k.batch = a.batch; k.nsrc = a.nsrc; k.hid = a.hid; k.out_ch = a.out_ch;
    k.ctot = 0;
    for (int s = 0; s < 4; ++s) {
        k.src[s] = s < a.nsrc ? a.src[s] : nullptr;
        k.gsrc[s] = s < a.nsrc ? a.gsrc[s] : nullptr;
        k.src_ch[s] = s < a.nsrc ? a.src_ch[s] : 0;
        k.ctot += k.src_ch[s];
    }
    k.w1 = a.w1; k.b1 = a.b1; k.w2 = a.w2; k.b2 = a.b2; k.out = a.out; k.gout = a.gout;
    k.pre_out = a.pre_out; k.pre_in = a.pre_in;
    k.gw1 = a.gw1; k.gb1 = a.gb1; k.gw2 = a.gw2; k.gb2 = a.gb2;
    return k;
}

template <int CT>
int launch_proj(const ProjK& k, bool bwd, cudaStream_t st) {
    if (!bwd && k.ctot > 32 && k.hid <= 64 && !cfg(CFG_PROJ_SIMT) && tc_enabled()) {
        // DEFAULT for wide inputs: tcgen05 forward (pixel_mlp_tc.cuh), 1.00 -> 0.74 ms per Darcy step.  The fp32 kernel below
        // for hid > 64, switch proj_simt, and for at most 32 input channels, where it measured faster (the 24 channels of
        // Uno3D_T10: 0.23 against 0.27 ms -- the tensor-core kernel stages 64-channel images whatever the real count)
        const int N_t = std::max(16, (k.hid + 15) & ~15);
        const size_t smem = proj_fwd_tc_smem(N_t, k.hid, k.out_ch);
        const long ntiles = ((long)k.batch * k.g.nraw + kPtPix - 1) / kPtPix;
        const unsigned grid = (unsigned)std::min<long>(ntiles, (long)num_sms());
        auto go = [&](auto kern) -> int {
            int rc = ensure_smem(kern, smem);
            if (rc) return rc;
            kern<<<grid, kPfThreads, smem, st>>>(k, ntiles, N_t);
            return 0;
        };
        int rc;
        if (k.out_ch == 1) rc = k.pre_out ? go(proj_fwd_tc_kernel<1, true>) : go(proj_fwd_tc_kernel<1, false>);
        else rc = k.pre_out ? go(proj_fwd_tc_kernel<4, true>) : go(proj_fwd_tc_kernel<4, false>);
        if (rc) return rc;
    } else if (!bwd) {
        const size_t smem = proj_fwd_smem(CT, k.hid, k.out_ch);
        int rc = ensure_smem(proj_fwd_kernel<CT>, smem);
        if (rc) return rc;
        const long ntiles = ((long)k.batch * k.g.nraw + kPixTP - 1) / kPixTP;
        const unsigned grid = (unsigned)std::min<long>(ntiles, 148L * 2);     // persistent, two CTAs per SM
        proj_fwd_kernel<CT><<<grid, kPixTP, smem, st>>>(k, ntiles);
    } else if (k.hid <= kProjHC && k.out_ch == 1 && k.pre_in != nullptr && !cfg(CFG_PROJ_SIMT) && tc_enabled()) {
        // (also for fewer than 33 channels -- the 24 of Uno3D_T10: the kernel pads to 64 internally; 0.91 -> 0.64 ms per NS-3D step)
        // DEFAULT for the shipped shapes (64 channels, hid <= 32, one output, pre-activations kept by forward):
        // warp-specialised tcgen05 kernel, 1.6 ms at Darcy size against 3.4 ms for the fp32 kernel (switch proj_simt)
        const size_t smem = proj_bwd_tcp_smem(k.hid, k.out_ch);
        int rc = ensure_smem(proj_bwd_tcp_kernel, smem);
        if (rc) return rc;
        const long ntiles = ((long)k.batch * k.g.nraw + kPtPix - 1) / kPtPix;
        const unsigned grid = (unsigned)std::min<long>(ntiles, 148L);
        proj_bwd_tcp_kernel<<<grid, kTcpThreads, smem, st>>>(k, ntiles);
    } else {
        // double-buffer the staged inputs when both copies fit beside the weights (hid <= 64 at 64 channels)
        const int nbuf = proj_bwd_smem(CT, k.hid, k.out_ch, 2) <= 224 * 1024 ? 2 : 1;
        const size_t smem = proj_bwd_smem(CT, k.hid, k.out_ch, nbuf);
        int rc = ensure_smem(proj_bwd_kernel<CT>, smem);
        if (rc) return rc;
        const long ntiles = ((long)k.batch * k.g.nraw + kPixTP - 1) / kPixTP;   // tiles of the cropped grid
        const unsigned grid = (unsigned)std::min<long>(ntiles, 148L);            // persistent, one CTA per SM
        proj_bwd_kernel<CT><<<grid, kPixTP, smem, st>>>(k, ntiles, nbuf);
    }
    CU_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int be_lift_supported(const LiftArgs& a) {
    const int cin = a.raw_ch + a.grid_ch;
    return a.raw_ch >= 1 && a.grid_ch >= 0 && cin <= 16 && a.hid >= 1 && a.hid <= 32 && a.out_ch >= 1 && a.out_ch <= 64;
}
int be_lift_fwd(const LiftArgs& a, stream_t s) {
    if (!be_lift_supported(a)) return (int)cudaErrorInvalidValue;
    const LiftK k = lift_k(a);
    const double px = (double)a.batch * k.g.nraw, ppx = (double)a.batch * k.g.npad;
    ProfScope ps("lift_fwd", 4.0 * (px * a.raw_ch + ppx * a.out_ch), 2.0 * px * (k.cin * a.hid + a.hid * a.out_ch), S(s));
    return dispatch_lift(k, false, S(s));
}
int be_lift_bwd(const LiftArgs& a, stream_t s) {
    if (!be_lift_supported(a)) return (int)cudaErrorInvalidValue;
    const LiftK k = lift_k(a);
    const double px = (double)a.batch * k.g.nraw;
    ProfScope ps("lift_bwd", 4.0 * px * (a.raw_ch * (a.ga ? 2 : 1) + a.out_ch), 2.0 * px * (2 * k.cin * a.hid + 3 * a.hid * a.out_ch), S(s));
    return dispatch_lift(k, true, S(s));
}

int be_proj_supported(const ProjArgs& a) {
    int ctot = 0;
    if (a.nsrc < 1 || a.nsrc > 4) return 0;
    for (int s = 0; s < a.nsrc; ++s) { if (a.src_ch[s] < 1) return 0; ctot += a.src_ch[s]; }
    return ctot <= 64 && a.hid >= 1 && a.hid <= 128 && a.out_ch >= 1 && a.out_ch <= kProjMaxOut;
}
int be_proj_fwd(const ProjArgs& a, stream_t s) {
    if (!be_proj_supported(a)) return (int)cudaErrorInvalidValue;
    const ProjK k = proj_k(a);
    const double px = (double)a.batch * k.g.nraw;
    ProfScope ps("project_fwd", 4.0 * px * (k.ctot + a.out_ch + (a.pre_out ? a.hid : 0)), 2.0 * px * (k.ctot * a.hid + a.hid * a.out_ch), S(s));
    return k.ctot <= 32 ? launch_proj<32>(k, false, S(s)) : launch_proj<64>(k, false, S(s));
}
int be_proj_bwd(const ProjArgs& a, stream_t s) {
    if (!be_proj_supported(a)) return (int)cudaErrorInvalidValue;
    const ProjK k = proj_k(a);
    const double px = (double)a.batch * k.g.nraw, ppx = (double)a.batch * k.g.npad;
    ProfScope ps("project_bwd", 4.0 * (px * (k.ctot + a.out_ch + (a.pre_in ? a.hid : 0)) + ppx * k.ctot),
                 2.0 * px * ((a.pre_in ? 2 : 3) * k.ctot * a.hid + 2 * a.hid * a.out_ch), S(s));
    return k.ctot <= 32 ? launch_proj<32>(k, true, S(s)) : launch_proj<64>(k, true, S(s));
}

// =====================================================================================================
// training-step ops (train_ops.cuh)
// =====================================================================================================
int be_adam_step(const AdamTensor* t, int n, const AdamHyper& h, stream_t s) {
    const double bc1 = 1.0 - pow(h.beta1, (double)h.step), bc2 = 1.0 - pow(h.beta2, (double)h.step);
    for (int i0 = 0; i0 < n; i0 += kAdamMaxTensors) {
        AdamBatch b;
        memset(&b, 0, sizeof b);
        b.n = std::min(kAdamMaxTensors, n - i0);
        double floats = 0;
        int chunks = 0;
        for (int i = 0; i < b.n; ++i) {
            const AdamTensor& a = t[i0 + i];
            b.param[i] = a.param; b.grad[i] = a.grad; b.m[i] = a.exp_avg; b.v[i] = a.exp_avg_sq; b.vmax[i] = a.max_exp_avg_sq;
            b.numel[i] = a.numel; b.is_complex[i] = a.is_complex;
            b.chunk_start[i] = chunks;
            chunks += (int)((a.numel + kAdamChunk - 1) / kAdamChunk);
            floats += (double)a.numel;
        }
        b.chunk_start[b.n] = chunks;
        b.beta1 = (float)h.beta1; b.one_minus_beta1 = (float)(1.0 - h.beta1);
        b.beta2 = (float)h.beta2; b.one_minus_beta2 = (float)(1.0 - h.beta2);
        b.eps = (float)h.eps; b.weight_decay = (float)h.weight_decay;
        b.step_size = (float)(h.lr / bc1);
        b.sqrt_bc2 = (float)sqrt(bc2);
        b.amsgrad = h.amsgrad;
        if (chunks == 0) continue;
        ProfScope ps("adam", 4.0 * floats * (h.amsgrad ? 8 : 6), 0, S(s));
        adam_kernel<<<(unsigned)chunks, 256, 0, S(s)>>>(b);
        CU_LAUNCH_CHECK();
    }
    return 0;
}

int be_lp_loss_fwd(const float* x, const float* y, int B, long N, int reduction, float* loss, float* norms, double* acc, stream_t s) {
    if (B <= 0 || N <= 0) return 0;
    cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double) * 2 * B, S(s));
    if (e != cudaSuccess) return (int)e;
    const unsigned gy = (unsigned)std::max<long>(1, std::min<long>((N + 8191) / 8192, (148L * 8 + B - 1) / B));
    ProfScope ps("lp_loss", 8.0 * B * N, 0, S(s));
    lp_partial_kernel<<<dim3((unsigned)B, gy), 256, 0, S(s)>>>(x, y, N, acc);
    lp_finish_kernel<<<1, 256, 0, S(s)>>>(acc, B, reduction, norms, loss);
    CU_LAUNCH_CHECK();
    return 0;
}
int be_lp_loss_bwd(const float* x, const float* y, const float* norms, const float* gl, int B, long N, int reduction, float* gx,
                   stream_t s) {
    if (B <= 0 || N <= 0) return 0;
    const unsigned gy = (unsigned)std::max<long>(1, std::min<long>((N + 8191) / 8192, (148L * 8 + B - 1) / B));
    ProfScope ps("lp_loss_bwd", 12.0 * B * N, 0, S(s));
    lp_bwd_kernel<<<dim3((unsigned)B, gy), 256, 0, S(s)>>>(x, y, norms, gl, N, B, reduction, gx);
    CU_LAUNCH_CHECK();
    return 0;
}

}  // namespace uno
