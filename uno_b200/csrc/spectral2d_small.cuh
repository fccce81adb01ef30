// Truncated 2-D DFT analysis / synthesis of WHOLE PLANES in one kernel, for grids small enough that a few planes, both transform
// matrices and the intermediate fit in shared memory (H, W <= 64: every level of the Navier-Stokes 2-D U-NO).  Included by
// backend_cuda.cu inside namespace uno::{anonymous}.
//
//   analysis   X^[p, j, k] = sum_h A[j, h] * ( sum_w x[p, h, w] * F[w, (k, re|im)] )          x real, A complex, X^ complex
//   synthesis  y[p, h', w'] (+)= sum_c ( sum_j S[h', j] * Y^[p, j, k] )_(k, re|im) * E[c, w']   + the block epilogue
//
// The two-launch path (last-axis GEMM + leading-axis transform, each a persistent tcgen05 kernel) spends most of its 25-60 us per
// level at these sizes on fixed costs -- launch, TMEM allocation, pipeline fill, a round trip of the intermediate through L2 --
// not on the few MFLOP a level needs.  Here a CTA takes a group of PP planes (PP * H <= 128 rows), runs both stages out of
// shared memory with a register-tiled fp32 SIMT product (no tf32 split needed) and never writes the intermediate.
//
// Complex products are real ones: with T [H x 2m] holding (re, im) pairs and T' its rotation (-im, re),
//   X^ (as [J x 2m] real) = [Are | Aim] [J x 2H] * [T ; T'] [2H x 2m],       and the same with S and Y^ for the synthesis.
#pragma once

struct Spec2Small {
    // shared
    int P;                 // planes
    int H, W;              // spatial grid of a plane (analysis: input, synthesis: output)
    int J, m;              // kept rows (2 * modes1) and complex columns (modes2) of a plane's spectrum
    int PP;                // planes per group
    int ldn;               // padded row length of the [.. x 2m] matrices in shared memory (multiple of 4)
    int ldw;               // padded row length of the [.. x W] matrices (multiple of 4)
    const float* last;     // analysis: F [W x 2m]; synthesis: E [2m x W]   (real, row-major)
    const float* mid;      // analysis: A [J x H]; synthesis: S [H x J]     (complex interleaved, row-major)
    // analysis
    const float* x; float* xhat;
    // synthesis
    const float* yhat; float* y; float* y2; int epi;
};

// C[M x N] (ldc) = A[M x K] (lda) * B[K x N] (ldb), all in shared memory, N <= 64, ldb/ldc multiples of 4 with zero padding up to
// the next multiple of 4 of N.  256 threads: tx = tid % nx owns columns 4 tx .. 4 tx + 3, ty = tid / nx owns rows ty + ny * i; every
// 16-byte load of A (4 k of a row, broadcast over tx) and of B (4 columns of a k, contiguous over tx) feeds 16 products.
// `emit(row, col4, acc4)` receives each finished 1 x 4 strip.  K is rounded up to a multiple of 4 by the caller (zero padded).
template <typename Emit>
__device__ __forceinline__ void smem_gemm4(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int M, int N, int K,
                                           Emit emit) {
    const int nx = (N + 3) >> 2;
    const int ny = 256 / nx;
    const int tid = threadIdx.x;
    const int ty = tid / nx, tx = tid - ty * nx;
    if (ty >= ny) return;
    const float* bp = B + 4 * tx;
    for (int r0 = ty; r0 < M; r0 += 4 * ny) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        const float* ap[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ap[i] = A + (size_t)min(r0 + i * ny, M - 1) * lda;   // rows past M recompute the last row, not emitted
#pragma unroll 2
        for (int k = 0; k < K; k += 4) {
            float4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(ap[i] + k);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) b[kk] = *reinterpret_cast<const float4*>(bp + (size_t)(k + kk) * ldb);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(a[i].x, b[0].x, acc[i][0]); acc[i][1] = fmaf(a[i].x, b[0].y, acc[i][1]);
                acc[i][2] = fmaf(a[i].x, b[0].z, acc[i][2]); acc[i][3] = fmaf(a[i].x, b[0].w, acc[i][3]);
                acc[i][0] = fmaf(a[i].y, b[1].x, acc[i][0]); acc[i][1] = fmaf(a[i].y, b[1].y, acc[i][1]);
                acc[i][2] = fmaf(a[i].y, b[1].z, acc[i][2]); acc[i][3] = fmaf(a[i].y, b[1].w, acc[i][3]);
                acc[i][0] = fmaf(a[i].z, b[2].x, acc[i][0]); acc[i][1] = fmaf(a[i].z, b[2].y, acc[i][1]);
                acc[i][2] = fmaf(a[i].z, b[2].z, acc[i][2]); acc[i][3] = fmaf(a[i].z, b[2].w, acc[i][3]);
                acc[i][0] = fmaf(a[i].w, b[3].x, acc[i][0]); acc[i][1] = fmaf(a[i].w, b[3].y, acc[i][1]);
                acc[i][2] = fmaf(a[i].w, b[3].z, acc[i][2]); acc[i][3] = fmaf(a[i].w, b[3].w, acc[i][3]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + i * ny;
            if (r < M) emit(r, 4 * tx, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        }
    }
}

__host__ __device__ inline int s2_round4(int n) { return (n + 3) & ~3; }

// shared-memory layout (floats); K dimensions are padded to multiples of 4 with zeros
struct Spec2Layout {
    int kW, kH2, kJ2, kM2;          // padded contraction lengths: W, 2H, 2J, 2m
    int ld_in, ld_mid, ld_z;        // row pitches of the A operands: contraction length + 4, so that the 16-byte loads of the (up to
                                    // eight) different rows a warp reads at once fall in different bank groups
    size_t o_last, o_mid, o_in, o_t, o_out, total;
};
__host__ __device__ inline Spec2Layout spec2_layout_analysis(int H, int W, int J, int m, int PP) {
    Spec2Layout L;
    L.kW = s2_round4(W); L.kH2 = s2_round4(2 * H); L.kJ2 = 0; L.kM2 = 0;
    L.ld_in = L.kW + 4; L.ld_mid = L.kH2 + 4; L.ld_z = 0;
    const int ldn = s2_round4(2 * m);
    size_t o = 0;
    L.o_last = o; o += (size_t)L.kW * ldn;                  // F  [kW x ldn]
    L.o_mid = o; o += (size_t)J * L.ld_mid;                 // A' [J x kH2] = [Are | Aim]
    L.o_in = o; o += (size_t)PP * H * L.ld_in;              // x  [PP*H x kW]
    L.o_t = o; o += (size_t)PP * L.kH2 * ldn;               // per plane [T ; T'] [kH2 x ldn]
    L.o_out = L.o_in;                                       // X^ of the group reuses the input buffer: [PP][J x ldn]
    if ((size_t)PP * J * ldn > (size_t)PP * H * L.ld_in) { L.o_out = o; o += (size_t)PP * J * ldn; }
    L.total = o;
    return L;
}
__host__ __device__ inline Spec2Layout spec2_layout_synthesis(int H, int W, int J, int m, int PP) {
    Spec2Layout L;
    L.kW = 0; L.kH2 = 0; L.kJ2 = s2_round4(2 * J); L.kM2 = s2_round4(2 * m);
    L.ld_in = 0; L.ld_mid = L.kJ2 + 4; L.ld_z = L.kM2 + 4;
    const int ldn = s2_round4(2 * m), ldw = s2_round4(W);
    size_t o = 0;
    L.o_last = o; o += (size_t)L.kM2 * ldw;                 // E  [kM2 x ldw]
    L.o_mid = o; o += (size_t)H * L.ld_mid;                 // S' [H x kJ2] = [Sre | Sim]
    L.o_in = o; o += (size_t)PP * L.kJ2 * ldn;              // per plane [Y^ ; Y^'] [kJ2 x ldn]
    L.o_t = o; o += (size_t)PP * H * L.ld_z;                // Z  [PP*H x kM2]
    L.o_out = 0;
    L.total = o;
    return L;
}

__global__ void __launch_bounds__(256, 2) analysis2d_small_kernel(const Spec2Small k) {
    extern __shared__ __align__(16) float s2m[];
    const Spec2Layout L = spec2_layout_analysis(k.H, k.W, k.J, k.m, k.PP);
    float* sF = s2m + L.o_last;
    float* sA = s2m + L.o_mid;
    float* sX = s2m + L.o_in;
    float* sT = s2m + L.o_t;
    float* sO = s2m + L.o_out;
    const int tid = threadIdx.x;
    const int N = 2 * k.m, ldn = k.ldn;
    // constants: F zero-padded to [kW x ldn]; A' = [Are | Aim] zero-padded to [J x kH2]
    for (int i = tid; i < L.kW * ldn; i += 256) {
        const int w = i / ldn, c = i - w * ldn;
        sF[i] = (w < k.W && c < N) ? __ldg(k.last + (size_t)w * N + c) : 0.f;
    }
    for (int i = tid; i < k.J * L.kH2; i += 256) {
        const int j = i / L.kH2, c = i - j * L.kH2;
        float v = 0.f;
        if (c < k.H) v = __ldg(k.mid + ((size_t)j * k.H + c) * 2);
        else if (c < 2 * k.H) v = __ldg(k.mid + ((size_t)j * k.H + (c - k.H)) * 2 + 1);
        sA[(size_t)j * L.ld_mid + c] = v;
    }
    // rows of T / T' past 2H (contraction padding) stay zero for the whole kernel
    for (int i = tid; i < k.PP * L.kH2 * ldn; i += 256) sT[i] = 0.f;
    __syncthreads();
    const long groups = ((long)k.P + k.PP - 1) / k.PP;
    const int HW = k.H * k.W;
    const bool vec = (k.W & 3) == 0 && (HW & 3) == 0 && (reinterpret_cast<uintptr_t>(k.x) & 15) == 0;
    for (long g = blockIdx.x; g < groups; g += gridDim.x) {
        const long p0 = g * k.PP;
        const int np = (int)min((long)k.PP, k.P - p0);
        const int M = np * k.H;
        // ---- stage the planes of the group: rows of W floats into rows of kW (zero padded)
        {
            const float* src = k.x + p0 * HW;
            if (vec) {
                const int w4 = k.W >> 2;
                for (int i = tid; i < M * w4; i += 256) {
                    const int r = i / w4, c = i - r * w4;
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sX + (size_t)r * L.ld_in + 4 * c);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + (size_t)r * k.W + 4 * c) : "memory");
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
            } else {
                for (int i = tid; i < M * L.kW; i += 256) {
                    const int r = i / L.kW, c = i - r * L.kW;
                    sX[(size_t)r * L.ld_in + c] = c < k.W ? __ldg(src + (size_t)r * k.W + c) : 0.f;
                }
            }
        }
        __syncthreads();
        // ---- stage 1: T[(p,h), c] = sum_w x[(p,h), w] F[w, c]; written as [T ; T'] per plane
        smem_gemm4(sX, L.ld_in, sF, ldn, M, N, L.kW, [&](int r, int c4, float4 v) {
            const int p = r / k.H, h = r - p * k.H;
            float* t = sT + ((size_t)p * L.kH2 + h) * ldn + c4;
            *reinterpret_cast<float4*>(t) = v;                                                     // (re0, im0, re1, im1)
            *reinterpret_cast<float4*>(t + (size_t)k.H * ldn) = make_float4(-v.y, v.x, -v.w, v.z);    // T' = i * T
        });
        __syncthreads();
        // ---- stage 2, plane by plane: X^[j, c] = sum_{h'} A'[j, h'] [T ; T'][h', c]
        for (int p = 0; p < np; ++p) {
            smem_gemm4(sA, L.ld_mid, sT + (size_t)p * L.kH2 * ldn, ldn, k.J, N, L.kH2, [&](int r, int c4, float4 v) {
                *reinterpret_cast<float4*>(sO + ((size_t)p * k.J + r) * ldn + c4) = v;
            });
        }
        __syncthreads();
        // ---- store the group's spectra: [np * J] rows of N floats, contiguous in global memory
        {
            float* dst = k.xhat + p0 * (size_t)k.J * N;
            const int rows = np * k.J;
            if ((N & 3) == 0 && (reinterpret_cast<uintptr_t>(k.xhat) & 15) == 0) {
                const int n4 = N >> 2;
                for (int i = tid; i < rows * n4; i += 256) {
                    const int r = i / n4, c = i - r * n4;
                    *reinterpret_cast<float4*>(dst + (size_t)r * N + 4 * c) = *reinterpret_cast<const float4*>(sO + (size_t)r * ldn + 4 * c);
                }
            } else {
                for (int i = tid; i < rows * N; i += 256) {
                    const int r = i / N, c = i - r * N;
                    dst[i] = sO[(size_t)r * ldn + c];
                }
            }
        }
        __syncthreads();      // sX / sO and sT are rewritten by the next group
    }
}

__global__ void __launch_bounds__(256, 2) synthesis2d_small_kernel(const Spec2Small k) {
    extern __shared__ __align__(16) float s2m[];
    const Spec2Layout L = spec2_layout_synthesis(k.H, k.W, k.J, k.m, k.PP);
    float* sE = s2m + L.o_last;
    float* sS = s2m + L.o_mid;
    float* sY = s2m + L.o_in;
    float* sZ = s2m + L.o_t;
    const int tid = threadIdx.x;
    const int N = 2 * k.m, ldn = k.ldn, ldw = k.ldw;
    for (int i = tid; i < L.kM2 * ldw; i += 256) {
        const int c = i / ldw, w = i - c * ldw;
        sE[i] = (c < N && w < k.W) ? __ldg(k.last + (size_t)c * k.W + w) : 0.f;
    }
    for (int i = tid; i < k.H * L.kJ2; i += 256) {
        const int h = i / L.kJ2, c = i - h * L.kJ2;
        float v = 0.f;
        if (c < k.J) v = __ldg(k.mid + ((size_t)h * k.J + c) * 2);
        else if (c < 2 * k.J) v = __ldg(k.mid + ((size_t)h * k.J + (c - k.J)) * 2 + 1);
        sS[(size_t)h * L.ld_mid + c] = v;
    }
    for (int i = tid; i < k.PP * L.kJ2 * ldn; i += 256) sY[i] = 0.f;      // contraction padding rows / columns stay zero
    for (int i = tid; i < k.PP * k.H * L.ld_z; i += 256) sZ[i] = 0.f;
    __syncthreads();
    const long groups = ((long)k.P + k.PP - 1) / k.PP;
    const long HW = (long)k.H * k.W;
    const bool vec_out = (k.W & 3) == 0 && (reinterpret_cast<uintptr_t>(k.y) & 15) == 0 && (k.y2 == nullptr || (reinterpret_cast<uintptr_t>(k.y2) & 15) == 0);
    for (long g = blockIdx.x; g < groups; g += gridDim.x) {
        const long p0 = g * k.PP;
        const int np = (int)min((long)k.PP, k.P - p0);
        // ---- stage the spectra of the group as [Y^ ; Y^'] per plane
        {
            const float* src = k.yhat + p0 * (size_t)k.J * N;
            const int n2 = N >> 1;       // complex columns
            for (int i = tid; i < np * k.J * n2; i += 256) {
                const int r = i / n2, c = i - r * n2;          // r = (p, j)
                const int p = r / k.J, j = r - p * k.J;
                const float2 v = __ldg(reinterpret_cast<const float2*>(src) + i);
                float* d = sY + ((size_t)p * L.kJ2 + j) * ldn + 2 * c;
                *reinterpret_cast<float2*>(d) = v;
                *reinterpret_cast<float2*>(d + (size_t)k.J * ldn) = make_float2(-v.y, v.x);
            }
        }
        __syncthreads();
        // ---- stage 1, plane by plane: Z[h, c] = sum_{j'} S'[h, j'] [Y^ ; Y^'][j', c]
        for (int p = 0; p < np; ++p) {
            smem_gemm4(sS, L.ld_mid, sY + (size_t)p * L.kJ2 * ldn, ldn, k.H, N, L.kJ2, [&](int r, int c4, float4 v) {
                float* z = sZ + ((size_t)p * k.H + r) * L.ld_z + c4;
                // columns past N are contraction padding: keep them zero
                if (c4 + 3 < N) *reinterpret_cast<float4*>(z) = v;
                else {
                    if (c4 + 0 < N) z[0] = v.x;
                    if (c4 + 1 < N) z[1] = v.y;
                    if (c4 + 2 < N) z[2] = v.z;
                }
            });
        }
        __syncthreads();
        // ---- stage 2: y[(p,h), w] = sum_c Z[(p,h), c] E[c, w], with the block epilogue, straight to global memory
        {
            float* yp = k.y + p0 * HW;
            float* y2p = k.y2 ? k.y2 + p0 * HW : nullptr;
            const int epi = k.epi;
            smem_gemm4(sZ, L.ld_z, sE, ldw, np * k.H, k.W, L.kM2, [&](int r, int c4, float4 v) {
                float* q = yp + (size_t)r * k.W + c4;
                const int left = k.W - c4;
                if (vec_out && left >= 4) {
                    float4 o = v;
                    if (epi != EPI_STORE) {
                        const float4 a = *reinterpret_cast<const float4*>(q);
                        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
                    }
                    if (epi == EPI_ACCUM_GELU) {
                        *reinterpret_cast<float4*>(q) = o;
                        *reinterpret_cast<float4*>(y2p + (size_t)r * k.W + c4) = make_float4(gelu_f(o.x), gelu_f(o.y), gelu_f(o.z), gelu_f(o.w));
                    } else if (epi == EPI_ACCUM_GELU_INPLACE) {
                        *reinterpret_cast<float4*>(q) = make_float4(gelu_f(o.x), gelu_f(o.y), gelu_f(o.z), gelu_f(o.w));
                    } else {
                        *reinterpret_cast<float4*>(q) = o;
                    }
                } else {
                    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (e < left) {
                            float o = vv[e];
                            if (epi != EPI_STORE) o += q[e];
                            if (epi == EPI_ACCUM_GELU) { q[e] = o; y2p[(size_t)r * k.W + c4 + e] = gelu_f(o); }
                            else if (epi == EPI_ACCUM_GELU_INPLACE) q[e] = gelu_f(o);
                            else q[e] = o;
                        }
                    }
                }
            });
        }
        __syncthreads();
    }
}
