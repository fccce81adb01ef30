// Fused separable 2-D banded resample, register-blocked (pointwise_op_2D's anti-aliased bicubic and its transpose;
// integral_operators.py:240-242).  Included by backend_cuda.cu inside namespace uno::{anonymous}.
//
//   y[p] = R0 * x[p] * R1^T  for every plane p,     x [P, n_in0, n_in1] -> y [P, n_out0, n_out1]
//
// Both bands come as "group images" (plan.h BandGroups): G consecutive outputs share a window of W inputs and a
// dense G x W weight block.  One CTA = one plane x (TH x 64) output tile:
//   1. the input window [rin x cin] is staged in shared memory with 4-byte LDGSTS (rows of 481 floats are not
//      16-byte aligned), every element of x is read once per tile;
//   2. pass A (contiguous axis): a warp owns one column group (its G1 x W1 weights sit in registers), lanes run over
//      the window rows -> odd row pitch makes every shared access conflict-free; W1 loads feed G1*W1 FMAs;
//   3. pass B (row axis): a warp owns one row group (G0 x W0 weights in registers), lanes run over 32 output
//      columns; W0 loads feed G0*W0 FMAs and G0 coalesced 128-byte row stores.
// Bound: HBM (4 B read per input element + 4 B written per output element); the FMA work is
// planes * (n_in0*n_out1*W1 + n_out0*n_out1*W0), 25-40 flop per byte of traffic for the 2x resamples of the U-NO levels.
#pragma once

struct Resample2K {
    const float* x; float* y; long planes;
    int n_in0, n_out0, n_in1, n_out1;
    const int* gs0; const float* D0; int ng0;   // row band groups
    const int* gs1; const float* D1; int ng1;   // column band groups
    int TH;                                     // output rows per tile (multiple of G0); 64 output columns per tile
    int RIN, ldin;                              // rows of the staged window, its (odd) pitch
    int tiles_h, tiles_w;
};
// output columns per tile: 128 for the small-window (G = 8, up-sampling) image, whose tiles are otherwise dominated by
// per-tile overheads, 64 for the others
__host__ __device__ constexpr int rs_tile_w(int G1) { return G1 == 8 ? 128 : 64; }

// (A variant that staged the window as two cp.async groups so that pass A could start on the upper half measured SLOWER on
// B200 -- 3.98 vs 3.43 ms per Darcy step -- and was removed: the second resident CTA already covers the staging latency.)
template <int G0, int W0, int G1, int W1, int MINB>
__global__ void __launch_bounds__(256, MINB) resample2d_kernel(const Resample2K k) {
    extern __shared__ __align__(16) float rsm[];
    constexpr int kRsTW = rs_tile_w(G1), kRsMidLd = kRsTW + 1;
    constexpr int NG1 = kRsTW / G1;                      // column groups per tile
    const int NG0 = k.TH / G0;                           // row groups per tile
    float* in_s = rsm;                                   // [RIN][ldin]
    float* mid_s = in_s + (size_t)k.RIN * k.ldin;        // [RIN][kRsMidLd]
    float* d1s = mid_s + (size_t)k.RIN * kRsMidLd;       // [NG1][W1][G1]   (16-byte aligned by construction of the sizes below)
    float* d0s = d1s + NG1 * W1 * G1;                    // [NG0][W0][G0]
    int* gs1s = reinterpret_cast<int*>(d0s + NG0 * W0 * G0);   // [NG1]
    int* gs0s = gs1s + NG1;                              // [NG0]
    const int tw = blockIdx.x, th = blockIdx.y;          // grid = (tiles_w, tiles_h, planes)
    const long p = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ga0 = th * NG0, ga1 = tw * NG1;            // first row / column group of the tile
    const int ngh = min(NG0, k.ng0 - ga0), ngw = min(NG1, k.ng1 - ga1);
    const int r0 = __ldg(k.gs0 + ga0), c0 = __ldg(k.gs1 + ga1);
    const int rin = __ldg(k.gs0 + ga0 + ngh - 1) + W0 - r0;   // windows never leave the input (plan guarantee)
    const int cin = __ldg(k.gs1 + ga1 + ngw - 1) + W1 - c0;
    // ---- stage the input window: a warp per row, lanes over columns (up to 160 columns unrolled, predicated)
    {
        const float* xp = k.x + (p * k.n_in0 + r0) * (long)k.n_in1 + c0 + lane;
        const uint32_t in_base = (uint32_t)__cvta_generic_to_shared(in_s) + 4u * lane;
        for (int r = warp; r < rin; r += 8) {
            const float* src = xp + (long)r * k.n_in1;
            const uint32_t dst = in_base + 4u * (uint32_t)(r * k.ldin);
            constexpr int CIT = W1 == 16 ? 5 : 3;       // unrolled, predicated column chunks (wider windows: loop below)
#pragma unroll
            for (int it = 0; it < CIT; ++it)
                if (lane + 32 * it < cin)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 128u * it), "l"(src + 32 * it) : "memory");
            for (int c = lane + 32 * CIT; c < cin; c += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * (c - lane)), "l"(src + (c - lane)) : "memory");
        }
    }
    for (int i = tid; i < ngw * W1 * G1; i += 256) d1s[i] = __ldg(k.D1 + (size_t)ga1 * W1 * G1 + i);
    for (int i = tid; i < ngh * W0 * G0; i += 256) d0s[i] = __ldg(k.D0 + (size_t)ga0 * W0 * G0 + i);
    if (tid < ngw) gs1s[tid] = __ldg(k.gs1 + ga1 + tid) - c0;
    if (tid >= 64 && tid < 64 + ngh) gs0s[tid - 64] = __ldg(k.gs0 + ga0 + tid - 64) - r0;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // ---- pass A: mid[r][j] = sum_u D1[g][u][q] * in[r][gs1[g] + u],  j = g*G1 + q
    for (int cg = warp; cg < ngw; cg += 8) {
        float w[W1][G1];
        const float4* wsrc = reinterpret_cast<const float4*>(d1s + cg * W1 * G1);
#pragma unroll
        for (int u = 0; u < W1; ++u)
#pragma unroll
            for (int q4 = 0; q4 < G1 / 4; ++q4) {
                const float4 v = wsrc[u * (G1 / 4) + q4];
                w[u][4 * q4 + 0] = v.x; w[u][4 * q4 + 1] = v.y; w[u][4 * q4 + 2] = v.z; w[u][4 * q4 + 3] = v.w;
            }
        const int cs = gs1s[cg];
        for (int rb = 0; rb < rin; rb += 32) {
            // uniform control flow: lanes past the window recompute its last row and skip the store
            const int r = min(rb + lane, rin - 1);
            const float* src = in_s + r * k.ldin + cs;
            float acc[G1];
#pragma unroll
            for (int q = 0; q < G1; ++q) acc[q] = 0.f;
#pragma unroll
            for (int u = 0; u < W1; ++u) {
                const float v = src[u];
#pragma unroll
                for (int q = 0; q < G1; ++q) acc[q] = fmaf(w[u][q], v, acc[q]);
            }
            if (rb + lane < rin) {
                float* dst = mid_s + r * kRsMidLd + cg * G1;
#pragma unroll
                for (int q = 0; q < G1; ++q) dst[q] = acc[q];
            }
        }
    }
    __syncthreads();
    // ---- pass B: y[i][j] = sum_u D0[g][u][q] * mid[gs0[g] + u][j],  i = g*G0 + q
    const int i0 = ga0 * G0, j0 = tw * kRsTW;
    for (int rg = warp; rg < ngh; rg += 8) {
        float w[W0][G0];
        const float4* wsrc = reinterpret_cast<const float4*>(d0s + rg * W0 * G0);
#pragma unroll
        for (int u = 0; u < W0; ++u)
#pragma unroll
            for (int q4 = 0; q4 < G0 / 4; ++q4) {
                const float4 v = wsrc[u * (G0 / 4) + q4];
                w[u][4 * q4 + 0] = v.x; w[u][4 * q4 + 1] = v.y; w[u][4 * q4 + 2] = v.z; w[u][4 * q4 + 3] = v.w;
            }
        const float* srow = mid_s + gs0s[rg] * kRsMidLd + lane;
        float* yrow = k.y + (p * k.n_out0 + i0 + rg * G0) * (long)k.n_out1 + j0 + lane;
        const int rows_left = k.n_out0 - (i0 + rg * G0);
#pragma unroll 1
        for (int half = 0; half < kRsTW / 32; ++half) { // the 32-column slices of the tile share the weights
            const float* src = srow + 32 * half;
            float acc[G0];
#pragma unroll
            for (int q = 0; q < G0; ++q) acc[q] = 0.f;
#pragma unroll
            for (int u = 0; u < W0; ++u) {
                const float v = src[u * kRsMidLd];
#pragma unroll
                for (int q = 0; q < G0; ++q) acc[q] = fmaf(w[u][q], v, acc[q]);
            }
            const bool col_ok = j0 + 32 * half + lane < k.n_out1;
            float* yp = yrow + 32 * half;
#pragma unroll
            for (int q = 0; q < G0; ++q) {              // running row pointer: one 64-bit add per store, predicated stores
                if (col_ok && q < rows_left) *yp = acc[q];
                yp += k.n_out1;
            }
        }
    }
}

inline size_t resample2d_smem(int RIN, int ldin, int TH, int G0, int W0, int G1, int W1) {
    const int TW = rs_tile_w(G1);
    return sizeof(float) * ((size_t)RIN * ldin + (size_t)RIN * (TW + 1) + (size_t)(TW / G1) * W1 * G1 + (size_t)(TH / G0) * W0 * G0 +
                            TW / G1 + TH / G0);
}

// (A persistent, software-pipelined form of this kernel -- the LDGSTS copies of tile t+1 in flight while tile t runs its two
// passes -- was measured on B200 at 3.36 against 3.43 ms per Darcy step and removed: the kernel is bound by its shared-memory
// and FMA instruction stream, not by exposed load latency.)
