// Training-step ops next to the hot path: the reference's Adam (Adam.py:8-52) as ONE
// multi-tensor kernel, and the relative L2 loss (utilities3.py:86-100, LpLoss.rel) forward / backward.
// Included by backend_cuda.cu inside namespace uno::{anonymous}.
//
// Adam.py differs from torch.optim.Adam on complex parameters: the second moment is the EMA of g*conj(g) = |g|^2 (a real
// number kept in a complex tensor with zero imaginary part), so real and imaginary parts of a weight share one
// denominator.  The kernel keeps that: a complex element is an (re, im) pair of floats and exp_avg_sq stores (v, 0).
// The reference runs ~10 elementwise launches per parameter tensor from a Python loop; this is one launch per 24
// tensors, every value read once and written once (HBM-bound: 4 B x (3 reads + 3 writes) per float).
#pragma once

constexpr int kAdamMaxTensors = 24;
constexpr int kAdamChunk = 8192;            // floats per CTA

struct AdamBatch {
    int n;
    int chunk_start[kAdamMaxTensors + 1];
    float* param[kAdamMaxTensors];
    const float* grad[kAdamMaxTensors];
    float* m[kAdamMaxTensors];
    float* v[kAdamMaxTensors];
    float* vmax[kAdamMaxTensors];
    long numel[kAdamMaxTensors];
    int is_complex[kAdamMaxTensors];
    float beta1, one_minus_beta1, beta2, one_minus_beta2, eps, weight_decay, step_size, sqrt_bc2;
    int amsgrad;
};

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamBatch b) {
    int t = 0;
    while (t + 1 < b.n && (int)blockIdx.x >= b.chunk_start[t + 1]) ++t;
    const long base = (long)(blockIdx.x - b.chunk_start[t]) * kAdamChunk;
    const long n = b.numel[t];
    float* __restrict__ p = b.param[t];
    const float* __restrict__ g = b.grad[t];
    float* __restrict__ m = b.m[t];
    float* __restrict__ v = b.v[t];
    float* __restrict__ vm = b.vmax[t];
    const bool cx = b.is_complex[t] != 0;
    // a thread owns float pairs (2k, 2k+1): one complex element, or two independent reals
    for (long i = base + 2 * threadIdx.x; i < min(base + kAdamChunk, n); i += 2 * 256) {
        const bool has1 = i + 1 < n;
        float p0 = p[i], p1 = has1 ? p[i + 1] : 0.f;
        float g0 = g[i], g1 = has1 ? g[i + 1] : 0.f;
        if (b.weight_decay != 0.f) { g0 = fmaf(b.weight_decay, p0, g0); g1 = fmaf(b.weight_decay, p1, g1); }
        const float m0 = fmaf(b.one_minus_beta1, g0, m[i] * b.beta1);
        const float m1 = has1 ? fmaf(b.one_minus_beta1, g1, m[i + 1] * b.beta1) : 0.f;
        float v0, v1, d0, d1;
        if (cx) {
            v0 = fmaf(b.one_minus_beta2, fmaf(g0, g0, g1 * g1), v[i] * b.beta2);
            v1 = v[i + 1] * b.beta2;                                  // imaginary part of g*conj(g) is exactly zero
            d0 = d1 = sqrtf(v0) / b.sqrt_bc2 + b.eps;
        } else {
            v0 = fmaf(b.one_minus_beta2, g0 * g0, v[i] * b.beta2);
            v1 = has1 ? fmaf(b.one_minus_beta2, g1 * g1, v[i + 1] * b.beta2) : 0.f;
            float h0 = v0, h1 = v1;
            if (b.amsgrad) {
                h0 = fmaxf(vm[i], v0);
                vm[i] = h0;
                if (has1) { h1 = fmaxf(vm[i + 1], v1); vm[i + 1] = h1; }
            }
            d0 = sqrtf(h0) / b.sqrt_bc2 + b.eps;
            d1 = sqrtf(h1) / b.sqrt_bc2 + b.eps;
        }
        m[i] = m0; v[i] = v0;
        p[i] = fmaf(-b.step_size, m0 / d0, p0);
        if (has1) {
            m[i + 1] = m1; v[i + 1] = v1;
            p[i + 1] = fmaf(-b.step_size, m1 / d1, p1);
        }
    }
}

// ---- relative L2 loss ------------------------------------------------------------------------------------------
// partial sums of (x-y)^2 and y^2 per sample, double accumulators acc[b][2] (pre-zeroed)
__global__ void __launch_bounds__(256) lp_partial_kernel(const float* __restrict__ x, const float* __restrict__ y, long N,
                                                         double* __restrict__ acc) {
    __shared__ double sh[32];
    const long b = blockIdx.x;
    const float* xp = x + b * N;
    const float* yp = y + b * N;
    float sd = 0.f, sy = 0.f;
    double dd = 0.0, dy = 0.0;
    int cnt = 0;
    for (long i = (long)blockIdx.y * 256 + threadIdx.x; i < N; i += (long)gridDim.y * 256) {
        const float yv = yp[i], d = xp[i] - yv;
        sd = fmaf(d, d, sd);
        sy = fmaf(yv, yv, sy);
        if (++cnt == 32) { dd += sd; dy += sy; sd = sy = 0.f; cnt = 0; }
    }
    dd += sd; dy += sy;
    const double td = block_sum(dd, sh);
    const double ty = block_sum(dy, sh);
    if (threadIdx.x == 0) {
        atomicAdd(acc + 2 * b, td);
        atomicAdd(acc + 2 * b + 1, ty);
    }
}

// norms[b] = (||x-y||, ||y||); loss = per-sample ratios (reduction 0), their sum (1) or mean (2)
__global__ void lp_finish_kernel(const double* __restrict__ acc, int B, int reduction, float* __restrict__ norms, float* __restrict__ loss) {
    __shared__ double sh[32];
    double s = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float dn = (float)sqrt(acc[2 * b]), yn = (float)sqrt(acc[2 * b + 1]);
        norms[2 * b] = dn;
        norms[2 * b + 1] = yn;
        const float r = dn / yn;
        if (reduction == 0) loss[b] = r;
        s += (double)r;
    }
    const double t = block_sum(s, sh);
    if (threadIdx.x == 0 && reduction != 0) loss[0] = (float)(reduction == 2 ? t / B : t);
}

// gx = gl_b * (x - y) / (||x-y|| * ||y||),  gl_b = gl[b] (reduction 0), gl[0] (sum) or gl[0] / B (mean)
__global__ void __launch_bounds__(256) lp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ norms,
                                                     const float* __restrict__ gl, long N, int B, int reduction, float* __restrict__ gx) {
    const long b = blockIdx.x;
    const float dn = norms[2 * b], yn = norms[2 * b + 1];
    float scale = reduction == 0 ? gl[b] : gl[0];
    if (reduction == 2) scale /= (float)B;
    // d||d||/dd = d / ||d||; torch's norm backward masks ||d|| == 0 and returns a zero gradient there (a sample whose
    // prediction matches its target exactly must not put NaN into every parameter gradient)
    scale = dn > 0.0f ? scale / (dn * yn) : 0.0f;
    const float* xp = x + b * N;
    const float* yp = y + b * N;
    float* gp = gx + b * N;
    for (long i = (long)blockIdx.y * 256 + threadIdx.x; i < N; i += (long)gridDim.y * 256) gp[i] = (xp[i] - yp[i]) * scale;
}
