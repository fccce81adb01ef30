// Fused per-pixel MLP kernels for the model glue around the operator blocks.
// Included by backend_cuda.cu inside namespace uno::{anonymous}, after gelu_f / gelu_grad_f.
//
//   lift    (darcy_flow_uno2d.py:96-107, navier_stokes_uno2d.py:191-201, navier_stokes_uno3d.py:497-511)
//           cat(a, grid) -> Linear -> GELU -> Linear -> GELU -> permute to channels-first -> zero pad
//           x is read once (channels-last), h written once (channels-first, padded); nothing else touches HBM.
//   project (darcy_flow_uno2d.py:121-131, navier_stokes_uno2d.py:215-225, navier_stokes_uno3d.py:551-575)
//           cat(src_0, src_1, ..) -> crop -> permute to channels-last -> Linear -> GELU -> Linear
//           the concatenated / cropped / permuted tensors are never materialised.
//
// Backward kernels recompute the hidden activations from the inputs (nothing is saved by forward), write the
// input gradients, and reduce the weight gradients over pixels in two levels: a 256-pixel tile is staged in
// shared memory and contracted by all threads ("phase 2", a [rows x 256] * [256 x cols] product), accumulated
// in shared memory across the tiles of a persistent CTA, then flushed with one atomicAdd per element per CTA.
//
// Bound: fwd kernels are HBM-bound (lift writes 4*out_ch B/pixel, project reads 4*sum(src_ch) B/pixel);
// the backward kernels are fp32-FMA-bound on the SIMT pipes (3x the forward MACs).
#pragma once

struct PixGeom {
    int n0, n1, n2;      // raw grid
    int N0, N1, N2;      // padded grid
    int lo0, lo1, lo2;   // raw origin inside the padded grid
    long nraw, npad;     // pixels per sample
};

constexpr int kPixTP = 256;        // pixels per tile = threads per CTA
constexpr int kPixTPP = 260;       // padded row pitch of the staging buffers (floats, multiple of 4)
constexpr int kProjHC = 32;        // hidden units per round of the projection backward (rows of the staged dL/dpre tile)

// Exact-erf GELU and its derivative from ONE exponential: with E = exp(-x^2/2) and t = 1/(1 + p|x|/sqrt2),
// 1 - erf(|x|/sqrt2) = E * t*(a1 + t*(a2 + t*(a3 + t*(a4 + t*a5))))  (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7),
// so Phi(x) = h for x < 0 and 1 - h for x >= 0 with h = E*poly/2 (no cancellation in the negative tail), and the
// derivative Phi(x) + x*E/sqrt(2 pi) reuses E.  Measured in fp32 against fp64: |gelu error| <= 4.3e-7,
// |gelu' error| <= 3.2e-7 over [-12, 12] -- two orders below the parity tolerance; ~16 instructions instead of ~60.
__device__ __forceinline__ float gelu_act(float x) { return gelu_f(x); }
__device__ __forceinline__ float gelu_der(float x) {
    float a, g;
    gelu_both(x, a, g);
    return g;
}

__device__ __forceinline__ float dot_tile(const float* __restrict__ r, const float* __restrict__ c) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
    for (int p = 0; p < kPixTP; p += 4) {
        const float4 a = *reinterpret_cast<const float4*>(r + p);
        const float4 b = *reinterpret_cast<const float4*>(c + p);
        s0 = fmaf(a.x, b.x, s0); s1 = fmaf(a.y, b.y, s1); s2 = fmaf(a.z, b.z, s2); s3 = fmaf(a.w, b.w, s3);
    }
    return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ float sum_tile(const float* __restrict__ r) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 8
    for (int p = 0; p < kPixTP; p += 4) {
        const float4 a = *reinterpret_cast<const float4*>(r + p);
        s0 += a.x; s1 += a.y; s2 += a.z; s3 += a.w;
    }
    return (s0 + s1) + (s2 + s3);
}
// acc[(4 ab + r) * ldo + 4 bb + j] += sum over the tile's 256 pixels of A[4 ab + r][px] * Bm[4 bb + j][px]   (rows kPixTPP apart)
// for ab < nab, bb < nbb, skipping rows >= a_valid and columns >= b_valid.  Register-tiled: a warp takes a block of four A rows,
// its lanes split into four B blocks x eight pixel phases (lane = 8 * bq + pr, pixels 4 (pr + 8 step) .. +3), so every 16-byte
// shared-memory load feeds sixteen products -- the one-dot-product-per-thread form it replaces read two floats per product and
// was bound by shared-memory bandwidth (256 16-byte loads per thread and tile against 64 here).  Partial sums of the eight
// pixel phases meet in three shuffles.  Every lane of the CTA must call it (uniform control flow).
__device__ __forceinline__ void tile_outer_4x4(const float* __restrict__ A, int nab, int a_valid, const float* __restrict__ Bm, int nbb,
                                               int b_valid, float* __restrict__ acc, int ldo) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pr = lane & 7, bq = lane >> 3;
    for (int ab = warp; ab < nab; ab += kPixTP / 32) {
        for (int bb0 = 0; bb0 < nbb; bb0 += 4) {
            const int bb = bb0 + bq;
            const bool active = bb < nbb;
            const float* ap = A + (size_t)(4 * ab) * kPixTPP + 4 * pr;
            const float* bp = Bm + (size_t)(4 * (active ? bb : 0)) * kPixTPP + 4 * pr;
            float s[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[r][j] = 0.f;
#pragma unroll 2
            for (int step = 0; step < kPixTP / 32; ++step) {
                float4 a[4], b[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(ap + r * kPixTPP + 32 * step);
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(bp + j * kPixTPP + 32 * step);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        s[r][j] = fmaf(a[r].x, b[j].x, s[r][j]);
                        s[r][j] = fmaf(a[r].y, b[j].y, s[r][j]);
                        s[r][j] = fmaf(a[r].z, b[j].z, s[r][j]);
                        s[r][j] = fmaf(a[r].w, b[j].w, s[r][j]);
                    }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v = s[r][j];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    if (pr == 0 && active && 4 * ab + r < a_valid && 4 * bb + j < b_valid) acc[(4 * ab + r) * ldo + 4 * bb + j] += v;
                }
        }
    }
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__host__ __device__ inline int round4(int n) { return (n + 3) & ~3; }

// linear index over [batch, raw grid] -> (sample, linear index inside the PADDED grid); 32-bit arithmetic when it fits
__device__ __forceinline__ void raw_to_padded(const PixGeom& g, long idx, long& b, long& rp, long& pp) {
    int r0, r1, r2;
    if (idx < 0x7fffffffL) {
        const unsigned i = (unsigned)idx, nraw = (unsigned)g.nraw;
        const unsigned bb = i / nraw, r = i - bb * nraw;
        const unsigned t = r / (unsigned)g.n2;
        r2 = (int)(r - t * (unsigned)g.n2);
        r0 = (int)(t / (unsigned)g.n1);
        r1 = (int)(t - (unsigned)r0 * (unsigned)g.n1);
        b = bb; rp = r;
    } else {
        b = idx / g.nraw;
        rp = idx - b * g.nraw;
        r2 = (int)(rp % g.n2);
        const long t = rp / g.n2;
        r1 = (int)(t % g.n1); r0 = (int)(t / g.n1);
    }
    pp = ((long)(r0 + g.lo0) * g.N1 + (r1 + g.lo1)) * g.N2 + (r2 + g.lo2);
}

// =====================================================================================================
// lift
// =====================================================================================================
struct LiftK {
    PixGeom g;
    int batch, raw_ch, grid_ch, cin, hid, out_ch;
    const float *a, *grid, *w_a, *b_a, *w_b, *b_b;
    float* h;
    const float* gh;
    const float* gh2;     // optional second upstream gradient (h feeds two consumers): added on the fly
    float *ga, *gw_a, *gb_a, *gw_b, *gb_b;
};

// weights into shared memory, zero-padded to the template sizes (padded hidden units output gelu(0) = 0)
template <int CIN, int HID>
__device__ __forceinline__ void lift_stage_weights(const LiftK& k, float* sWa, float* sba, float* sWb, float* sbb) {
    const int tid = threadIdx.x;
    for (int i = tid; i < HID * CIN; i += kPixTP) {
        const int kk = i / CIN, ci = i % CIN;
        sWa[i] = (kk < k.hid && ci < k.cin) ? __ldg(k.w_a + kk * k.cin + ci) : 0.f;
    }
    for (int i = tid; i < HID; i += kPixTP) sba[i] = i < k.hid ? __ldg(k.b_a + i) : 0.f;
    for (int i = tid; i < k.out_ch * HID; i += kPixTP) {
        const int c = i / HID, kk = i % HID;
        sWb[i] = kk < k.hid ? __ldg(k.w_b + c * k.hid + kk) : 0.f;
    }
    for (int i = tid; i < k.out_ch; i += kPixTP) sbb[i] = __ldg(k.b_b + i);
}

template <int CIN>
__device__ __forceinline__ void lift_load_in(const LiftK& k, long b, long rpix, bool valid, float (&in)[CIN]) {
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
        float v = 0.f;
        if (valid) {
            if (ci < k.raw_ch) v = __ldg(k.a + (b * k.g.nraw + rpix) * k.raw_ch + ci);
            else if (ci < k.cin) v = __ldg(k.grid + rpix * k.grid_ch + (ci - k.raw_ch));
        }
        in[ci] = v;
    }
}

template <int CIN, int HID>
__device__ __forceinline__ void lift_first_layer(const float* sWa, const float* sba, const float (&in)[CIN], float (&pre0)[HID]) {
#pragma unroll
    for (int kk = 0; kk < HID; ++kk) {
        float acc = sba[kk];
        const float4* w4 = reinterpret_cast<const float4*>(sWa + kk * CIN);
#pragma unroll
        for (int q = 0; q < CIN / 4; ++q) {
            const float4 w = w4[q];
            acc = fmaf(w.x, in[4 * q + 0], acc);
            acc = fmaf(w.y, in[4 * q + 1], acc);
            acc = fmaf(w.z, in[4 * q + 2], acc);
            acc = fmaf(w.w, in[4 * q + 3], acc);
        }
        pre0[kk] = acc;
    }
}

template <int HID>
__device__ __forceinline__ float lift_second_layer_row(const float* sWb_row, float bias, const float (&a0)[HID]) {
    float acc0 = bias, acc1 = 0.f;
    const float4* w4 = reinterpret_cast<const float4*>(sWb_row);
#pragma unroll
    for (int q = 0; q < HID / 4; ++q) {
        const float4 w = w4[q];
        acc0 = fmaf(w.x, a0[4 * q + 0], acc0);
        acc1 = fmaf(w.y, a0[4 * q + 1], acc1);
        acc0 = fmaf(w.z, a0[4 * q + 2], acc0);
        acc1 = fmaf(w.w, a0[4 * q + 3], acc1);
    }
    return acc0 + acc1;
}

template <int CIN, int HID>
__global__ void __launch_bounds__(kPixTP) lift_fwd_kernel(const LiftK k) {
    extern __shared__ __align__(16) float psm[];
    float* sWa = psm;                         // [HID][CIN]
    float* sba = sWa + HID * CIN;             // [HID]
    float* sWb = sba + HID;                   // [out_ch][HID]
    float* sbb = sWb + round4(k.out_ch) * HID;
    lift_stage_weights<CIN, HID>(k, sWa, sba, sWb, sbb);
    __syncthreads();
    const PixGeom g = k.g;
    const long total = (long)k.batch * g.npad;
    for (long idx = (long)blockIdx.x * kPixTP + threadIdx.x; idx < total; idx += (long)gridDim.x * kPixTP) {
        const long b = idx / g.npad;
        const long pp = idx - b * g.npad;
        const int i2 = (int)(pp % g.N2);
        const long t = pp / g.N2;
        const int i1 = (int)(t % g.N1), i0 = (int)(t / g.N1);
        const int r0 = i0 - g.lo0, r1 = i1 - g.lo1, r2 = i2 - g.lo2;
        const bool inside = (unsigned)r0 < (unsigned)g.n0 && (unsigned)r1 < (unsigned)g.n1 && (unsigned)r2 < (unsigned)g.n2;
        float* hp = k.h + b * k.out_ch * g.npad + pp;
        if (inside) {
            const long rpix = ((long)r0 * g.n1 + r1) * g.n2 + r2;
            float in[CIN], a0[HID];
            lift_load_in<CIN>(k, b, rpix, true, in);
            lift_first_layer<CIN, HID>(sWa, sba, in, a0);
#pragma unroll
            for (int kk = 0; kk < HID; ++kk) a0[kk] = gelu_act(a0[kk]);
#pragma unroll 4
            for (int c = 0; c < k.out_ch; ++c)
                hp[(long)c * g.npad] = gelu_act(lift_second_layer_row<HID>(sWb + c * HID, sbb[c], a0));
        } else {
            for (int c = 0; c < k.out_ch; ++c) hp[(long)c * g.npad] = 0.f;
        }
    }
}

inline size_t lift_fwd_smem(int CIN, int HID, int out_ch) {
    return sizeof(float) * (size_t)(HID * CIN + HID + round4(out_ch) * HID + round4(out_ch));
}

template <int CIN, int HID>
__global__ void __launch_bounds__(kPixTP) lift_bwd_kernel(const LiftK k, long ntiles) {
    extern __shared__ __align__(16) float psm[];
    const int OC4 = round4(k.out_ch);
    float* sWa = psm;                         // [HID][CIN]
    float* sba = sWa + HID * CIN;             // [HID]
    float* sWb = sba + HID;                   // [out_ch][HID]
    float* sbb = sWb + OC4 * HID;             // [out_ch]
    float* accWa = sbb + OC4;                 // same shapes, gradient accumulators
    float* accba = accWa + HID * CIN;
    float* accWb = accba + HID;
    float* accbb = accWb + OC4 * HID;
    float* GH = accbb + OC4;                  // 2 x [out_ch][TPP]  staged dL/dh of a tile, turned in place into dL/dpre1
    float* A0 = GH + (size_t)2 * OC4 * kPixTPP;   // [HID][TPP]     hidden activations
    float* D0 = A0 + HID * kPixTPP;           // [HID][TPP]     dL/dpre0
    float* IN = D0 + HID * kPixTPP;           // [CIN][TPP]     layer-0 inputs
    const int tid = threadIdx.x;
    lift_stage_weights<CIN, HID>(k, sWa, sba, sWb, sbb);
    for (int i = tid; i < HID * CIN + HID + OC4 * HID + OC4; i += kPixTP) accWa[i] = 0.f;
    __syncthreads();
    const PixGeom g = k.g;
    const long total = (long)k.batch * g.nraw;
    // the upstream gradient of a tile is copied asynchronously (LDGSTS) one tile ahead: thread t fetches pixel t, all
    // channels (consecutive channels are npad elements apart), zero-filled past the end
    auto stage_gh = [&](long tile, float* dstbuf) {
        const long idx = tile * kPixTP + tid;
        const bool valid = idx < total;
        long b = 0, rp = 0, pp = 0;
        if (valid) raw_to_padded(g, idx, b, rp, pp);
        const float* src = valid ? k.gh + b * k.out_ch * g.npad + pp : k.w_a;
        const long step = valid ? g.npad : 0;
        const int sz = valid ? 4 : 0;
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(dstbuf + tid);
#pragma unroll 4
        for (int c = 0; c < k.out_ch; ++c) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
            dst += (uint32_t)(kPixTPP * 4);
            src += step;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int cur = 0;
    if ((long)blockIdx.x < ntiles) stage_gh(blockIdx.x, GH);
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        float* D1 = GH + (size_t)cur * OC4 * kPixTPP;
        {
            const long next = tile + gridDim.x;
            if (next < ntiles) {
                stage_gh(next, GH + (size_t)(cur ^ 1) * OC4 * kPixTPP);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            cur ^= 1;
        }
        const long idx = tile * kPixTP + tid;
        const bool valid = idx < total;
        long b = 0, rp = 0, pp = 0;
        if (valid) raw_to_padded(g, idx, b, rp, pp);
        float in[CIN], a0[HID], gp0[HID], da0[HID];
        const float* gh2p = (k.gh2 != nullptr && valid) ? k.gh2 + b * k.out_ch * g.npad + pp : nullptr;
        lift_load_in<CIN>(k, b, rp, valid, in);
        lift_first_layer<CIN, HID>(sWa, sba, in, a0);
#pragma unroll
        for (int kk = 0; kk < HID; ++kk) {
            float act, grad;
            gelu_both(a0[kk], act, grad);
            a0[kk] = act; gp0[kk] = grad; da0[kk] = 0.f;
        }
        // the second upstream gradient (when h has two consumers) is not staged: its values for the NEXT eight channels are
        // loaded into registers while the current eight are processed (coalesced: consecutive threads, consecutive pixels)
        constexpr int G2 = 8;
        float g2n[G2];
#pragma unroll
        for (int j = 0; j < G2; ++j) g2n[j] = (gh2p && j < k.out_ch) ? __ldg(gh2p + (long)j * g.npad) : 0.f;
        for (int c0 = 0; c0 < k.out_ch; c0 += G2) {
            float g2c[G2];
#pragma unroll
            for (int j = 0; j < G2; ++j) {
                g2c[j] = g2n[j];
                g2n[j] = (gh2p && c0 + G2 + j < k.out_ch) ? __ldg(gh2p + (long)(c0 + G2 + j) * g.npad) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < G2; ++j) {
                const int c = c0 + j;
                if (c < k.out_ch) {
                    const float pre1 = lift_second_layer_row<HID>(sWb + c * HID, sbb[c], a0);
                    const float gv = D1[c * kPixTPP + tid] + g2c[j];   // own asynchronous copy (complete after wait_group) + second source
                    const float d1 = gv * gelu_der(pre1);
                    D1[c * kPixTPP + tid] = d1;
                    const float4* w4 = reinterpret_cast<const float4*>(sWb + c * HID);
#pragma unroll
                    for (int q = 0; q < HID / 4; ++q) {
                        const float4 w = w4[q];
                        da0[4 * q + 0] = fmaf(w.x, d1, da0[4 * q + 0]);
                        da0[4 * q + 1] = fmaf(w.y, d1, da0[4 * q + 1]);
                        da0[4 * q + 2] = fmaf(w.z, d1, da0[4 * q + 2]);
                        da0[4 * q + 3] = fmaf(w.w, d1, da0[4 * q + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int kk = 0; kk < HID; ++kk) {
            da0[kk] *= gp0[kk];                       // dL/dpre0
            D0[kk * kPixTPP + tid] = da0[kk];
            A0[kk * kPixTPP + tid] = a0[kk];
        }
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) IN[ci * kPixTPP + tid] = in[ci];
        if (k.ga != nullptr && valid) {
            float* gap = k.ga + (b * g.nraw + rp) * k.raw_ch;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                if (ci < k.raw_ch) {
                    float s = 0.f;
#pragma unroll
                    for (int kk = 0; kk < HID; ++kk) s = fmaf(sWa[kk * CIN + ci], da0[kk], s);
                    gap[ci] = s;
                }
            }
        }
        __syncthreads();
        // ---- phase 2: contract the staged tile over its 256 pixels
        tile_outer_4x4(D1, OC4 / 4, k.out_ch, A0, HID / 4, HID, accWb, HID);        // dW_b[c][kk] += D1[c] . A0[kk]
        tile_outer_4x4(D0, HID / 4, HID, IN, CIN / 4, CIN, accWa, CIN);             // dW_a[kk][ci] += D0[kk] . IN[ci]
        if (tid < k.out_ch) accbb[tid] += sum_tile(D1 + tid * kPixTPP);
        else if (tid >= 128 && tid < 128 + HID) accba[tid - 128] += sum_tile(D0 + (tid - 128) * kPixTPP);
        __syncthreads();
    }
    // ---- flush: one atomic per element per CTA
    for (int e = tid; e < k.out_ch * HID; e += kPixTP) {
        const int c = e / HID, kk = e % HID;
        if (kk < k.hid) atomicAdd(k.gw_b + c * k.hid + kk, accWb[e]);
    }
    for (int e = tid; e < HID * CIN; e += kPixTP) {
        const int kk = e / CIN, ci = e % CIN;
        if (kk < k.hid && ci < k.cin) atomicAdd(k.gw_a + kk * k.cin + ci, accWa[e]);
    }
    for (int e = tid; e < k.out_ch; e += kPixTP) atomicAdd(k.gb_b + e, accbb[e]);
    for (int e = tid; e < k.hid; e += kPixTP) atomicAdd(k.gb_a + e, accba[e]);
}

inline size_t lift_bwd_smem(int CIN, int HID, int out_ch) {
    const int OC4 = round4(out_ch);
    return sizeof(float) * ((size_t)2 * (HID * CIN + HID + OC4 * HID + OC4) + (size_t)(2 * OC4 + 2 * HID + CIN) * kPixTPP);
}

// =====================================================================================================
// project
// =====================================================================================================
struct ProjK {
    PixGeom g;
    int batch, nsrc, ctot, hid, out_ch;
    const float* src[4];
    float* gsrc[4];
    int src_ch[4];
    const float *w1, *b1, *w2, *b2;
    float* out;
    float* pre_out;          // optional [hid][batch * nraw]
    const float* pre_in;     // optional, same layout
    const float* gout;
    float *gw1, *gb1, *gw2, *gb2;
};
constexpr int kProjMaxOut = 4;

// per-channel base pointers (channel c of the virtual concatenation) and batch strides
template <int CT>
__device__ __forceinline__ void proj_stage_tables(const ProjK& k, const float** sbase, float** gbase, long* sstride) {
    for (int c = threadIdx.x; c < CT; c += kPixTP) {
        int s = 0, cl = c;
        while (s < k.nsrc && cl >= k.src_ch[s]) { cl -= k.src_ch[s]; ++s; }
        if (s < k.nsrc) {
            sbase[c] = k.src[s] + (long)cl * k.g.npad;
            gbase[c] = k.gsrc[s] ? k.gsrc[s] + (long)cl * k.g.npad : nullptr;
            sstride[c] = (long)k.src_ch[s] * k.g.npad;
        } else {
            sbase[c] = nullptr; gbase[c] = nullptr; sstride[c] = 0;
        }
    }
}

template <int CT>
__device__ __forceinline__ void proj_stage_weights(const ProjK& k, float* sW1, float* sb1, float* sW2, float* sb2) {
    const int tid = threadIdx.x;
    for (int i = tid; i < k.hid * CT; i += kPixTP) {
        const int n = i / CT, c = i % CT;
        sW1[i] = c < k.ctot ? __ldg(k.w1 + n * k.ctot + c) : 0.f;
    }
    for (int i = tid; i < k.hid; i += kPixTP) sb1[i] = __ldg(k.b1 + i);
    for (int i = tid; i < k.out_ch * k.hid; i += kPixTP) sW2[i] = __ldg(k.w2 + i);
    if (sb2) for (int i = tid; i < kProjMaxOut; i += kPixTP) sb2[i] = i < k.out_ch ? __ldg(k.b2 + i) : 0.f;
}

template <int CT>
__device__ __forceinline__ float proj_hidden_row(const float* sW1_row, float bias, const float (&in)[CT]) {
    float acc0 = bias, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const float4* w4 = reinterpret_cast<const float4*>(sW1_row);
#pragma unroll
    for (int q = 0; q < CT / 4; ++q) {
        const float4 w = w4[q];
        acc0 = fmaf(w.x, in[4 * q + 0], acc0);
        acc1 = fmaf(w.y, in[4 * q + 1], acc1);
        acc2 = fmaf(w.z, in[4 * q + 2], acc2);
        acc3 = fmaf(w.w, in[4 * q + 3], acc3);
    }
    return (acc0 + acc1) + (acc2 + acc3);
}

__host__ __device__ inline size_t proj_table_bytes(int CT) { return (size_t)CT * (2 * sizeof(void*) + sizeof(long)); }

// Forward of the projection: a CTA walks 256-pixel tiles of the cropped grid; the tile's inputs are staged with LDGSTS
// (thread t copies pixel t, all channels) and the hidden layer is a register-tiled product out of shared memory,
// PRE[32 x 256] = W1[chunk] * IN with 8 hidden x 4 pixel thread tiles (pixels tp, tp+64, tp+128, tp+192: every access
// of a warp is to 32 consecutive pixels).  GELU and the fc2 dot product are applied to the thread tile; the four
// hidden-block partial sums of a pixel meet in shared memory.  Optionally writes the pre-activations for backward.
template <int CT>
__global__ void __launch_bounds__(kPixTP, 2) proj_fwd_kernel(const ProjK k, long ntiles) {
    extern __shared__ __align__(16) float psm[];
    const float** sbase = reinterpret_cast<const float**>(psm);
    float** gbase = reinterpret_cast<float**>(psm) + CT;
    long* sstride = reinterpret_cast<long*>(gbase + CT);
    float* sW1 = reinterpret_cast<float*>(sstride + CT);   // [hid][CT]
    float* sb1 = sW1 + (size_t)k.hid * CT;                 // [hid]
    float* sW2 = sb1 + round4(k.hid);                      // [out_ch][hid]
    float* sb2 = sW2 + round4(k.out_ch * k.hid);           // [4]
    float* OUTP = sb2 + 4;                                 // [4 hidden blocks][256 pixels][4 outputs] partial fc2 sums
    float* IN = OUTP + 4 * kPixTP * kProjMaxOut;           // [CT][TPP]
    const int tid = threadIdx.x;
    proj_stage_tables<CT>(k, sbase, gbase, sstride);
    proj_stage_weights<CT>(k, sW1, sb1, sW2, sb2);
    for (int i = tid; i < (CT - k.ctot) * kPixTPP; i += kPixTP) IN[(size_t)k.ctot * kPixTPP + i] = 0.f;
    __syncthreads();
    const PixGeom g = k.g;
    const long total = (long)k.batch * g.nraw;
    const int tn = tid >> 6, tp = tid & 63;
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long base = tile * kPixTP;
        {   // stage the tile (two CTAs per SM overlap each other's copies)
            const long idx = base + tid;
            const bool valid = idx < total;
            long b = 0, rp = 0, pp = 0;
            if (valid) raw_to_padded(g, idx, b, rp, pp);
            uint32_t dst = (uint32_t)__cvta_generic_to_shared(IN + tid);
            const int sz = valid ? 4 : 0;
            const long step = valid ? g.npad : 0;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (s < k.nsrc) {
                    const int nch = k.src_ch[s];
                    const float* src = valid ? k.src[s] + b * nch * g.npad + pp : k.w1;
#pragma unroll 4
                    for (int cl = 0; cl < nch; ++cl) {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
                        dst += (uint32_t)(kPixTPP * 4);
                        src += step;
                    }
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        __syncthreads();
        float o[4][kProjMaxOut];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int u = 0; u < kProjMaxOut; ++u) o[q][u] = 0.f;
        for (int ch0 = 0; ch0 < k.hid; ch0 += kProjHC) {
            float acc[8][4];
            const float* wrow[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int n = min(ch0 + 8 * tn + i, k.hid - 1);      // rows past hid run on a clamped row and are discarded
                wrow[i] = sW1 + (size_t)n * CT;
                const float bv = sb1[n];
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[i][q] = bv;
            }
            const float* xin = IN + tp;
#pragma unroll 2
            for (int c4 = 0; c4 < CT; c4 += 4) {
                float x[4][4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                    for (int q = 0; q < 4; ++q) x[cc][q] = xin[(size_t)(c4 + cc) * kPixTPP + 64 * q];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 w = *reinterpret_cast<const float4*>(wrow[i] + c4);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        acc[i][q] = fmaf(w.x, x[0][q], acc[i][q]);
                        acc[i][q] = fmaf(w.y, x[1][q], acc[i][q]);
                        acc[i][q] = fmaf(w.z, x[2][q], acc[i][q]);
                        acc[i][q] = fmaf(w.w, x[3][q], acc[i][q]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int n = ch0 + 8 * tn + i;
                if (n < k.hid) {                                     // uniform over the warp (tn is)
                    float w2v[kProjMaxOut];
#pragma unroll
                    for (int u = 0; u < kProjMaxOut; ++u) w2v[u] = u < k.out_ch ? sW2[u * k.hid + n] : 0.f;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const long idx = base + tp + 64 * q;
                        if (k.pre_out != nullptr && idx < total) k.pre_out[(size_t)n * total + idx] = acc[i][q];
                        const float a = gelu_act(acc[i][q]);
#pragma unroll
                        for (int u = 0; u < kProjMaxOut; ++u) o[q][u] = fmaf(w2v[u], a, o[q][u]);
                    }
                }
            }
        }
        // the four hidden blocks of a pixel meet in shared memory
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(OUTP + ((size_t)tn * kPixTP + tp + 64 * q) * kProjMaxOut) = make_float4(o[q][0], o[q][1], o[q][2], o[q][3]);
        __syncthreads();
        {
            const long idx = base + tid;
            if (idx < total) {
                float4 r = make_float4(sb2[0], sb2[1], sb2[2], sb2[3]);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float4 v = *reinterpret_cast<const float4*>(OUTP + ((size_t)t * kPixTP + tid) * kProjMaxOut);
                    r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
                }
                const float rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int u = 0; u < kProjMaxOut; ++u)
                    if (u < k.out_ch) k.out[idx * k.out_ch + u] = rr[u];
            }
        }
        __syncthreads();       // IN and OUTP are rewritten by the next tile
    }
}

inline size_t proj_fwd_smem(int CT, int hid, int out_ch) {
    return proj_table_bytes(CT) + sizeof(float) * ((size_t)hid * CT + round4(hid) + round4(out_ch * hid) + 4 + 4 * kPixTP * kProjMaxOut +
                                                   (size_t)CT * kPixTPP);
}

// The crop has no gradient in the padding: zero it.  A warp takes one row (b, i0, i1) of the padded grid at a time and its
// lanes the columns that lie outside the crop (all of them for a padded row), so only the padding is ever visited.
__device__ __forceinline__ void proj_zero_padding(const ProjK& k, float* const* gbase, const long* sstride) {
    const PixGeom g = k.g;
    if (g.npad == g.nraw) return;
    const int lane = threadIdx.x & 31;
    const long warps = (long)gridDim.x * (blockDim.x >> 5), w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long rows = (long)k.batch * g.N0 * g.N1;
    for (long row = w; row < rows; row += warps) {
        const long b = row / ((long)g.N0 * g.N1);
        const int rr = (int)(row - b * (long)g.N0 * g.N1);
        const int i0 = rr / g.N1, i1 = rr - i0 * g.N1;
        const bool row_inside = (unsigned)(i0 - g.lo0) < (unsigned)g.n0 && (unsigned)(i1 - g.lo1) < (unsigned)g.n1;
        const long pp0 = ((long)i0 * g.N1 + i1) * g.N2;
        // columns outside [lo2, lo2 + n2) -- or every column of a padded row
        const int left = row_inside ? g.lo2 : g.N2, right0 = g.lo2 + g.n2, nright = row_inside ? g.N2 - right0 : 0;
        for (int j = lane; j < left + nright; j += 32) {
            const int i2 = j < left ? j : right0 + (j - left);
            for (int c = 0; c < k.ctot; ++c)
                if (gbase[c] != nullptr) gbase[c][b * sstride[c] + pp0 + i2] = 0.f;
        }
    }
}

// Backward of the projection.  A CTA walks 256-pixel tiles of the CROPPED grid; per tile and per chunk of 32 hidden
// units it runs three register-tiled products out of shared memory
//     PRE [32 x 256] = W1[chunk] * IN            (+ GELU / GELU' and the fc2 back-substitution -> D = dL/dpre)
//     DIN [CT x 256] += W1[chunk]^T * D          (accumulated in registers over the chunks, then stored to gsrc)
//     dW1[chunk]    += D * IN^T                  (accumulated in shared memory over the tiles of the CTA)
// with thread tiles 8x4, (CT/4)x4 and 2x4.  A thread's four pixels are tp, tp+64, tp+128, tp+192 so that every
// shared / global access of a warp is to 32 consecutive pixels.
template <int CT>
__global__ void __launch_bounds__(kPixTP, 1) proj_bwd_kernel(const ProjK k, long ntiles, int nbuf) {
    extern __shared__ __align__(16) float psm[];
    const float** sbase = reinterpret_cast<const float**>(psm);
    float** gbase = reinterpret_cast<float**>(psm) + CT;
    long* sstride = reinterpret_cast<long*>(gbase + CT);
    const int H4 = round4(k.hid), OH4 = round4(k.out_ch * k.hid);
    float* sW1 = reinterpret_cast<float*>(sstride + CT);   // [hid][CT]
    float* sb1 = sW1 + (size_t)k.hid * CT;                 // [hid]
    float* sW2 = sb1 + H4;                                 // [out_ch][hid]
    float* accW1 = sW2 + OH4;                              // gradient accumulators, same shapes
    float* accb1 = accW1 + (size_t)k.hid * CT;
    float* accW2 = accb1 + H4;
    float* accb2 = accW2 + OH4;                            // [4]
    float* D = accb2 + 4;                                  // [kProjHC][TPP]  dL/dpre1 of the current hidden chunk
    float* INbuf = D + (size_t)kProjHC * kPixTPP;          // nbuf x [CT][TPP] cropped inputs of a tile (double-buffered when it fits)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    proj_stage_tables<CT>(k, sbase, gbase, sstride);
    proj_stage_weights<CT>(k, sW1, sb1, sW2, nullptr);
    for (int i = tid; i < k.hid * CT + H4 + OH4 + 4; i += kPixTP) accW1[i] = 0.f;
    __syncthreads();
    const PixGeom g = k.g;
    const long total = (long)k.batch * g.nraw;
    // asynchronous staging of a tile's inputs (LDGSTS: no registers, every load of the tile in flight at once);
    // thread t copies pixel t of the tile, all channels; missing pixels / channels are zero-filled (src-size 0)
    auto stage_tile = [&](long tile, float* dstbuf) {
        const long idx = tile * kPixTP + tid;
        const bool valid = idx < total;
        long b = 0, rp = 0, pp = 0;
        if (valid) raw_to_padded(g, idx, b, rp, pp);
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(dstbuf + tid);
        const int sz = valid ? 4 : 0;
        const long step = valid ? g.npad : 0;
        // source by source: consecutive channels of one source are npad elements apart, so the address is a running sum
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (s < k.nsrc) {
                const int nch = k.src_ch[s];
                const float* src = valid ? k.src[s] + b * nch * g.npad + pp : k.w1;
#pragma unroll 4
                for (int cl = 0; cl < nch; ++cl) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
                    dst += (uint32_t)(kPixTPP * 4);
                    src += step;
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // rows of channels past ctot (template padding) stay zero for the whole kernel
    for (int i = tid; i < (CT - k.ctot) * kPixTPP; i += kPixTP) {
        INbuf[(size_t)k.ctot * kPixTPP + i] = 0.f;
        if (nbuf == 2) INbuf[(size_t)(CT + k.ctot) * kPixTPP + i] = 0.f;
    }
    // ---- the padding of the source gradients is zero (the crop has no gradient there)
    proj_zero_padding(k, gbase, sstride);
    constexpr int CB = CT / 4;                             // channels per thread in the DIN product
    constexpr int WC = CT / 32, WN = 8 / WC;               // dW1 product: a warp owns 8 (hidden) x 32 (channel), a thread 2 x 4
    const int tn = tid >> 6, tp = tid & 63;                // PRE / DIN products: hidden (channel) block, pixel lane
    const int c_l = lane & 7, n_l = lane >> 3;
    const int cbase = (warp % WC) * 32, nwarp = (warp / WC) * 8;
    int cur = 0;
    if (nbuf == 2 && (long)blockIdx.x < ntiles) stage_tile(blockIdx.x, INbuf);
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long base = tile * kPixTP;
        float* IN = INbuf + (size_t)cur * CT * kPixTPP;
        if (nbuf == 2) {
            const long next = tile + gridDim.x;
            if (next < ntiles) {
                stage_tile(next, INbuf + (size_t)(cur ^ 1) * CT * kPixTPP);   // overlaps this tile's arithmetic
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            cur ^= 1;
        } else {
            stage_tile(tile, IN);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        // ---- this thread's four pixels in the register-tiled products
        long pb[4], ppx[4];
        float go[4][kProjMaxOut];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long idx = base + tp + 64 * q;
            const bool valid = idx < total;
            long b = -1, rp = 0, pp = 0;
            if (valid) raw_to_padded(g, idx, b, rp, pp);
            pb[q] = b;
            ppx[q] = pp;
#pragma unroll
            for (int o = 0; o < kProjMaxOut; ++o) go[q][o] = (valid && o < k.out_ch) ? __ldg(k.gout + idx * k.out_ch + o) : 0.f;
        }
        float dacc[CB][4];
#pragma unroll
        for (int i = 0; i < CB; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) dacc[i][q] = 0.f;
        __syncthreads();
        for (int ch0 = 0; ch0 < k.hid; ch0 += kProjHC) {
            const int nn = min(kProjHC, k.hid - ch0);
            // ---- PRE = W1[chunk] * IN for hidden units ch0 + 8 tn + i, then D = (W2^T gout) * gelu'(PRE)
            {
                float acc[8][4];
                if (k.pre_in != nullptr) {
                    // saved by the forward pass: [hid][pixels], a warp reads 32 consecutive pixels of one hidden unit
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int n = min(ch0 + 8 * tn + i, k.hid - 1);
                        const float* src = k.pre_in + (size_t)n * total + base + tp;
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[i][q] = pb[q] >= 0 ? __ldg(src + 64 * q) : 0.f;
                    }
                } else {
                    const float* wrow[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int n = min(ch0 + 8 * tn + i, k.hid - 1);     // rows past hid: computed on a clamped row, discarded below
                        wrow[i] = sW1 + (size_t)n * CT;
                        const float bv = sb1[n];
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[i][q] = bv;
                    }
                    const float* xin = IN + tp;
#pragma unroll 2
                    for (int c4 = 0; c4 < CT; c4 += 4) {
                        float x[4][4];
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                            for (int q = 0; q < 4; ++q) x[cc][q] = xin[(size_t)(c4 + cc) * kPixTPP + 64 * q];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 w = *reinterpret_cast<const float4*>(wrow[i] + c4);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                acc[i][q] = fmaf(w.x, x[0][q], acc[i][q]);
                                acc[i][q] = fmaf(w.y, x[1][q], acc[i][q]);
                                acc[i][q] = fmaf(w.z, x[2][q], acc[i][q]);
                                acc[i][q] = fmaf(w.w, x[3][q], acc[i][q]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = 8 * tn + i;                // row of the chunk
                    const int n = ch0 + j;
                    const bool live = j < nn;
                    const int nc = live ? n : k.hid - 1;
                    float w2v[kProjMaxOut];
#pragma unroll
                    for (int o = 0; o < kProjMaxOut; ++o) w2v[o] = o < k.out_ch ? sW2[o * k.hid + nc] : 0.f;
                    float sb = 0.f, sw[kProjMaxOut];
#pragma unroll
                    for (int o = 0; o < kProjMaxOut; ++o) sw[o] = 0.f;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float a, gp;
                        gelu_both(acc[i][q], a, gp);
                        float s = 0.f;
#pragma unroll
                        for (int o = 0; o < kProjMaxOut; ++o) {
                            s = fmaf(go[q][o], w2v[o], s);
                            sw[o] = fmaf(go[q][o], a, sw[o]);
                        }
                        const float dp = live ? s * gp : 0.f;
                        if (j < kProjHC) D[(size_t)j * kPixTPP + tp + 64 * q] = dp;
                        sb += dp;
                    }
                    // weight-gradient pieces that reduce over pixels only: dW2[o][n], db1[n]
                    sb = warp_sum(sb);
#pragma unroll
                    for (int o = 0; o < kProjMaxOut; ++o)
                        if (o < k.out_ch) sw[o] = warp_sum(sw[o]);
                    if (lane == 0 && live) {
                        atomicAdd(accb1 + n, sb);
#pragma unroll
                        for (int o = 0; o < kProjMaxOut; ++o)
                            if (o < k.out_ch) atomicAdd(accW2 + o * k.hid + n, sw[o]);
                    }
                }
            }
            __syncthreads();
            // ---- DIN += W1[chunk]^T * D : channels tn*CB .. +CB, this thread's four pixels
            {
                const float* dcol = D + tp;
                const float* wbase = sW1 + (size_t)ch0 * CT + tn * CB;
#pragma unroll 4
                for (int j = 0; j < nn; ++j) {
                    float d[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) d[q] = dcol[(size_t)j * kPixTPP + 64 * q];
#pragma unroll
                    for (int i4 = 0; i4 < CB; i4 += 4) {
                        const float4 w = *reinterpret_cast<const float4*>(wbase + (size_t)j * CT + i4);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            dacc[i4 + 0][q] = fmaf(w.x, d[q], dacc[i4 + 0][q]);
                            dacc[i4 + 1][q] = fmaf(w.y, d[q], dacc[i4 + 1][q]);
                            dacc[i4 + 2][q] = fmaf(w.z, d[q], dacc[i4 + 2][q]);
                            dacc[i4 + 3][q] = fmaf(w.w, d[q], dacc[i4 + 3][q]);
                        }
                    }
                }
            }
            // ---- dW1[ch0 + r][c] += sum_p D[r][p] * IN[c][p]
            for (int n0 = 0; n0 < nn; n0 += 8 * WN) {
                const int nb = n0 + nwarp;
                if (nb < nn) {
                    float acc[2][4];
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
                    const int ra = min(nb + n_l, kProjHC - 1), rb = min(nb + n_l + 4, kProjHC - 1);
                    const float* drow0 = D + (size_t)ra * kPixTPP;
                    const float* drow1 = D + (size_t)rb * kPixTPP;
                    const float* icol = IN + (size_t)(cbase + c_l) * kPixTPP;
#pragma unroll 2
                    for (int p = 0; p < kPixTP; p += 4) {
                        const float4 d0 = *reinterpret_cast<const float4*>(drow0 + p);
                        const float4 d1 = *reinterpret_cast<const float4*>(drow1 + p);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 x = *reinterpret_cast<const float4*>(icol + (size_t)(8 * j) * kPixTPP + p);
                            acc[0][j] = fmaf(d0.x, x.x, acc[0][j]); acc[0][j] = fmaf(d0.y, x.y, acc[0][j]);
                            acc[0][j] = fmaf(d0.z, x.z, acc[0][j]); acc[0][j] = fmaf(d0.w, x.w, acc[0][j]);
                            acc[1][j] = fmaf(d1.x, x.x, acc[1][j]); acc[1][j] = fmaf(d1.y, x.y, acc[1][j]);
                            acc[1][j] = fmaf(d1.z, x.z, acc[1][j]); acc[1][j] = fmaf(d1.w, x.w, acc[1][j]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int r = nb + n_l + 4 * i;
                        if (r < nn) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) accW1[(size_t)(ch0 + r) * CT + cbase + c_l + 8 * j] += acc[i][j];
                        }
                    }
                }
            }
            __syncthreads();
        }
        // ---- input gradients of the tile (a warp stores 32 consecutive pixels of one channel); the per-pixel offset
        //      only depends on the batch stride of the channel's source, so it is recomputed when that changes
        {
            long cur_stride = -1, off[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < CB; ++i) {
                const int c = tn * CB + i;
                if (c < k.ctot) {
                    float* gb = gbase[c];
                    const long st = sstride[c];
                    if (st != cur_stride) {
                        cur_stride = st;
#pragma unroll
                        for (int q = 0; q < 4; ++q) off[q] = pb[q] * st + ppx[q];
                    }
                    if (gb != nullptr) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (pb[q] >= 0) gb[off[q]] = dacc[i][q];
                    }
                }
            }
        }
        if (tn == 0) {
#pragma unroll
            for (int o = 0; o < kProjMaxOut; ++o)
                if (o < k.out_ch) {
                    const float v = warp_sum(go[0][o] + go[1][o] + go[2][o] + go[3][o]);
                    if (lane == 0) atomicAdd(accb2 + o, v);
                }
        }
    }
    __syncthreads();
    for (int i = tid; i < k.hid * CT; i += kPixTP) {
        const int n = i / CT, c = i % CT;
        if (c < k.ctot) atomicAdd(k.gw1 + n * k.ctot + c, accW1[i]);
    }
    for (int i = tid; i < k.hid; i += kPixTP) atomicAdd(k.gb1 + i, accb1[i]);
    for (int i = tid; i < k.out_ch * k.hid; i += kPixTP) atomicAdd(k.gw2 + i, accW2[i]);
    for (int i = tid; i < k.out_ch; i += kPixTP) atomicAdd(k.gb2 + i, accb2[i]);
}

inline size_t proj_bwd_smem(int CT, int hid, int out_ch, int nbuf) {
    return proj_table_bytes(CT) +
           sizeof(float) * ((size_t)2 * ((size_t)hid * CT + round4(hid) + round4(out_ch * hid)) + 4 + (size_t)(kProjHC + nbuf * CT) * kPixTPP);
}
