// Single-pass InstanceNorm (+ GELU) forward and backward with the plane held on chip.
// Included by backend_cuda.cu inside namespace uno::{anonymous}, after block_sum / gelu_f / gelu_grad_f.
//
// nn.InstanceNorm{2,3}d(affine) needs two sweeps over every (b, c) plane -- statistics, then the normalisation --
// and its backward two more.  A 240x240 plane is 230 KB: too big for one CTA's shared memory, so a plane is spread over a
// thread-block CLUSTER of 1, 2, 4 or 8 CTAs; each CTA keeps its slice in shared memory, the partial sums are exchanged
// through distributed shared memory (cluster.map_shared_rank) and every element crosses HBM exactly once per
// direction: forward 4 B read + 4 B written, backward 8 B read + 4 B written per element (the two-kernel path moves
// 12-16 B and 20 B, and evaluates GELU' twice).  Arithmetic is the same as plane_stats / norm_act_fwd / norm_act_bwd
// (two-pass variance, double accumulation of the partial sums).
#pragma once
// (backend_cuda.cu includes <cooperative_groups.h> at global scope)
namespace cgx = ::cooperative_groups;

constexpr int kNormThreads = 512;

// sum over the whole cluster of one double per CTA: `slot` is this CTA's shared-memory mailbox
__device__ __forceinline__ double cluster_total(cgx::cluster_group& cluster, double* slot, double mine) {
    if (threadIdx.x == 0) *slot = mine;
    cluster.sync();
    double t = 0.0;
    const unsigned n = cluster.num_blocks();
    for (unsigned r = 0; r < n; ++r) t += *cluster.map_shared_rank(slot, r);
    return t;
}

__global__ void __launch_bounds__(kNormThreads) norm_fwd_cluster_kernel(const float* __restrict__ x, float* __restrict__ stats,
                                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                        float* __restrict__ y, int C, long L, int slice, float eps,
                                                                        int non_lin) {
    extern __shared__ __align__(16) float nsm[];
    __shared__ double sh[32];
    __shared__ double mail[2];
    cgx::cluster_group cluster = cgx::this_cluster();
    const unsigned cs = cluster.num_blocks(), rank = cluster.block_rank();
    const long p = blockIdx.x / cs;
    const long lo = (long)rank * slice;
    const int n = (int)max(0L, min((long)slice, L - lo));
    const float* xp = x + p * L + lo;
    float* yp = y + p * L + lo;
    // 16-byte body when every slice starts on a 16-byte boundary (L % 4 == 0; slices are multiples of 4), scalar tail
    const bool vec = (L & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const int nv = vec ? n >> 2 : 0;
    const float4* x4 = reinterpret_cast<const float4*>(xp);
    float4* s4 = reinterpret_cast<float4*>(nsm);
    float s = 0.f;
    double sd = 0.0;
    int cnt = 0;
    for (int i = threadIdx.x; i < nv; i += kNormThreads) {
        const float4 v = __ldcs(x4 + i);
        s4[i] = v;
        s += (v.x + v.y) + (v.z + v.w);
        if (++cnt == 16) { sd += s; s = 0.f; cnt = 0; }
    }
    for (int i = 4 * nv + threadIdx.x; i < n; i += kNormThreads) {
        const float v = xp[i];
        nsm[i] = v;
        s += v;
        if (++cnt == 16) { sd += s; s = 0.f; cnt = 0; }
    }
    sd += s;
    const double mu = cluster_total(cluster, &mail[0], block_sum(sd, sh)) / (double)L;
    const float muf = (float)mu;
    s = 0.f; sd = 0.0; cnt = 0;
    for (int i = threadIdx.x; i < nv; i += kNormThreads) {
        const float4 v = s4[i];
        const float a = v.x - muf, b = v.y - muf, c2 = v.z - muf, d = v.w - muf;
        s = fmaf(a, a, s); s = fmaf(b, b, s); s = fmaf(c2, c2, s); s = fmaf(d, d, s);
        if (++cnt == 16) { sd += s; s = 0.f; cnt = 0; }
    }
    for (int i = 4 * nv + threadIdx.x; i < n; i += kNormThreads) {
        const float d = nsm[i] - muf;
        s = fmaf(d, d, s);
        if (++cnt == 16) { sd += s; s = 0.f; cnt = 0; }
    }
    sd += s;
    const double var = cluster_total(cluster, &mail[1], block_sum(sd, sh)) / (double)L;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    if (rank == 0 && threadIdx.x == 0 && stats != nullptr) {
        stats[2 * p] = muf;
        stats[2 * p + 1] = rstd;
    }
    const int c = (int)(p % C);
    const float g = gamma[c] * rstd, b = beta[c] - muf * rstd * gamma[c];
    float4* y4 = reinterpret_cast<float4*>(yp);
    for (int i = threadIdx.x; i < nv; i += kNormThreads) {
        float4 v = s4[i];
        v.x = fmaf(v.x, g, b); v.y = fmaf(v.y, g, b); v.z = fmaf(v.z, g, b); v.w = fmaf(v.w, g, b);
        if (non_lin) { v.x = gelu_f(v.x); v.y = gelu_f(v.y); v.z = gelu_f(v.z); v.w = gelu_f(v.w); }
        y4[i] = v;
    }
    for (int i = 4 * nv + threadIdx.x; i < n; i += kNormThreads) {
        const float v = fmaf(nsm[i], g, b);
        yp[i] = non_lin ? gelu_f(v) : v;
    }
    cluster.sync();      // no CTA may exit while a neighbour can still read its mailbox
}

__global__ void __launch_bounds__(kNormThreads) norm_bwd_cluster_kernel(const UpGradK gy, const float* __restrict__ x,
                                                                        const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta, float* __restrict__ g,
                                                                        float* __restrict__ ggamma, float* __restrict__ gbeta, int C, long L,
                                                                        int slice, int non_lin) {
    extern __shared__ __align__(16) float nsm[];
    __shared__ double sh[32];
    __shared__ double mail[2];
    cgx::cluster_group cluster = cgx::this_cluster();
    const unsigned cs = cluster.num_blocks(), rank = cluster.block_rank();
    const long p = blockIdx.x / cs;
    const int c = (int)(p % C);
    const long lo = (long)rank * slice;
    const int n = (int)max(0L, min((long)slice, L - lo));
    float* gn_s = nsm;             // [slice] upstream gradient through the activation
    float* xh_s = nsm + slice;     // [slice] normalised input
    const float mu = stats[2 * p], rstd = stats[2 * p + 1];
    const float ga = gamma[c], be = beta[c];
    const float* xp = x + p * L + lo;
    // the upstream gradient is the sum of up to two strided sources (backend.h UpGrad)
    const float* gp = upgrad_plane(gy.p0, gy.bs0, p, C, L) + lo;
    const float* hp = gy.p1 ? upgrad_plane(gy.p1, gy.bs1, p, C, L) + lo : nullptr;
    const bool vec = (L & 3) == 0 && ((gy.bs0 | (gy.p1 ? gy.bs1 : 0)) & 3) == 0 &&
                     ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gy.p0) | reinterpret_cast<uintptr_t>(gy.p1) |
                       reinterpret_cast<uintptr_t>(g)) & 15) == 0;
    const int nv = vec ? n >> 2 : 0;
    float s1 = 0.f, s2 = 0.f;
    double d1 = 0.0, d2 = 0.0;
    int cnt = 0;
    {
        const float4* x4 = reinterpret_cast<const float4*>(xp);
        const float4* g4 = reinterpret_cast<const float4*>(gp);
        const float4* h4 = reinterpret_cast<const float4*>(hp);
        float4* gn4 = reinterpret_cast<float4*>(gn_s);
        float4* xh4 = reinterpret_cast<float4*>(xh_s);
        for (int i = threadIdx.x; i < nv; i += kNormThreads) {
            const float4 xv = __ldcs(x4 + i);
            float4 gv = __ldcs(g4 + i);
            if (hp) {
                const float4 hv = __ldcs(h4 + i);
                gv.x += hv.x; gv.y += hv.y; gv.z += hv.z; gv.w += hv.w;
            }
            float4 xh, gn;
            xh.x = (xv.x - mu) * rstd; xh.y = (xv.y - mu) * rstd; xh.z = (xv.z - mu) * rstd; xh.w = (xv.w - mu) * rstd;
            gn = gv;
            if (non_lin) {
                gn.x *= gelu_grad_f(fmaf(xh.x, ga, be)); gn.y *= gelu_grad_f(fmaf(xh.y, ga, be));
                gn.z *= gelu_grad_f(fmaf(xh.z, ga, be)); gn.w *= gelu_grad_f(fmaf(xh.w, ga, be));
            }
            gn4[i] = gn;
            xh4[i] = xh;
            s1 += (gn.x + gn.y) + (gn.z + gn.w);
            s2 = fmaf(gn.x, xh.x, s2); s2 = fmaf(gn.y, xh.y, s2); s2 = fmaf(gn.z, xh.z, s2); s2 = fmaf(gn.w, xh.w, s2);
            if (++cnt == 16) { d1 += s1; d2 += s2; s1 = s2 = 0.f; cnt = 0; }
        }
    }
    for (int i = 4 * nv + threadIdx.x; i < n; i += kNormThreads) {
        const float xh = (xp[i] - mu) * rstd;
        const float up = hp ? gp[i] + hp[i] : gp[i];
        const float gn = non_lin ? up * gelu_grad_f(fmaf(xh, ga, be)) : up;
        gn_s[i] = gn;
        xh_s[i] = xh;
        s1 += gn;
        s2 = fmaf(gn, xh, s2);
        if (++cnt == 16) { d1 += s1; d2 += s2; s1 = s2 = 0.f; cnt = 0; }
    }
    d1 += s1; d2 += s2;
    const double b1 = block_sum(d1, sh);
    const double b2 = block_sum(d2, sh);
    if (threadIdx.x == 0) { mail[0] = b1; mail[1] = b2; }
    cluster.sync();
    double t1 = 0.0, t2 = 0.0;
    for (unsigned r = 0; r < cs; ++r) {
        const double* m = cluster.map_shared_rank(mail, r);
        t1 += m[0];
        t2 += m[1];
    }
    if (rank == 0 && threadIdx.x == 0) {
        atomicAdd(ggamma + c, (float)t2);
        atomicAdd(gbeta + c, (float)t1);
    }
    const float m1 = (float)(t1 / (double)L), m2 = (float)(t2 / (double)L);
    const float k = ga * rstd;
    float* op = g + p * L + lo;
    {
        const float4* gn4 = reinterpret_cast<const float4*>(gn_s);
        const float4* xh4 = reinterpret_cast<const float4*>(xh_s);
        float4* o4 = reinterpret_cast<float4*>(op);
        for (int i = threadIdx.x; i < nv; i += kNormThreads) {
            const float4 a = gn4[i], b = xh4[i];
            o4[i] = make_float4(k * (a.x - m1 - b.x * m2), k * (a.y - m1 - b.y * m2), k * (a.z - m1 - b.z * m2), k * (a.w - m1 - b.w * m2));
        }
    }
    for (int i = 4 * nv + threadIdx.x; i < n; i += kNormThreads) op[i] = k * (gn_s[i] - m1 - xh_s[i] * m2);
    cluster.sync();
}

// cluster size (1, 2, 4, 8) whose per-CTA slice of `bytes_per_elem * L` fits `budget` bytes of shared memory; 0 if none
inline int norm_cluster_size(long L, int bytes_per_elem, size_t budget, int* slice) {
    for (int cs = 1; cs <= 8; cs *= 2) {
        const long sl = ((L + cs - 1) / cs + 3) & ~3L;
        if ((size_t)sl * bytes_per_elem <= budget) { *slice = (int)sl; return cs; }
    }
    return 0;
}

template <typename Kern, typename... Args>
int launch_cluster(Kern kernel, long planes, int cs, size_t smem, cudaStream_t st, Args... args) {
    int rc = ensure_smem(kernel, smem);
    if (rc) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(planes * cs), 1, 1);
    cfg.blockDim = dim3(kNormThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, kernel, args...);
}
