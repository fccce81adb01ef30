// Projection backward on the tensor cores (warp-level mma.sync.m16n8k8 TF32 with the 3-term split that keeps fp32-grade
// accuracy: x = hi + lo, D += a_lo*b_hi + a_hi*b_lo + a_hi*b_hi).  Included by backend_cuda.cu after pixel_mlp.cuh.
//
// Same tile loop, staging and reductions as proj_bwd_kernel (pixel_mlp.cuh); the two large products of a tile move from
// the fp32 pipes (19 k + 18 k warp instructions) to ~4 k + ~4.5 k:
//     DIN[p][c] += sum_n D[n][p] * W1[n][c]      warp w: pixels 32w..32w+31 (2 m-tiles) x all channels, k = the chunk's 32 hidden units
//     dW1[n][c] += sum_p D[n][p] * IN[c][p]      warp w: the k-slice p = 32w..32w+31 of the tile, all 32 x CT outputs
// W1 is constant during the kernel and is split into tf32 hi / lo images once (row pitch CT+8: conflict-free B fragments);
// D and IN fragments are split on the fly.  With a single hidden chunk (hid <= 32: the Darcy and 3-D models) every warp
// keeps its dW1 partial sums in registers across ALL tiles of the persistent CTA and reduces them once at the end.
// Requires the fc1 pre-activations saved by the forward pass (pre_in).
//
// MEASURED (B200, Darcy 421^2, batch 32): parity-green on all projection cases but 5.14 ms against 3.53 ms for the fp32
// kernel: the 3072 mma.sync per tile issue at ~1 per 22 cycles per SM, i.e. legacy warp-level TF32 MMA runs below the
// fp32 FMA pipe on sm_100a.  Kept as an opt-in (UNO_B200_PROJ_MMA=1) record of that experiment; the tensor-core route
// for these products is tcgen05 (UMMA from shared memory into TMEM), as in the tc_*.cuh kernels.
#pragma once

__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__host__ __device__ inline int round32(int n) { return (n + 31) & ~31; }

template <int CT>
__global__ void __launch_bounds__(kPixTP, 1) proj_bwd_mma_kernel(const ProjK k, long ntiles, int nbuf) {
    extern __shared__ __align__(16) float psm[];
    constexpr int CTP = CT + 8, NT = CT / 8;
    const float** sbase = reinterpret_cast<const float**>(psm);
    float** gbase = reinterpret_cast<float**>(psm) + CT;
    long* sstride = reinterpret_cast<long*>(gbase + CT);
    const int H4 = round4(k.hid), OH4 = round4(k.out_ch * k.hid), H32 = round32(k.hid);
    uint32_t* W1hi = reinterpret_cast<uint32_t*>(sstride + CT);   // [H32][CTP] tf32 images of fc1.weight (zero rows past hid)
    uint32_t* W1lo = W1hi + (size_t)H32 * CTP;
    float* sW2 = reinterpret_cast<float*>(W1lo + (size_t)H32 * CTP);   // [out_ch][hid]
    float* accW1 = sW2 + OH4;                              // gradient accumulators
    float* accb1 = accW1 + (size_t)k.hid * CT;
    float* accW2 = accb1 + H4;
    float* accb2 = accW2 + OH4;                            // [4]
    float* D = accb2 + 4;                                  // [32][TPP]  dL/dpre1 of the current hidden chunk
    float* INbuf = D + (size_t)kProjHC * kPixTPP;          // nbuf x [CT][TPP]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g8 = lane >> 2, tig = lane & 3;
    proj_stage_tables<CT>(k, sbase, gbase, sstride);
    for (int i = tid; i < H32 * CTP; i += kPixTP) {
        const int n = i / CTP, c = i % CTP;
        uint32_t hi = 0u, lo = 0u;
        if (n < k.hid && c < k.ctot) tf32_split(__ldg(k.w1 + n * k.ctot + c), hi, lo);
        W1hi[i] = hi;
        W1lo[i] = lo;
    }
    for (int i = tid; i < k.out_ch * k.hid; i += kPixTP) sW2[i] = __ldg(k.w2 + i);
    for (int i = tid; i < k.hid * CT + H4 + OH4 + 4; i += kPixTP) accW1[i] = 0.f;
    __syncthreads();
    const PixGeom g = k.g;
    const long total = (long)k.batch * g.nraw;
    auto stage_tile = [&](long tile, float* dstbuf) {
        const long idx = tile * kPixTP + tid;
        const bool valid = idx < total;
        long b = 0, rp = 0, pp = 0;
        if (valid) raw_to_padded(g, idx, b, rp, pp);
        uint32_t dst = (uint32_t)__cvta_generic_to_shared(dstbuf + tid);
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
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int i = tid; i < (CT - k.ctot) * kPixTPP; i += kPixTP) {
        INbuf[(size_t)k.ctot * kPixTPP + i] = 0.f;
        if (nbuf == 2) INbuf[(size_t)(CT + k.ctot) * kPixTPP + i] = 0.f;
    }
    proj_zero_padding(k, gbase, sstride);
    const int tn = tid >> 6, tp = tid & 63;                // activation phase: hidden block, pixel lane (pixels tp + 64 q)
    const bool one_chunk = k.hid <= kProjHC;
    float wacc[2][NT][4];                                  // dW1 partial sums of this warp's pixel slices (registers across tiles)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) wacc[mt][nt][e] = 0.f;
    auto flush_wacc = [&](int ch0) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int n = ch0 + 16 * mt + g8 + 8 * (e >> 1), c = 8 * nt + 2 * tig + (e & 1);
                    if (n < k.hid) atomicAdd(accW1 + (size_t)n * CT + c, wacc[mt][nt][e]);
                    wacc[mt][nt][e] = 0.f;
                }
    };
    int cur = 0;
    if (nbuf == 2 && (long)blockIdx.x < ntiles) stage_tile(blockIdx.x, INbuf);
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long base = tile * kPixTP;
        float* IN = INbuf + (size_t)cur * CT * kPixTPP;
        if (nbuf == 2) {
            const long next = tile + gridDim.x;
            if (next < ntiles) {
                stage_tile(next, INbuf + (size_t)(cur ^ 1) * CT * kPixTPP);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            cur ^= 1;
        } else {
            stage_tile(tile, IN);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        bool vq[4];
        float go[4][kProjMaxOut];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long idx = base + tp + 64 * q;
            vq[q] = idx < total;
#pragma unroll
            for (int o = 0; o < kProjMaxOut; ++o) go[q][o] = (vq[q] && o < k.out_ch) ? __ldg(k.gout + idx * k.out_ch + o) : 0.f;
        }
        float dacc[2][NT][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) dacc[mt][nt][e] = 0.f;
        __syncthreads();
        for (int ch0 = 0; ch0 < k.hid; ch0 += kProjHC) {
            const int nn = min(kProjHC, k.hid - ch0);
            // ---- activation phase (fp32 pipes): D = (W2^T gout) * gelu'(pre), dW2 / db1 partial sums
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = 8 * tn + i;
                const int n = ch0 + j;
                const bool live = j < nn;
                const int nc = live ? n : k.hid - 1;
                const float* src = k.pre_in + (size_t)nc * total + base + tp;
                float w2v[kProjMaxOut];
#pragma unroll
                for (int o = 0; o < kProjMaxOut; ++o) w2v[o] = o < k.out_ch ? sW2[o * k.hid + nc] : 0.f;
                float sb = 0.f, sw[kProjMaxOut];
#pragma unroll
                for (int o = 0; o < kProjMaxOut; ++o) sw[o] = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float a, gp;
                    gelu_both(vq[q] ? __ldg(src + 64 * q) : 0.f, a, gp);
                    float s = 0.f;
#pragma unroll
                    for (int o = 0; o < kProjMaxOut; ++o) {
                        s = fmaf(go[q][o], w2v[o], s);
                        sw[o] = fmaf(go[q][o], a, sw[o]);
                    }
                    const float dp = live ? s * gp : 0.f;
                    D[(size_t)j * kPixTPP + tp + 64 * q] = dp;
                    sb += dp;
                }
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
            __syncthreads();
            // ---- tensor-core phase.  Fragment coordinates of m16n8k8: A (row g8 | g8+8, col tig | tig+4),
            //      B (k tig | tig+4, n g8), C (row g8 | g8+8, col 2 tig | 2 tig + 1)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                // DIN: A[m = pixel][k = hidden] = D[hidden][pixel];  B[k = hidden][n = channel] = W1[hidden][channel]
                uint32_t ahi[2][4], alo[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const float* d0 = D + (size_t)(8 * ks + tig) * kPixTPP + 32 * warp + 16 * mt + g8;
                    const float* d1 = d0 + 4 * kPixTPP;
                    tf32_split(d0[0], ahi[mt][0], alo[mt][0]);
                    tf32_split(d0[8], ahi[mt][1], alo[mt][1]);
                    tf32_split(d1[0], ahi[mt][2], alo[mt][2]);
                    tf32_split(d1[8], ahi[mt][3], alo[mt][3]);
                }
                const uint32_t* wh = W1hi + (size_t)(ch0 + 8 * ks + tig) * CTP + g8;
                const uint32_t* wl = W1lo + (size_t)(ch0 + 8 * ks + tig) * CTP + g8;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const uint32_t bh0 = wh[8 * nt], bh1 = wh[8 * nt + 4 * CTP], bl0 = wl[8 * nt], bl1 = wl[8 * nt + 4 * CTP];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        mma_tf32(dacc[mt][nt], alo[mt], bh0, bh1);
                        mma_tf32(dacc[mt][nt], ahi[mt], bl0, bl1);
                        mma_tf32(dacc[mt][nt], ahi[mt], bh0, bh1);
                    }
                }
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                // dW1: A[m = hidden][k = pixel] = D[hidden][pixel];  B[k = pixel][n = channel] = IN[channel][pixel]
                const int pk = 32 * warp + 8 * ks;
                uint32_t ahi[2][4], alo[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const float* d0 = D + (size_t)(16 * mt + g8) * kPixTPP + pk + tig;
                    const float* d1 = d0 + 8 * kPixTPP;
                    tf32_split(d0[0], ahi[mt][0], alo[mt][0]);
                    tf32_split(d1[0], ahi[mt][1], alo[mt][1]);
                    tf32_split(d0[4], ahi[mt][2], alo[mt][2]);
                    tf32_split(d1[4], ahi[mt][3], alo[mt][3]);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const float* x0 = IN + (size_t)(8 * nt + g8) * kPixTPP + pk + tig;
                    uint32_t bh0, bl0, bh1, bl1;
                    tf32_split(x0[0], bh0, bl0);
                    tf32_split(x0[4], bh1, bl1);
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        mma_tf32(wacc[mt][nt], alo[mt], bh0, bh1);
                        mma_tf32(wacc[mt][nt], ahi[mt], bl0, bl1);
                        mma_tf32(wacc[mt][nt], ahi[mt], bh0, bh1);
                    }
                }
            }
            if (!one_chunk) flush_wacc(ch0);       // several chunks: the register tile is reused by the next one
            __syncthreads();
        }
        // ---- input gradients of the tile: this thread holds pixels 32 warp + 16 mt + g8 + 8 h, channels 8 nt + 2 tig + e
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long idx = base + 32 * warp + 16 * mt + g8 + 8 * h;
                if (idx < total) {
                    long b, rp, pp;
                    raw_to_padded(g, idx, b, rp, pp);
                    long cur_stride = -1, off = 0;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int c = 8 * nt + 2 * tig + e;
                            if (c < k.ctot) {
                                float* gb = gbase[c];
                                const long st = sstride[c];
                                if (st != cur_stride) { cur_stride = st; off = b * st + pp; }
                                if (gb != nullptr) gb[off] = dacc[mt][nt][2 * h + e];
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
    if (one_chunk) flush_wacc(0);
    __syncthreads();
    for (int i = tid; i < k.hid * CT; i += kPixTP) {
        const int n = i / CT, c = i % CT;
        if (c < k.ctot) atomicAdd(k.gw1 + n * k.ctot + c, accW1[i]);
    }
    for (int i = tid; i < k.hid; i += kPixTP) atomicAdd(k.gb1 + i, accb1[i]);
    for (int i = tid; i < k.out_ch * k.hid; i += kPixTP) atomicAdd(k.gw2 + i, accW2[i]);
    for (int i = tid; i < k.out_ch; i += kPixTP) atomicAdd(k.gb2 + i, accb2[i]);
}

inline size_t proj_bwd_mma_smem(int CT, int hid, int out_ch, int nbuf) {
    return proj_table_bytes(CT) + sizeof(float) * ((size_t)2 * round32(hid) * (CT + 8) + (size_t)hid * CT + 2 * (size_t)round4(out_ch * hid) +
                                                   round4(hid) + 4 + (size_t)(kProjHC + nbuf * CT) * kPixTPP);
}
