// Projection backward with its two large products on the 5th-generation tensor cores (tcgen05 / UMMA, accumulators in
// TMEM).  Included by backend_cuda.cu inside namespace uno::{anonymous}, after pixel_mlp.cuh and tc_common.cuh.
//
// Per 128-pixel tile and per chunk of 32 hidden units:
//   activation (fp32 pipes)   D[n][p] = (W2^T gout)[n][p] * gelu'(pre[n][p])          pre = fc1 pre-activations kept by forward
//   DIN[p][c] (+)= sum_n D[n][p] * W1[n][c]      UMMA 128 (pixels) x 64 (channels) x 32,  3xTF32: hi*hi + hi*lo + lo*hi
//   dW1[c][n]  += sum_p IN[c][p] * D[n][p]       UMMA 128 x 32 x 128 with the hi and lo parts of IN STACKED as rows
//                                                0-63 / 64-127 of one A operand, so two MMAs per k-step (B = D_hi, D_lo)
//                                                give all four split products; the two row blocks are summed by the
//                                                atomics of the final flush.  Its accumulator stays in TMEM for the whole
//                                                kernel (one 32-column block per hidden chunk): split-K over all tiles of the CTA.
// Operands are K-major, no swizzle ("interleave": element (row, k) at (k>>2)*LBO + row*16 + (k&3)*4, SBO = 128).
// The tile's inputs arrive by LDGSTS into a raw buffer one tile ahead; a pass turns them into the stacked tf32 hi / lo
// image.  D is written by the activation phase straight into its two operand images (pixel-major for DIN, hidden-major
// for dW1).  One thread issues the MMAs; everybody waits on one mbarrier (tcgen05.commit) before the images are reused.
// Takes: 64-channel template, hid <= 64, saved pre-activations.
//
// History (measured on B200, Darcy 421^2, batch 32): a first version that ran these phases back to back behind block-wide
// barriers took 4.07 ms per launch against 3.4 ms for the fp32 kernel (pixel_mlp.cuh) -- the tensor-core work is ~1 k cycles
// of a ~27 k-cycle tile -- and was removed.  The warp-specialised kernel below (proj_bwd_tcp_kernel) keeps the barriers off
// the critical path: 1.6 ms, the default for the shipped shapes; other shapes run the fp32 kernel.
#pragma once

constexpr int kPtPix = 128;                          // pixels per tile
constexpr uint32_t kPtLboA = 128 * 16 + 16;          // 128-row operands (pixels x hidden, stacked channels x pixels)
constexpr uint32_t kPtLboW = 64 * 16;                // fc1 weight chunk: 64 rows (channels) x 32 k (hidden)
constexpr uint32_t kPtLboD = 32 * 16 + 16;           // D for dW1: 32 rows (hidden) x 128 k (pixels); +16 B keeps the scalar stores conflict-free
constexpr uint32_t kPtImgBytes = (kPtPix / 4) * kPtLboA;       // stacked IN image
constexpr uint32_t kPtA1Bytes = (kProjHC / 4) * kPtLboA;       // one D^T image (hi or lo)
constexpr uint32_t kPtDBytes = (kPtPix / 4) * kPtLboD;         // one D image (hi or lo)
constexpr uint32_t kPtWBytes = (kProjHC / 4) * kPtLboW;        // one fc1 chunk image (hi or lo)
constexpr int kPtMaxChunks = 2;                                // hid <= 64
constexpr uint32_t kPtTmemCols = 128;                          // 64 (DIN) + 32 per chunk (dW1)

// =====================================================================================================
// Warp-specialised form (hid <= 32: one hidden chunk).  The same tensor-core products as above, but the fp32 phases of a
// tile are split between two groups of four warps that only meet at two mbarriers:
//   group A (warps 0-3, thread = pixel)   INPUT(t): its pixel's 64 channels, copied by LDGSTS one tile ahead, become the
//                                         stacked tf32 image; then EPI(t-1): its TMEM lane (= pixel) of the DIN
//                                         accumulator -> gsrc; then it issues the copies of tile t+1
//   group B (warps 4-7, thread = pixel)   ACT(t): 32 pre-activations (prefetched into registers one tile ahead) ->
//                                         D in both operand images; the sums over pixels (db1, dW2, db2) stay in
//                                         registers for the whole kernel
//   thread 128                            waits for all 256 arrivals (bar_full), issues the 12 + 32 MMAs of the tile and
//                                         commits to bar_mma, which both groups wait on before touching the images again
// No block-wide barrier in the loop; the only exposed latency per tile is the tensor core's.
// =====================================================================================================
constexpr int kTcpThreads = 512;      // 8 warps per group: two threads per pixel (half the channels / hidden units each)
__global__ void __launch_bounds__(kTcpThreads, 1) proj_bwd_tcp_kernel(const ProjK k, long ntiles) {
    extern __shared__ __align__(128) uint8_t tsm[];
    constexpr int CT = 64;
    uint8_t* p0 = tsm + ((128u - (tc::smem_u32(tsm) & 127u)) & 127u);
    uint8_t* IMG = p0;
    uint8_t* A1hi = IMG + kPtImgBytes;
    uint8_t* A1lo = A1hi + kPtA1Bytes;
    uint8_t* Dhi = A1lo + kPtA1Bytes;
    uint8_t* Dlo = Dhi + kPtDBytes;
    uint8_t* Wimg = Dlo + kPtDBytes;                     // [hi | lo] fc1 image (one chunk)
    float* RAW = reinterpret_cast<float*>(Wimg + (size_t)2 * kPtWBytes);
    const float** sbase = reinterpret_cast<const float**>(RAW + 64 * kPtPix);
    float** gbase = reinterpret_cast<float**>(const_cast<float**>(sbase) + CT);
    long* sstride = reinterpret_cast<long*>(gbase + CT);
    const int OH4 = round4(k.out_ch * k.hid);
    float* sW2 = reinterpret_cast<float*>(sstride + CT);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(sW2 + OH4);
    uint64_t* bar_mma = bar_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool groupA = tid < 256;
    const int px = tid & 127;                            // this thread's pixel of every tile
    const int half = (tid >> 7) & 1;                     // group A: channels 32 half .. +32; group B: hidden units 16 half .. +16

    proj_stage_tables<CT>(k, sbase, gbase, sstride);
    for (uint32_t i = tid; i < (kPtImgBytes + 2 * kPtA1Bytes + 2 * kPtDBytes) / 16; i += kTcpThreads)
        reinterpret_cast<float4*>(IMG)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < kProjHC * CT; i += kTcpThreads) {
        const int c = i % CT, j = i / CT;
        float hi = 0.f, lo = 0.f;
        if (j < k.hid && c < k.ctot) tc::split_tf32(__ldg(k.w1 + j * k.ctot + c), hi, lo);
        uint8_t* d = Wimg + (uint32_t)(j >> 2) * kPtLboW + (uint32_t)c * 16 + (uint32_t)(j & 3) * 4;
        *reinterpret_cast<float*>(d) = hi;
        *reinterpret_cast<float*>(d + kPtWBytes) = lo;
    }
    for (int i = tid; i < k.out_ch * k.hid; i += kTcpThreads) sW2[i] = __ldg(k.w2 + i);
    if (tid == 0) {
        tc::mbar_init(bar_full, kTcpThreads);
        tc::mbar_init(bar_mma, 1);
        tc::fence_barrier_init();
    }
    constexpr uint32_t kCols = 256;       // DIN accumulators at columns 0 / 64 (alternating tiles), dW1 at 128
    if (warp == 0) tc::tmem_alloc(tmem_slot, kCols);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const PixGeom g = k.g;
    const long total = (long)k.batch * g.nraw;
    proj_zero_padding(k, gbase, sstride);
    const long step_tiles = gridDim.x;
    long it = 0;
    if (groupA) {
        // ------------------------------------------------------------------ group A: input image + gradient stores
        long b_prev = 0, pp_prev = 0, b_cur = 0, pp_cur = 0;
        bool v_prev = false, v_cur = false;
        auto stage_raw = [&](long tile, long& b, long& pp, bool& valid) {
            const long idx = tile * kPtPix + px;
            valid = idx < total;
            long rp = 0;
            b = 0; pp = 0;
            if (valid) raw_to_padded(g, idx, b, rp, pp);
            uint32_t dst = tc::smem_u32(RAW + (32 * half) * kPtPix + px);
#pragma unroll 4
            for (int c = 32 * half; c < 32 * half + 32; ++c) {
                const bool on = valid && c < k.ctot;
                const float* src = on ? sbase[c] + b * sstride[c] + pp : k.w1;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(on ? 4 : 0) : "memory");
                dst += (uint32_t)(kPtPix * 4);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // every source a multiple of 16 channels wide (the shipped models: 32 + 32): a 16-column TMEM read never straddles
        // two sources, so its stores run on one pointer advanced by the plane size
        const bool src16 = (k.src_ch[0] % 16 == 0) && (k.src_ch[1] % 16 == 0) && (k.src_ch[2] % 16 == 0) && (k.src_ch[3] % 16 == 0);
        auto store_gradients = [&](long b, long pp, bool valid, uint32_t acc_col) {
            long cur_stride = -1, off = 0;
#pragma unroll
            for (int cc = 0; cc < 32; cc += 16) {
                const int c0 = 32 * half + cc;
                uint32_t r[16];
                tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + acc_col + (uint32_t)c0, r);
                tc::tmem_ld_wait();
                if (!valid || c0 >= k.ctot) continue;
                if (src16) {
                    float* gb = gbase[c0];
                    if (gb != nullptr) {
                        gb += b * sstride[c0] + pp;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            *gb = __uint_as_float(r[j]);
                            gb += g.npad;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int c = c0 + j;
                        if (c < k.ctot) {
                            float* gb = gbase[c];
                            const long st = sstride[c];
                            if (st != cur_stride) { cur_stride = st; off = b * st + pp; }
                            if (gb != nullptr) gb[off] = __uint_as_float(r[j]);
                        }
                    }
                }
            }
        };
        if ((long)blockIdx.x < ntiles) stage_raw(blockIdx.x, b_cur, pp_cur, v_cur);
        for (long tile = blockIdx.x; tile < ntiles; tile += step_tiles, ++it) {
            if (it > 0) {                                // MMA(t-1) complete: image free, DIN accumulator final
                tc::mbar_wait(bar_mma, (uint32_t)(it - 1) & 1u);
                tc::tc_fence_after();
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");        // this thread's own copies of tile t
            {
                uint8_t* d = IMG + (uint32_t)(px >> 2) * kPtLboA + (uint32_t)(px & 3) * 4;
#pragma unroll 8
                for (int c = 32 * half; c < 32 * half + 32; ++c) {
                    float hi, lo;
                    tc::split_tf32(RAW[c * kPtPix + px], hi, lo);
                    *reinterpret_cast<float*>(d + (uint32_t)c * 16) = hi;
                    *reinterpret_cast<float*>(d + (uint32_t)(64 + c) * 16) = lo;
                }
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(bar_full);                   // the image of tile t is ready: everything below is off the critical path
            const long b_t = b_cur, pp_t = pp_cur;
            const bool v_t = v_cur;
            if (tile + step_tiles < ntiles) stage_raw(tile + step_tiles, b_cur, pp_cur, v_cur);   // this thread's RAW column is free again
            // EPI(t-1) reads the accumulator MMA(t) does not write (they alternate); it finishes before this thread arrives
            // for tile t+1, hence before MMA(t+1) reuses that accumulator
            if (it > 0) {
                store_gradients(b_prev, pp_prev, v_prev, (uint32_t)((it - 1) & 1) * 64u);
                tc::tc_fence_before();
            }
            b_prev = b_t; pp_prev = pp_t; v_prev = v_t;
        }
        if (it > 0) {
            tc::mbar_wait(bar_mma, (uint32_t)(it - 1) & 1u);
            tc::tc_fence_after();
            store_gradients(b_prev, pp_prev, v_prev, (uint32_t)((it - 1) & 1) * 64u);
            // dW1: TMEM lanes 0-63 hold the hi-part rows (channel = lane), 64-127 the lo-part rows: both add into gw1
            const int c = (32 * (warp & 3) + lane) & 63;
            {
                const int c0 = 16 * half;
                uint32_t r[16];
                tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + 128u + (uint32_t)c0, r);
                tc::tmem_ld_wait();
                if (c < k.ctot) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < k.hid) atomicAdd(k.gw1 + (c0 + j) * k.ctot + c, __uint_as_float(r[j]));
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ group B: activation backward + MMA issue
        const uint32_t idesc_din = tc::make_idesc_tf32(128, 64, 0, 0), idesc_dw = tc::make_idesc_tf32(128, kProjHC, 0, 0);
        const uint64_t d_img = tc::make_smem_desc(tc::smem_u32(IMG), kPtLboA, 128), d_a1hi = tc::make_smem_desc(tc::smem_u32(A1hi), kPtLboA, 128),
                       d_dhi = tc::make_smem_desc(tc::smem_u32(Dhi), kPtLboD, 128), d_w = tc::make_smem_desc(tc::smem_u32(Wimg), kPtLboW, 128);
        const uint32_t a1lo_off = (tc::smem_u32(A1lo) - tc::smem_u32(A1hi)) >> 4, dlo_off = (tc::smem_u32(Dlo) - tc::smem_u32(Dhi)) >> 4;
        // (single output channel: every shipped model projects to one field; other widths run the fp32 kernel)
        constexpr int NH = kProjHC / 2;                       // hidden units per thread
        const int n0 = NH * half;
        float pre[NH], go = 0.f;
        float s_b1[NH], s_w2[NH], s_b2 = 0.f;                 // sums over this thread's pixels, whole kernel
#pragma unroll
        for (int n = 0; n < NH; ++n) { s_b1[n] = 0.f; s_w2[n] = 0.f; }
        auto prefetch = [&](long tile) {
            const long idx = tile * kPtPix + px;
            const bool valid = tile < ntiles && idx < total;
#pragma unroll
            for (int n = 0; n < NH; ++n) pre[n] = (valid && n0 + n < k.hid) ? __ldg(k.pre_in + (size_t)(n0 + n) * total + idx) : 0.f;
            go = valid ? __ldg(k.gout + idx) : 0.f;
        };
        prefetch(blockIdx.x);
        for (long tile = blockIdx.x; tile < ntiles; tile += step_tiles, ++it) {
            float hi[NH], lo[NH];
#pragma unroll
            for (int n = 0; n < NH; ++n) {
                float a, gp;
                gelu_both(pre[n], a, gp);
                const float w2 = n0 + n < k.hid ? sW2[n0 + n] : 0.f;
                s_w2[n] = fmaf(go, a, s_w2[n]);
                const float dp = go * w2 * gp;
                s_b1[n] += dp;
                tc::split_tf32(dp, hi[n], lo[n]);
            }
            if (half == 0) s_b2 += go;
            prefetch(tile + step_tiles);                 // next tile's loads fly while this tile's images are written
            if (it > 0) {                                // MMA(t-1) complete: the D images may be rewritten
                tc::mbar_wait(bar_mma, (uint32_t)(it - 1) & 1u);
                tc::tc_fence_after();
            }
#pragma unroll
            for (int m = 0; m < NH / 4; ++m) {
                const uint32_t oa = (uint32_t)(n0 / 4 + m) * kPtLboA + (uint32_t)px * 16;
                *reinterpret_cast<float4*>(A1hi + oa) = make_float4(hi[4 * m], hi[4 * m + 1], hi[4 * m + 2], hi[4 * m + 3]);
                *reinterpret_cast<float4*>(A1lo + oa) = make_float4(lo[4 * m], lo[4 * m + 1], lo[4 * m + 2], lo[4 * m + 3]);
            }
            const uint32_t od = (uint32_t)(px >> 2) * kPtLboD + (uint32_t)(px & 3) * 4 + (uint32_t)n0 * 16;
#pragma unroll
            for (int n = 0; n < NH; ++n) {
                *reinterpret_cast<float*>(Dhi + od + n * 16) = hi[n];
                *reinterpret_cast<float*>(Dlo + od + n * 16) = lo[n];
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(bar_full);
            if (warp == 8) {                             // warp-uniform; one elected lane issues, descriptors are only advanced
                if (tc::elect_one()) {
                    tc::mbar_wait(bar_full, (uint32_t)it & 1u);
                    tc::tc_fence_after();
                    const uint32_t din = tmem_base + (uint32_t)(it & 1) * 64u;
                    uint64_t da = d_a1hi, db = d_w;
#pragma unroll
                    for (int ks = 0; ks < kProjHC / 8; ++ks) {
                        tc::mma_tf32(din, da, db, idesc_din, ks ? 1u : 0u);
                        tc::mma_tf32(din, da, db + (kPtWBytes >> 4), idesc_din, 1u);
                        tc::mma_tf32(din, da + a1lo_off, db, idesc_din, 1u);
                        da += (2 * kPtLboA) >> 4;
                        db += (2 * kPtLboW) >> 4;
                    }
                    uint64_t di = d_img, dd = d_dhi;
#pragma unroll 4
                    for (int ks = 0; ks < kPtPix / 8; ++ks) {
                        tc::mma_tf32(tmem_base + 128u, di, dd, idesc_dw, (it | ks) ? 1u : 0u);
                        tc::mma_tf32(tmem_base + 128u, di, dd + dlo_off, idesc_dw, 1u);
                        di += (2 * kPtLboA) >> 4;
                        dd += (2 * kPtLboD) >> 4;
                    }
                    tc::tc_commit(bar_mma);
                }
            }
            __syncwarp();
        }
        // sums over pixels: reduce over the warp, one atomic per value per warp
#pragma unroll
        for (int n = 0; n < NH; ++n) {
            const float v = warp_sum(s_b1[n]);
            const float w = warp_sum(s_w2[n]);
            if (lane == 0 && n0 + n < k.hid) {
                atomicAdd(k.gb1 + n0 + n, v);
                atomicAdd(k.gw2 + n0 + n, w);
            }
        }
        {
            const float v = warp_sum(s_b2);
            if (lane == 0) atomicAdd(k.gb2, v);
        }
        if (it > 0 && tid == 256) tc::mbar_wait(bar_mma, (uint32_t)(it - 1) & 1u);   // nothing of this CTA may still be in flight
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, kCols);
}

__host__ __device__ inline size_t proj_bwd_tcp_smem(int hid, int out_ch) {
    return 1024 + kPtImgBytes + 2 * kPtA1Bytes + 2 * kPtDBytes + (size_t)2 * kPtWBytes + 64 * kPtPix * 4 + proj_table_bytes(64) +
           sizeof(float) * round4(out_ch * hid) + 64;
}

// =====================================================================================================
// Projection FORWARD on tcgen05:   out = W_2 * gelu( W_1 * crop(cat(src..)) + b_1 ) + b_2
//   PRE[128 pixels, hid] = IN[128 pixels, 64 channels] * W_1^T     UMMA 128 x N_t x 64, 3xTF32 (24 MMAs per tile)
// The fp32 kernel (pixel_mlp.cuh proj_fwd_kernel) is bound by its FMA / shared-memory instruction stream (ncu: FMA pipe 41 %,
// 0.31 of HBM); here the product costs the SIMT pipes nothing and what is left per pixel is one load, one tf32 split and a
// quarter of a 16-byte shared store per input value, and bias + GELU + the fc2 dot product per hidden unit.
//   warps 0-15  loaders: thread = (pixel of the tile, quarter of the 64 channels); its 16 values of the tiles t+1 and t+2 are in
//               flight in registers while those of tile t are split into the K-major hi / lo images of stage t % 2
//   warps 16-23 epilogue: thread = (pixel = TMEM lane, every other 16-column chunk of the hidden units); PRE + b_1 -> optional
//               pre_out (kept for backward), GELU, partial fc2 sum; the two partial sums of a pixel meet in shared memory
//               (a first version with four epilogue warps and run-time out_ch / hid predicates in the unrolled loop was bound by
//               exactly those warps: ncu showed them busy 85 % of the time with the sixteen loader warps spinning on `empty`)
//   warp  20    one elected lane issues the MMAs of a tile into one of two TMEM accumulators and commits to the mbarriers that
//               free the operand stage and publish the accumulator
// Takes: at most 64 input channels, hid <= 64, out_ch <= 4.
// =====================================================================================================
constexpr int kPfLoadWarps = 16, kPfEpiWarps = 8;
constexpr int kPfThreads = (kPfLoadWarps + kPfEpiWarps + 1) * 32;
constexpr uint32_t kPfImgHalf = 16 * kPtLboA;          // one operand image (hi or lo): 16 channel groups x 128 pixels x 16 B (+ pad)
constexpr uint32_t kPfStage = 2 * kPfImgHalf;

__host__ __device__ inline size_t proj_fwd_tc_smem(int N_t, int hid, int out_ch) {
    (void)hid; (void)out_ch;
    return 1024 + (size_t)2 * kPfStage + (size_t)2 * N_t * 64 * 4 + proj_table_bytes(64) +
           sizeof(float) * (64 + kProjMaxOut * 64 + 4 + 2 * kPtPix * kProjMaxOut) + 16 * 8;
}

// NOUT = 1: one output channel (every shipped model); NOUT = 4: out_ch <= 4 at run time.  PRE: pre_out wanted.
template <int NOUT, bool PRE>
__global__ void __launch_bounds__(kPfThreads, 1) proj_fwd_tc_kernel(const ProjK k, long ntiles, int N_t) {
    extern __shared__ __align__(128) uint8_t tsm[];
    uint8_t* IMG = tsm + ((128u - (tc::smem_u32(tsm) & 127u)) & 127u);   // [2 stages][hi | lo]
    uint8_t* Wimg = IMG + 2 * kPfStage;                                  // [hi | lo] fc1 weights, K-major, N_t rows
    const uint32_t w_half = (uint32_t)N_t * 64 * 4;
    const float** sbase = reinterpret_cast<const float**>(Wimg + 2 * w_half);
    float** gbase = reinterpret_cast<float**>(const_cast<float**>(sbase) + 64);
    long* sstride = reinterpret_cast<long*>(gbase + 64);
    float* sb1 = reinterpret_cast<float*>(sstride + 64);      // [64], zero beyond hid
    float* sW2 = sb1 + 64;                                    // [kProjMaxOut][64], zero beyond hid / out_ch
    float* sb2 = sW2 + kProjMaxOut * 64;                      // [4]
    float* osum = sb2 + 4;                                    // [2 stages][128 pixels][kProjMaxOut] partial fc2 sums of the odd chunks
    uint64_t* bars = reinterpret_cast<uint64_t*>(osum + 2 * kPtPix * kProjMaxOut);
    uint64_t* full = bars;          // [2] loaders -> mma
    uint64_t* empty = bars + 2;     // [2] mma -> loaders
    uint64_t* d_full = bars + 4;    // [2] mma -> epilogue
    uint64_t* d_empty = bars + 6;   // [2] epilogue -> mma
    uint64_t* o_full = bars + 8;    // [2] odd-chunk epilogue warps -> even-chunk ones
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    proj_stage_tables<64>(k, sbase, gbase, sstride);
    for (int i = tid; i < N_t * 64; i += kPfThreads) {
        const int c = i & 63, n = i >> 6;
        float hi = 0.f, lo = 0.f;
        if (n < k.hid && c < k.ctot) {
            const float v = __ldg(k.w1 + n * k.ctot + c);
            const uint32_t u = __float_as_uint(v);
            hi = __uint_as_float((u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u);   // weights: hi rounded to nearest tf32
            lo = v - hi;
        }
        uint8_t* d = Wimg + (uint32_t)(c >> 2) * (uint32_t)(N_t * 16) + (uint32_t)n * 16 + (uint32_t)(c & 3) * 4;
        *reinterpret_cast<float*>(d) = hi;
        *reinterpret_cast<float*>(d + w_half) = lo;
    }
    for (int i = tid; i < 64; i += kPfThreads) sb1[i] = i < k.hid ? __ldg(k.b1 + i) : 0.f;
    for (int i = tid; i < kProjMaxOut * 64; i += kPfThreads) {
        const int u = i >> 6, n = i & 63;
        sW2[i] = (u < k.out_ch && n < k.hid) ? __ldg(k.w2 + u * k.hid + n) : 0.f;
    }
    if (tid < 4) sb2[tid] = tid < k.out_ch ? __ldg(k.b2 + tid) : 0.f;
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&full[s], kPfLoadWarps * 32);
            tc::mbar_init(&empty[s], 1);
            tc::mbar_init(&d_full[s], 1);
            tc::mbar_init(&d_empty[s], kPfEpiWarps * 32);
            tc::mbar_init(&o_full[s], kPfEpiWarps * 16);
        }
        tc::fence_barrier_init();
    }
    constexpr int kMmaWarp = kPfLoadWarps + kPfEpiWarps;
    constexpr uint32_t kCols = 128;                       // two accumulators of up to 64 columns
    if (warp == kMmaWarp) tc::tmem_alloc(tmem_slot, kCols);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const PixGeom g = k.g;
    const long total = (long)k.batch * g.nraw;
    const long step_tiles = gridDim.x;

    if (warp < kPfLoadWarps) {
        // ------------------------------------------------------------------ loaders
        const int px = tid & 127, qd = tid >> 7;
        const int c0 = 16 * qd;
        // all 16 channels of this thread inside one source: one pointer advanced by the plane size
        const bool contig = c0 + 15 < k.ctot && sbase[c0 + 15] == sbase[c0] + 15 * g.npad && sstride[c0 + 15] == sstride[c0];
        float ring[3][16];
        auto issue = [&](long tile, float (&v)[16]) {
            const long idx = tile * kPtPix + px;
            const bool valid = tile < ntiles && idx < total;
            long b = 0, rp = 0, pp = 0;
            if (valid) raw_to_padded(g, idx, b, rp, pp);
            if (contig) {
                const float* src = sbase[c0] + b * sstride[c0] + pp;
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = valid ? __ldg(src + (long)j * g.npad) : 0.f;
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const bool on = valid && c0 + j < k.ctot;
                    v[j] = on ? __ldg(sbase[c0 + j] + b * sstride[c0 + j] + pp) : 0.f;
                }
            }
        };
        long it = 0;
        auto process = [&](const float (&v)[16]) {
            const int s = (int)(it & 1);
            tc::mbar_wait(&empty[s], (((uint32_t)(it >> 1)) & 1u) ^ 1u);
            uint8_t* d = IMG + (uint32_t)s * kPfStage + (uint32_t)(4 * qd) * kPtLboA + (uint32_t)px * 16;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                float4 hi, lo;
                tc::split_tf32(v[4 * m + 0], hi.x, lo.x);
                tc::split_tf32(v[4 * m + 1], hi.y, lo.y);
                tc::split_tf32(v[4 * m + 2], hi.z, lo.z);
                tc::split_tf32(v[4 * m + 3], hi.w, lo.w);
                *reinterpret_cast<float4*>(d + (uint32_t)m * kPtLboA) = hi;
                *reinterpret_cast<float4*>(d + kPfImgHalf + (uint32_t)m * kPtLboA) = lo;
            }
            tc::fence_proxy_async();
            tc::mbar_arrive(&full[s]);
            ++it;
        };
        long tile = blockIdx.x;
        issue(tile, ring[0]);
        issue(tile + step_tiles, ring[1]);
        for (; tile < ntiles; tile += 3 * step_tiles) {
            issue(tile + 2 * step_tiles, ring[2]);
            process(ring[0]);
            if (tile + step_tiles >= ntiles) break;
            issue(tile + 3 * step_tiles, ring[0]);
            process(ring[1]);
            if (tile + 2 * step_tiles >= ntiles) break;
            issue(tile + 4 * step_tiles, ring[1]);
            process(ring[2]);
        }
    } else if (warp < kMmaWarp) {
        // ------------------------------------------------------------------ epilogue: thread = (pixel = TMEM lane, chunk parity)
        const int q = warp & 3, par = (warp - kPfLoadWarps) >> 2;
        const int px = q * 32 + lane;
        long it = 0;
        for (long tile = blockIdx.x; tile < ntiles; tile += step_tiles, ++it) {
            const int s = (int)(it & 1);
            const uint32_t ph = ((uint32_t)(it >> 1)) & 1u;
            tc::mbar_wait_relaxed(&d_full[s], ph);
            tc::tc_fence_after();
            const long idx = tile * kPtPix + px;
            const bool valid = idx < total;
            float o[NOUT];
#pragma unroll
            for (int u = 0; u < NOUT; ++u) o[u] = 0.f;
            const uint32_t t_base = tmem_base + (uint32_t)s * 64u + ((uint32_t)(q * 32) << 16);
            for (int n0 = 16 * par; n0 < k.hid; n0 += 32) {
                uint32_t r[16];
                tc::tmem_ld_32x32b_x16(t_base + (uint32_t)n0, r);
                tc::tmem_ld_wait();
                float* pre = PRE ? k.pre_out + (size_t)n0 * total + (valid ? idx : 0) : nullptr;
                const bool whole = n0 + 16 <= k.hid;          // warp-uniform
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float v = __uint_as_float(r[j]) + sb1[n0 + j];
                    if (PRE) {
                        if (valid && (whole || n0 + j < k.hid)) *pre = v;
                        pre += total;
                    }
                    const float a = gelu_act(v);              // padded units: weights, bias and fc2 row are zero
#pragma unroll
                    for (int u = 0; u < NOUT; ++u) o[u] = fmaf(sW2[u * 64 + n0 + j], a, o[u]);
                }
            }
            tc::tc_fence_before();
            float* os = osum + ((size_t)s * kPtPix + px) * kProjMaxOut;
            if (par) {
#pragma unroll
                for (int u = 0; u < NOUT; ++u) os[u] = o[u];
                tc::mbar_arrive(&o_full[s]);
                tc::mbar_arrive(&d_empty[s]);
            } else {
                tc::mbar_wait(&o_full[s], ph);
#pragma unroll
                for (int u = 0; u < NOUT; ++u) o[u] += os[u] + sb2[u];
                tc::mbar_arrive(&d_empty[s]);                 // after the read: the odd warps rewrite osum[s] two tiles later
                if (valid) {
#pragma unroll
                    for (int u = 0; u < NOUT; ++u)
                        if (NOUT == 1 || u < k.out_ch) k.out[idx * k.out_ch + u] = o[u];
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = tc::make_idesc_tf32(128, N_t, 0, 0);
        const uint32_t lbo_w = (uint32_t)N_t * 16;
        const uint64_t a0 = tc::make_smem_desc(tc::smem_u32(IMG), kPtLboA, 128);
        const uint64_t b0 = tc::make_smem_desc(tc::smem_u32(Wimg), lbo_w, 128);
        const uint32_t a_lo = kPfImgHalf >> 4, b_lo = w_half >> 4, a_step = (2 * kPtLboA) >> 4, b_step = (2 * lbo_w) >> 4;
        const int nks = (k.ctot + 7) / 8;
        long it = 0;
        for (long tile = blockIdx.x; tile < ntiles; tile += step_tiles, ++it) {
            const int s = (int)(it & 1);
            const uint32_t ph = ((uint32_t)(it >> 1)) & 1u;
            tc::mbar_wait(&full[s], ph);
            tc::mbar_wait(&d_empty[s], ph ^ 1u);
            tc::tc_fence_after();
            if (tc::elect_one()) {
                const uint32_t d_tmem = tmem_base + (uint32_t)s * 64u;
                uint64_t da = a0 + (uint64_t)((uint32_t)s * (kPfStage >> 4)), db = b0;
                for (int ks = 0; ks < nks; ++ks) {
                    tc::mma_tf32(d_tmem, da, db, idesc, ks ? 1u : 0u);
                    tc::mma_tf32(d_tmem, da, db + b_lo, idesc, 1u);
                    tc::mma_tf32(d_tmem, da + a_lo, db, idesc, 1u);
                    da += a_step;
                    db += b_step;
                }
                tc::tc_commit(&empty[s]);
                tc::tc_commit(&d_full[s]);
            }
            __syncwarp();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tc::tmem_dealloc(tmem_base, kCols);
}
