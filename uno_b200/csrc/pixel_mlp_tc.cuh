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
// STATUS (measured on B200, Darcy 421^2, batch 32): this first, barrier-synchronised version (proj_bwd_tc_kernel) is
// parity-green but takes 4.07 ms per launch against 3.4 ms for the fp32 kernel (pixel_mlp.cuh): the tensor-core work is ~1 k
// cycles of a ~27 k-cycle tile, the rest is the fp32 phases running back to back with five block-wide barriers per tile.
// It is opt-in (UNO_B200_PROJ_TC=1) and kept for hid in (32, 64].  The warp-specialised kernel at the end of this file
// (proj_bwd_tcp_kernel) removes the barriers from the critical path: 2.1 ms, the default for the shipped shapes.
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

__host__ __device__ inline size_t proj_bwd_tc_smem(int hid, int out_ch) {
    const int nch = (hid + kProjHC - 1) / kProjHC;
    return 1024 /* alignment slack */ + kPtImgBytes + 2 * kPtA1Bytes + 2 * kPtDBytes + (size_t)2 * nch * kPtWBytes + 64 * kPtPix * 4 /* raw */ +
           proj_table_bytes(64) + sizeof(float) * (2 * round4(out_ch * hid) + round4(hid) + 4) + 64;
}

__global__ void __launch_bounds__(256, 1) proj_bwd_tc_kernel(const ProjK k, long ntiles) {
    extern __shared__ __align__(128) uint8_t tsm[];
    constexpr int CT = 64;
    uint8_t* p0 = tsm + ((128u - (tc::smem_u32(tsm) & 127u)) & 127u);
    uint8_t* IMG = p0;                                   // stacked IN image: rows 0-63 tf32(x), rows 64-127 x - tf32(x)
    uint8_t* A1hi = IMG + kPtImgBytes;                   // D^T (pixels x hidden)
    uint8_t* A1lo = A1hi + kPtA1Bytes;
    uint8_t* Dhi = A1lo + kPtA1Bytes;                    // D (hidden x pixels)
    uint8_t* Dlo = Dhi + kPtDBytes;
    const int nchunks = (k.hid + kProjHC - 1) / kProjHC;
    uint8_t* Wimg = Dlo + kPtDBytes;                     // per chunk: [hi | lo] fc1 images
    float* RAW = reinterpret_cast<float*>(Wimg + (size_t)2 * nchunks * kPtWBytes);   // [64][128] inputs of the NEXT tile
    const float** sbase = reinterpret_cast<const float**>(RAW + 64 * kPtPix);
    float** gbase = reinterpret_cast<float**>(const_cast<float**>(sbase) + CT);
    long* sstride = reinterpret_cast<long*>(gbase + CT);
    const int H4 = round4(k.hid), OH4 = round4(k.out_ch * k.hid);
    float* sW2 = reinterpret_cast<float*>(sstride + CT);
    float* accb1 = sW2 + OH4;
    float* accW2 = accb1 + H4;
    float* accb2 = accW2 + OH4;
    uint64_t* bar = reinterpret_cast<uint64_t*>(accb2 + 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    proj_stage_tables<CT>(k, sbase, gbase, sstride);
    // zero every operand image once: rows of hidden units past hid, channels past ctot and pixels past the end stay zero
    for (uint32_t i = tid; i < (kPtImgBytes + 2 * kPtA1Bytes + 2 * kPtDBytes) / 16; i += 256)
        reinterpret_cast<float4*>(IMG)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // fc1 weight images: chunk ch, element (row = channel c, k = hidden j) = W1[32 ch + j][c]
    for (int i = tid; i < nchunks * kProjHC * CT; i += 256) {
        const int c = i % CT, j = (i / CT) % kProjHC, ch = i / (CT * kProjHC);
        const int n = ch * kProjHC + j;
        float hi = 0.f, lo = 0.f;
        if (n < k.hid && c < k.ctot) tc::split_tf32(__ldg(k.w1 + n * k.ctot + c), hi, lo);
        uint8_t* d = Wimg + (size_t)2 * ch * kPtWBytes + (uint32_t)(j >> 2) * kPtLboW + (uint32_t)c * 16 + (uint32_t)(j & 3) * 4;
        *reinterpret_cast<float*>(d) = hi;
        *reinterpret_cast<float*>(d + kPtWBytes) = lo;
    }
    for (int i = tid; i < k.out_ch * k.hid; i += 256) sW2[i] = __ldg(k.w2 + i);
    for (int i = tid; i < H4 + OH4 + 4; i += 256) accb1[i] = 0.f;
    if (tid == 0) {
        tc::mbar_init(bar, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(tmem_slot, kPtTmemCols);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const PixGeom g = k.g;
    const long total = (long)k.batch * g.nraw;
    proj_zero_padding(k, gbase, sstride);
    // LDGSTS of a tile's inputs into RAW[c][p]: threads 0-127 copy one pixel each, all channels (zero-filled past the end)
    auto stage_raw = [&](long tile) {
        if (tid < kPtPix) {
            const long idx = tile * kPtPix + tid;
            const bool valid = idx < total;
            long b = 0, rp = 0, pp = 0;
            if (valid) raw_to_padded(g, idx, b, rp, pp);
            uint32_t dst = tc::smem_u32(RAW + tid);
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
                        dst += (uint32_t)(kPtPix * 4);
                        src += step;
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const uint32_t idesc_din = tc::make_idesc_tf32(128, 64, 0, 0), idesc_dw = tc::make_idesc_tf32(128, kProjHC, 0, 0);
    const uint32_t img_a = tc::smem_u32(IMG), a1hi_a = tc::smem_u32(A1hi), a1lo_a = tc::smem_u32(A1lo), dhi_a = tc::smem_u32(Dhi),
                   dlo_a = tc::smem_u32(Dlo), w_a = tc::smem_u32(Wimg);
    const int tn = tid >> 6, tp = tid & 63;              // activation phase: hidden block of 8, pixels tp and tp + 64
    uint32_t phase = 0;
    long it = 0;
    if ((long)blockIdx.x < ntiles) stage_raw(blockIdx.x);
    for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const long base = tile * kPtPix;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                 // RAW complete; TMEM / images of the previous tile are free
        // ---- raw inputs -> stacked tf32 image (rows c: hi, rows 64 + c: lo); thread: pixel tid & 127, 32 channels
        {
            const int px = tid & (kPtPix - 1), c0 = (tid >> 7) * 32;
            uint8_t* d = IMG + (uint32_t)(px >> 2) * kPtLboA + (uint32_t)(px & 3) * 4;
#pragma unroll 8
            for (int c = c0; c < c0 + 32; ++c) {
                float hi, lo;
                tc::split_tf32(RAW[c * kPtPix + px], hi, lo);
                *reinterpret_cast<float*>(d + (uint32_t)c * 16) = hi;
                *reinterpret_cast<float*>(d + (uint32_t)(64 + c) * 16) = lo;
            }
        }
        __syncthreads();                                 // RAW consumed
        if (tile + gridDim.x < ntiles) stage_raw(tile + gridDim.x);     // overlaps this tile's work
        bool vq[2];
        float go[2][kProjMaxOut];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const long idx = base + tp + 64 * q;
            vq[q] = idx < total;
#pragma unroll
            for (int o = 0; o < kProjMaxOut; ++o) go[q][o] = (vq[q] && o < k.out_ch) ? __ldg(k.gout + idx * k.out_ch + o) : 0.f;
        }
        for (int ch = 0; ch < nchunks; ++ch) {
            const int ch0 = ch * kProjHC;
            const int nn = min(kProjHC, k.hid - ch0);
            // ---- activation phase: 8 hidden x 2 pixels per thread; all 16 pre-activations are requested before the first use
            float dp[8][2];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int nc = min(ch0 + 8 * tn + i, k.hid - 1);
                const float* src = k.pre_in + (size_t)nc * total + base + tp;
#pragma unroll
                for (int q = 0; q < 2; ++q) dp[i][q] = vq[q] ? __ldg(src + 64 * q) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = 8 * tn + i;
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
                for (int q = 0; q < 2; ++q) {
                    float a, gp;
                    gelu_both(dp[i][q], a, gp);
                    float s = 0.f;
#pragma unroll
                    for (int o = 0; o < kProjMaxOut; ++o) {
                        s = fmaf(go[q][o], w2v[o], s);
                        sw[o] = fmaf(go[q][o], a, sw[o]);
                    }
                    dp[i][q] = live ? s * gp : 0.f;
                    sb += dp[i][q];
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
            // D into its two operand images
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int px = tp + 64 * q;
                float hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) tc::split_tf32(dp[i][q], hi[i], lo[i]);
                // pixel-major (row = pixel, k = hidden 8 tn + i): hidden 4 m .. 4 m + 3 are 16 contiguous bytes
                const uint32_t oa = (uint32_t)(2 * tn) * kPtLboA + (uint32_t)px * 16;
                *reinterpret_cast<float4*>(A1hi + oa) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(A1hi + oa + kPtLboA) = make_float4(hi[4], hi[5], hi[6], hi[7]);
                *reinterpret_cast<float4*>(A1lo + oa) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<float4*>(A1lo + oa + kPtLboA) = make_float4(lo[4], lo[5], lo[6], lo[7]);
                // hidden-major (row = hidden, k = pixel)
                const uint32_t od = (uint32_t)(px >> 2) * kPtLboD + (uint32_t)(px & 3) * 4 + (uint32_t)(8 * tn) * 16;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    *reinterpret_cast<float*>(Dhi + od + i * 16) = hi[i];
                    *reinterpret_cast<float*>(Dlo + od + i * 16) = lo[i];
                }
            }
            tc::fence_proxy_async();                     // generic-proxy writes -> visible to the tensor core's reads
            tc::tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc::tc_fence_after();
                // DIN[pixel][channel]: accumulator columns 0-63
                const uint32_t wch = w_a + (uint32_t)(2 * ch) * kPtWBytes;
#pragma unroll
                for (int ks = 0; ks < kProjHC / 8; ++ks) {
                    const uint64_t da_hi = tc::make_smem_desc(a1hi_a + ks * 2 * kPtLboA, kPtLboA, 128);
                    const uint64_t da_lo = tc::make_smem_desc(a1lo_a + ks * 2 * kPtLboA, kPtLboA, 128);
                    const uint64_t db_hi = tc::make_smem_desc(wch + ks * 2 * kPtLboW, kPtLboW, 128);
                    const uint64_t db_lo = tc::make_smem_desc(wch + kPtWBytes + ks * 2 * kPtLboW, kPtLboW, 128);
                    tc::mma_tf32(tmem_base, da_hi, db_hi, idesc_din, (ch | ks) ? 1u : 0u);
                    tc::mma_tf32(tmem_base, da_hi, db_lo, idesc_din, 1u);
                    tc::mma_tf32(tmem_base, da_lo, db_hi, idesc_din, 1u);
                }
                // dW1[stacked channel][hidden of this chunk]: accumulator columns 64 + 32 ch .., never cleared after the first tile
                const uint32_t dw_tmem = tmem_base + 64u + (uint32_t)ch * kProjHC;
#pragma unroll 4
                for (int ks = 0; ks < kPtPix / 8; ++ks) {
                    const uint64_t da = tc::make_smem_desc(img_a + ks * 2 * kPtLboA, kPtLboA, 128);
                    const uint64_t db_hi = tc::make_smem_desc(dhi_a + ks * 2 * kPtLboD, kPtLboD, 128);
                    const uint64_t db_lo = tc::make_smem_desc(dlo_a + ks * 2 * kPtLboD, kPtLboD, 128);
                    tc::mma_tf32(dw_tmem, da, db_hi, idesc_dw, (it | ks) ? 1u : 0u);
                    tc::mma_tf32(dw_tmem, da, db_lo, idesc_dw, 1u);
                }
                tc::tc_commit(bar);
            }
            tc::mbar_wait(bar, phase);                   // the images may be rewritten, the accumulators read
            phase ^= 1u;
            tc::tc_fence_after();
        }
        // ---- input gradients: warps 0-3 own TMEM lanes 32 w .. 32 w + 31 = the tile's pixels
        if (warp < 4) {
            const long idx = base + 32 * warp + lane;
            long b = 0, rp = 0, pp = 0;
            const bool valid = idx < total;
            if (valid) raw_to_padded(g, idx, b, rp, pp);
            long cur_stride = -1, off = 0;
#pragma unroll
            for (int c0 = 0; c0 < CT; c0 += 16) {
                uint32_t r[16];
                tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0, r);
                tc::tmem_ld_wait();
                if (valid) {
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
            tc::tc_fence_before();
        }
        if (tn == 0) {
#pragma unroll
            for (int o = 0; o < kProjMaxOut; ++o)
                if (o < k.out_ch) {
                    const float v = warp_sum(go[0][o] + go[1][o]);
                    if (lane == 0) atomicAdd(accb2 + o, v);
                }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    tc::tc_fence_after();
    // ---- dW1: TMEM lanes 0-63 hold the hi-part rows (channel = lane), 64-127 the lo-part rows: both add into gw1
    if (warp < 4 && it > 0) {
        const int c = (32 * warp + lane) & 63;
        for (int ch = 0; ch < nchunks; ++ch)
            for (int c0 = 0; c0 < kProjHC; c0 += 16) {
                uint32_t r[16];
                tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(32 * warp) << 16) + 64u + (uint32_t)(ch * kProjHC + c0), r);
                tc::tmem_ld_wait();
                if (c < k.ctot) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = ch * kProjHC + c0 + j;
                        if (n < k.hid) atomicAdd(k.gw1 + n * k.ctot + c, __uint_as_float(r[j]));
                    }
                }
            }
    }
    for (int i = tid; i < k.hid; i += 256) atomicAdd(k.gb1 + i, accb1[i]);
    for (int i = tid; i < k.out_ch * k.hid; i += 256) atomicAdd(k.gw2 + i, accW2[i]);
    for (int i = tid; i < k.out_ch; i += 256) atomicAdd(k.gb2 + i, accb2[i]);
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, kPtTmemCols);
}

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
