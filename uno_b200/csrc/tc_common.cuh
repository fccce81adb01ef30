// sm_100a primitives used by the tensor-core kernels: mbarrier, bulk async copy (UBLKCP), proxy
// fences, tcgen05 (TMEM alloc / mma / commit / ld) and the shared-memory matrix descriptor.
//
// Numerics: fp32 operands are fed to the tf32 tensor-core path as an error-compensated split
//   x = hi + lo,  hi = x with the 13 low mantissa bits cleared (exactly tf32), lo = x - hi (exact in fp32)
//   A*B ~= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi     (fp32 accumulation in TMEM)
// which keeps ~2^-21 relative accuracy per product ("3xTF32"), i.e. fp32-class results.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace uno {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// for roles whose wait is long (a whole tile): back off so the spinning warp does not steal issue slots
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(128);
}

// ---- async proxy ---------------------------------------------------------------------------------------
// make generic-proxy st.shared visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (src, dst, size 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- tcgen05 -------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// arrive on an mbarrier when all tcgen05 operations previously issued by this thread have completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, one thread issues
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Instruction descriptor (cute::UMMA::InstrDescriptor bit layout): tf32 x tf32 -> f32, dense.
//   [4,6) c_format = 1 (F32)   [7,10) a_format = 2 (TF32)   [10,13) b_format = 2
//   [15] a_major (0 = K-major, 1 = MN-major)   [16] b_major   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layouts, 16-byte units:
//   K-major : 8 rows x 16 B core matrices (128 B contiguous); SBO = byte stride between 8-row groups,
//             LBO = byte stride between core matrices adjacent along K (next 4 fp32 of K)
//   MN-major: 8 k x 16 B (4 consecutive MN elements) core matrices; SBO = stride to the next 4 MN
//             elements, LBO = stride to the next 8 k
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 0) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version for sm_100
    d |= (uint64_t)(layout_type & 7) << 61;   // 0 = no swizzle, 1 = 128B swizzle with 32-byte base (MN-major 32-bit operands)
    return d;
}

// The issuing thread is the critical path of the pipelined kernels (measured: with loads, stores and MMAs of the
// analysis kernel switched off, the descriptor arithmetic and per-instruction election of the issuer alone kept
// 40-55 % of its run time), so descriptors are formed once and then only ADVANCED: the start address is the low
// 14 bits of the low word in 16-byte units, a byte offset adds (bytes >> 4) there (no carry: shared memory < 256 KB).
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }
// one lane of a converged warp (elect.sync); tcgen05.mma / commit issued under it need no per-instruction election loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}

// TMEM -> registers.  16x256b: a warp reads 16 lanes x 8 columns per repeat; thread t holds
//   (row t/4, cols 2(t%4), 2(t%4)+1) in regs 4r+0,4r+1 and (row t/4 + 8, same cols) in regs 4r+2,4r+3.
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
// 32x32b: thread t of the warp reads lane (base + t), 16 consecutive columns
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16-byte read-only load; HINT asks L2 to fetch the surrounding 256 bytes (the loaders of the streaming kernels walk every row
// 128 bytes at a time, so the neighbouring line is always wanted next: one DRAM access of 256 contiguous bytes instead of two)
template <bool HINT>
__device__ __forceinline__ float4 ldg_f4(const float4* p) {
    if constexpr (HINT) {
        float4 v;
        asm volatile("ld.global.nc.L2::256B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
        return v;
    } else {
        return __ldg(p);
    }
}

// Exact-erf GELU, forward only, from ONE special-function operation and no division:  with z = min(|x|, 6),
//   erfc(z / sqrt2) / 2 = 2^(z * r(z) - 1),   r = degree-6 polynomial fitted to log2(erfc(z / sqrt2)) / z on [0, 6]
//   gelu(x) = x * Phi(x) = max(x, 0) - |x| * erfc(|x| / sqrt2) / 2          (no cancellation in the negative tail)
// Measured in fp32 against the fp64 erf form over [-12, 12]: |error| <= 2.8e-7 (the rounding of the result itself), two orders
// below the parity tolerance.  ~12 instructions (6 FMA, ex2.approx) against ~35 for libm erff, whose two polynomial branches
// both execute in a diverged warp.  Used by the FMA-bound SIMT kernels (lift / projection forward: 1.14 -> 1.00 ms per Darcy
// step); NOT by the synthesis epilogue, which is latency-bound and measured slower with it (tc_rowgemm.cuh).
__device__ __forceinline__ float gelu_fwd_fast(float x) {
    const float a = fabsf(x);
    const float z = fminf(a, 6.0f);
    float r = fmaf(z, 4.659973911e-06f, -1.725336369e-05f);
    r = fmaf(z, r, -5.553429364e-04f);
    r = fmaf(z, r, 7.624045014e-03f);
    r = fmaf(z, r, -5.289301276e-02f);
    r = fmaf(z, r, -4.590760469e-01f);
    r = fmaf(z, r, -1.151119947e+00f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(z, r, -1.0f)));
    return fmaf(-a, e, fmaxf(x, 0.f));
}

// fp32 -> (hi, lo) split for the 3xTF32 scheme
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

}  // namespace tc
}  // namespace uno
