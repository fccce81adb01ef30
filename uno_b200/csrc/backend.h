// Primitive-kernel interface between the orchestration (uno_api.cpp) and the device code
// (backend_cuda.cu: hand-written sm_100a kernels).  tests/hostemu/backend_host.cpp implements the
// same interface with plain loops so the orchestration (matrices, strides, corner maps, adjoints)
// can be checked on a machine without a GPU; that build is test infrastructure and is never loaded
// by the product.
#pragma once
#include <cstddef>
#include <cstdint>

#include "config.h"

namespace uno {

typedef void* stream_t;

// ---- persistent constants (plan matrices) ---------------------------------------------------------
int be_upload(void** dptr, const void* host, size_t bytes);   // allocate + copy, synchronous
void be_free(void* dptr);
int be_memset(void* d, int v, size_t bytes, stream_t s);
const char* be_name();
int be_current_device();   // ordinal of the device the calling thread has current (plan constants are cached per device)
const char* be_error_string(int code);   // message for a non-zero return of any be_* call

// ---- fork / join of independent kernel sequences (the two branches of an operator block) -----------------
// be_side_stream: a library-owned stream on the current device (created on first use).  be_fork makes `side` wait
// for everything enqueued on `main` so far; be_join makes `main` wait for everything enqueued on `side` so far.
// Both are event record + stream-wait pairs: asynchronous, capturable in a CUDA graph, and invisible to the caller,
// whose stream still orders the whole call.  Returning NULL from be_side_stream disables the overlap.
stream_t be_side_stream(int which = 0);   // which = 0, 1: two independent side streams per device
int be_fork(stream_t main_stream, stream_t side);
int be_join(stream_t main_stream, stream_t side);

// ---- optional per-launch timing (CUDA events on the launching stream), off by default --------------
// While enabled every be_* launch is bracketed by an event pair and tagged with its role and its
// algorithmic bytes / flops.  be_profile_report synchronises and writes one JSON object:
//   {"tag": {"launches": n, "ms": total, "bytes": total, "flops": total}, ...}
void be_profile_enable(int on);
size_t be_profile_report(char* buf, size_t cap);
// Scopes group the launches of one operator call (a fused spectral convolution at one U-level): between
// be_profile_scope_begin and be_profile_scope_end every launch of the calling thread is also charged to `label`, together
// with the call's algorithmic bytes / flops (tensors in and out of the operator, not its intermediates).
// be_profile_report_scopes writes {"label": {"calls": c, "launches": n, "ms": total, "bytes": total, "flops": total}, ...}.
// No-ops while profiling is off.
int be_profile_enabled();
void be_profile_scope_begin(const char* label, double bytes, double flops);
void be_profile_scope_end();
size_t be_profile_report_scopes(char* buf, size_t cap);
long be_launch_count();                  // kernels launched by this library since load
// NVTX ranges: no-ops unless the switch `nvtx` is on
void be_range_push(const char* label);
void be_range_pop();

// ---- C[M,N] = A[M,K] * B[K,N]  (fp32, row-major B and C, strided A), optionally batched -----------
enum GemmEpi {
    EPI_STORE = 0,        // C = acc (+ bias[m])
    EPI_ACCUM = 1,        // C = C + acc
    EPI_ACCUM_GELU = 2,   // C = C + acc ; C2 = gelu(C)
    EPI_ACCUM_GELU_INPLACE = 3,   // C = gelu(C + acc)
};
struct GemmArgs {
    const float* A = nullptr; long a_rs = 0, a_cs = 1;   // A(m,k) = A[m*a_rs + k*a_cs]
    const float* B = nullptr; long ldb = 0;              // B(k,n) = B[k*ldb + n]
    float* C = nullptr; long ldc = 0;                    // C(m,n) = C[m*ldc + n]
    float* C2 = nullptr;                                 // second output (same ldc / batch stride)
    const float* bias = nullptr;                         // per-row bias [M] (EPI_STORE only)
    int M = 0, N = 0, K = 0;
    int batch = 1; long sA = 0, sB = 0, sC = 0;          // batch strides in elements
    int epi = EPI_STORE;
    const char* tag = "gemm";                            // role of this launch (profiling only)
    bool b_const = false;                                // B is a persistent plan constant (its tensor-core image may be cached)
    bool channel_mix = false;                            // A is a small weight matrix shared by the batch, B/C are [channels, pixels] per sample
};
int be_gemm(const GemmArgs& a, stream_t s);

// ---- C[M,N] (+)= sum_batch A_b[M,K] * B_b[N,K]^T   (weight-gradient reduction; C pre-zeroed) -------
struct GemmNtArgs {
    const float* A = nullptr; long lda = 0, sA = 0;
    const float* B = nullptr; long ldb = 0, sB = 0;
    float* C = nullptr; long ldc = 0;
    int M = 0, N = 0, K = 0, batch = 1;
};
int be_gemm_nt_atomic(const GemmNtArgs& a, stream_t s);

// ---- complex transform along a middle axis: Y[o,j,i] = sum_h Mat[j,h] * X[o,h,i] --------------------
struct MidArgs {
    const float* X = nullptr;    // complex64 interleaved [O, H, I]
    const float* Mat = nullptr;  // complex64 interleaved [J, H]
    float* Y = nullptr;          // complex64 interleaved [O, J, I]
    int O = 0, H = 0, J = 0, I = 0;
};
int be_mid(const MidArgs& a, stream_t s);

// ---- per-mode complex contraction: C[m,n,q] = sum_k opA(A[m,k,q]) * opB(B[k,n,q]) -----------------
// q = (qo, qi): qi contiguous (length q_inner), qo strided.  All strides in complex elements.
// Up to four independent problems of identical shape (the spectrum corners weights1..4) run in one launch.
struct CmmArgs {
    int ncorner = 1;
    const float* A[4] = {nullptr, nullptr, nullptr, nullptr};
    const float* B[4] = {nullptr, nullptr, nullptr, nullptr};
    float* C[4] = {nullptr, nullptr, nullptr, nullptr};
    long a_sm = 0, a_sk = 0, a_sqo = 0; int conjA = 0;
    long b_sk = 0, b_sn = 0, b_sqo = 0; int conjB = 0;
    long c_sm = 0, c_sn = 0, c_sqo = 0;
    int M = 0, N = 0, K = 0, q_outer = 1, q_inner = 0;
    // 1: the result must not depend on the order of atomic additions (forward outputs: repeated calls stay bit-identical);
    // 0: the backend may split the reduction over several CTAs that add into C (gradients, like the other weight-gradient kernels)
    int deterministic = 0;
};
int be_cmm(const CmmArgs& a, stream_t s);

// ---- banded (resample) operators -------------------------------------------------------------------
// last axis : y[r, j]    = sum_t w[j,t] * x[r, start[j]+t]          x [R, n_in]     y [R, n_out]
// mid axis  : y[o, j, i] = sum_t w[j,t] * x[o, start[j]+t, i]       x [O, n_in, I]  y [O, n_out, I]
struct BandedArgs {
    const float* x = nullptr; float* y = nullptr;
    const int* start = nullptr; const float* w = nullptr;
    int n_in = 0, n_out = 0, taps = 0;
    long outer = 0;    // R or O
    int inner = 1;     // I (1 for the last-axis form)
};
int be_banded(const BandedArgs& a, stream_t s);

// fused separable 2-D form: y[p] = R0 * x[p] * R1^T for every plane p   x [P, n_in0, n_in1] -> y [P, n_out0, n_out1]
// span0 / span1: largest input extent (rows / cols) needed by any 32-row / 64-column output tile (generic kernel);
// tile_span0 / tile_span1: the same for the register-blocked kernel's tiles (tile_groups0 row groups x 64 columns,
// 128 columns when G1 == 8).
struct Banded2DArgs {
    const float* x = nullptr; float* y = nullptr; long planes = 0;
    const int* start0 = nullptr; const float* w0 = nullptr; int n_in0 = 0, n_out0 = 0, taps0 = 0, span0 = 0;
    const int* start1 = nullptr; const float* w1 = nullptr; int n_in1 = 0, n_out1 = 0, taps1 = 0, span1 = 0;
    float* tmp = nullptr;   // scratch of planes * max(n_in0,n_out0) * max(n_in1,n_out1) floats for the two-pass fallback
    // optional register-blocked images of the same bands (plan.h BandGroups); when both axes have one the
    // register-blocked kernel runs: G outputs share a window of W inputs starting at gs[g], weights D[g][W][G]
    const int* gs0 = nullptr; const float* D0 = nullptr; int G0 = 0, W0 = 0, ng0 = 0, tile_groups0 = 0, tile_span0 = 0;
    const int* gs1 = nullptr; const float* D1 = nullptr; int G1 = 0, W1 = 0, ng1 = 0, tile_span1 = 0;
};
int be_banded2d(const Banded2DArgs& a, stream_t s);

// ---- elementwise / reductions ----------------------------------------------------------------------
int be_gelu_fwd(const float* pre, float* y, size_t n, stream_t s);
// Upstream gradient of an operator whose output has C planes of L elements per sample: the SUM of up to two tensors, each
// [B, C, L] with its own batch stride in floats (C*L when contiguous, larger when the tensor is a channel slice of a wider one:
// the halves of a skip concatenation's gradient).  The kernel that first reads the upstream gradient adds the two on the fly,
// so neither the slice copy nor the sum of a twice-used tensor's gradients ever exists in HBM.  p1 == NULL: one source.
struct UpGrad {
    const float* p0 = nullptr; long bs0 = 0;
    const float* p1 = nullptr; long bs1 = 0;
};
// plane-structured activation backward with the per-channel sum of the result folded in:
//   g = (gy.p0 + gy.p1) * gelu'(pre)  (pre == NULL: g = gy.p0 + gy.p1) ;
//   gbias[p % C] += alpha * sum over the plane's L elements of g   (gbias pre-zeroed; NULL: not wanted)
int be_gelu_bwd_bias(const UpGrad& gy, const float* pre, float* g, long planes, int C, long L, float* gbias,
                     float alpha, stream_t s);
// per-plane mean / rstd over L contiguous elements: stats[p] = (mean, rstd)
int be_plane_stats(const float* x, float* stats, long planes, long L, float eps, stream_t s);
// statistics + normalisation (+ GELU) in one call: stats[p] = (mean, rstd) out, y as below.  The CUDA backend keeps each
// plane on chip between the two (thread-block clusters) when it fits, else runs be_plane_stats + be_norm_act_fwd.
int be_norm_fused_fwd(const float* x, float* stats, const float* gamma, const float* beta, float* y, long planes, int C,
                      long L, float eps, int non_lin, stream_t s);
// y = [gelu]( (x-mean)*rstd*gamma[c] + beta[c] ), plane p -> channel p % C
int be_norm_act_fwd(const float* x, const float* stats, const float* gamma, const float* beta,
                    float* y, long planes, int C, long L, int non_lin, stream_t s);
// backward of the above: g = d/dx ; ggamma[c] += ..., gbeta[c] += ... (pre-zeroed by the caller).
// Every plane of g sums to zero in exact arithmetic (the mean is projected out).
int be_norm_act_bwd(const UpGrad& gy, const float* x, const float* stats, const float* gamma,
                    const float* beta, float* g, float* ggamma, float* gbeta, long planes, int C,
                    long L, int non_lin, stream_t s);
// out[c] += alpha * sum over planes p with p % C == c and over the plane's L elements  (out pre-zeroed)
int be_channel_sum(const float* x, float* out, long planes, int C, long L, float alpha, stream_t s);
// y[p, :] += v[p % C] * alpha
int be_add_channel_const(float* y, const float* v, float alpha, long planes, int C, long L, stream_t s);

// ---- model glue around the blocks: per-pixel MLPs, fused -----------------
// Pixels live on up to three spatial axes; a "raw" grid n[] sits at offset lo[] inside a zero-padded grid N[]
// (2-D callers pass n[0] = N[0] = 1).  Channels-last tensors: a [B, n.., raw_ch], grid [n.., grid_ch],
// out [B, n.., out_ch]; channels-first tensors: h / src / gsrc [B, C, N..].
//
// lift:    h = pad( gelu( W_b * gelu( W_a * [a ; grid] + b_a ) + b_b ) )      W_a [hid, raw_ch+grid_ch], W_b [out_ch, hid]
// project: out = W_2 * gelu( W_1 * crop(cat(src_0, src_1, ..)) + b_1 ) + b_2  W_1 [hid, sum src_ch],     W_2 [out_ch, hid]
struct LiftArgs {
    int batch = 0, n[3] = {1, 1, 1}, N[3] = {1, 1, 1}, lo[3] = {0, 0, 0};
    int raw_ch = 0, grid_ch = 0, hid = 0, out_ch = 0;
    const float* a = nullptr; const float* grid = nullptr;
    const float* w_a = nullptr; const float* b_a = nullptr; const float* w_b = nullptr; const float* b_b = nullptr;
    float* h = nullptr;                 // fwd output
    const float* gh = nullptr;          // bwd input  [B, out_ch, N..]
    const float* gh2 = nullptr;         // optional second upstream gradient of the same shape, added on the fly (h used twice)
    float* ga = nullptr;                // bwd output [B, n.., raw_ch] (optional)
    float* gw_a = nullptr; float* gb_a = nullptr; float* gw_b = nullptr; float* gb_b = nullptr;   // bwd outputs, PRE-ZEROED (accumulated with atomics)
};
int be_lift_supported(const LiftArgs& a);   // 1 if the kernels take this shape
int be_lift_fwd(const LiftArgs& a, stream_t s);
int be_lift_bwd(const LiftArgs& a, stream_t s);

struct ProjArgs {
    int batch = 0, n[3] = {1, 1, 1}, N[3] = {1, 1, 1}, lo[3] = {0, 0, 0};
    int nsrc = 0, src_ch[4] = {0, 0, 0, 0}, hid = 0, out_ch = 0;
    const float* src[4] = {nullptr, nullptr, nullptr, nullptr};
    const float* w1 = nullptr; const float* b1 = nullptr; const float* w2 = nullptr; const float* b2 = nullptr;
    float* out = nullptr;               // fwd output [B, n.., out_ch]
    // hidden pre-activations W_1*x + b_1, [hid, B * n..] (hidden unit major): optional fwd output / optional bwd input.
    // With it the backward skips recomputing the first layer (a third of its arithmetic) for 4*hid bytes per pixel.
    float* pre_out = nullptr;
    const float* pre_in = nullptr;
    const float* gout = nullptr;        // bwd input
    float* gsrc[4] = {nullptr, nullptr, nullptr, nullptr};   // bwd outputs [B, src_ch, N..], fully written (zero outside the crop)
    float* gw1 = nullptr; float* gb1 = nullptr; float* gw2 = nullptr; float* gb2 = nullptr;       // bwd outputs, PRE-ZEROED
};
int be_proj_supported(const ProjArgs& a);
int be_proj_fwd(const ProjArgs& a, stream_t s);
int be_proj_bwd(const ProjArgs& a, stream_t s);

// ---- training-step ops next to the path ---------------------------------------------
// The reference's Adam (Adam.py:8-52) over a list of tensors in one launch per 24 tensors.  numel counts floats; a complex
// tensor is (re, im) pairs whose second moment is |g|^2 stored as (v, 0).  vmax is used only when amsgrad (real tensors).
struct AdamTensor {
    float* param = nullptr; const float* grad = nullptr; float* exp_avg = nullptr; float* exp_avg_sq = nullptr;
    float* max_exp_avg_sq = nullptr;
    long numel = 0;
    int is_complex = 0;
};
struct AdamHyper {
    double lr = 1e-3, beta1 = 0.9, beta2 = 0.999, eps = 1e-8, weight_decay = 0.0;
    int amsgrad = 0;
    int step = 1;          // 1-based step count of every tensor in the call
};
int be_adam_step(const AdamTensor* t, int n, const AdamHyper& h, stream_t s);

// relative L2 loss of utilities3.LpLoss (p = 2): x, y [B, N]; reduction 0 = none ([B] out), 1 = sum, 2 = mean.
// norms [B, 2] = (||x-y||, ||y||) out of fwd / in to bwd; acc = scratch of 2*B doubles.
int be_lp_loss_fwd(const float* x, const float* y, int B, long N, int reduction, float* loss, float* norms, double* acc, stream_t s);
int be_lp_loss_bwd(const float* x, const float* y, const float* norms, const float* gl, int B, long N, int reduction, float* gx,
                   stream_t s);

}  // namespace uno
