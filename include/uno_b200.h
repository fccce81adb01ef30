/*
 * uno_b200 -- C ABI of the B200-native U-NO integral-operator path.
 *
 * The reference (ashiq24/UNO) is pure Python on PyTorch and has no FFI of its own: its boundary for
 * this path is the nn.Module API of integral_operators.py.  Each entry point below replaces the body
 * of one reference forward (or its autograd-generated backward); uno_b200/integral_operators.py keeps
 * the reference's module classes / signatures and calls these through ctypes.
 *
 * Conventions
 *   - every tensor argument is a DEVICE pointer to a dense, contiguous fp32 buffer in the reference's
 *     layout: activations [B, C, d1(, d2(, d3))]; spectral weights complex64 interleaved (re,im)
 *     [Ci, Co, m1(, m2(, m3))] -- exactly the memory of torch.view_as_real(weightsN);
 *     Conv{n}d(k=1) weight [Co, Ci], bias [Co]; InstanceNorm gamma/beta [Co].
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises.
 *   - `ws` is caller-owned scratch of at least *_workspace_bytes(desc); contents are undefined after.
 *   - return value 0 = ok; non-zero = error, message in uno_last_error() (thread-local).
 *       1 invalid argument / unsupported shape (the reference raises RuntimeError for these)
 *       2 workspace too small      3 CUDA runtime error
 *   - plan constants (twiddle matrices, resample bands) are built on first use of a shape and cached
 *     for the life of the process; the first call of a new shape therefore allocates device memory
 *     and must not happen inside CUDA-graph capture.
 */
#ifndef UNO_B200_H
#define UNO_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNO_OK 0
#define UNO_EINVAL 1
#define UNO_EWORKSPACE 2
#define UNO_ECUDA 3

/* Shape of one SpectralConv{1,2,3}d_Uno / pointwise_op_{2,3}D / OperatorBlock call.
 * Unused trailing dims must be 1 (in_dim/out_dim) or 0 (modes).                          */
typedef struct uno_conv_desc {
    int ndim;        /* 1, 2 or 3 */
    int batch;       /* B */
    int in_ch;       /* in_codim  */
    int out_ch;      /* out_codim */
    int in_dim[3];   /* input grid  */
    int out_dim[3];  /* output grid (dim1, dim2, dim3 of the forward call) */
    int modes[3];    /* modes1, modes2, modes3 (ignored by the pointwise entry points) */
} uno_conv_desc;

typedef struct uno_block_desc {
    uno_conv_desc conv;
    int normalize;   /* OperatorBlock(Normalize=...) : InstanceNorm(affine) on conv(x)+w(x) */
    int non_lin;     /* OperatorBlock(Non_Lin=...)   : exact-erf GELU                       */
    float eps;       /* InstanceNorm eps (1e-5)                                             */
} uno_block_desc;

const char* uno_last_error(void);
int uno_version(void);
const char* uno_backend_name(void);   /* "cuda-sm100a" for the product library */
void uno_clear_plans(void);

/* ---- SpectralConv{1,2,3}d_Uno -------------------------------------------------------------------
 * fwd replaces integral_operators.py:47-72 (1-D), :181-207 (2-D), :385-427 (3-D).
 *   x  [B,Ci,*in]   w[nW] (nW = 1,2,4 = weights1..)   y [B,Co,*out]
 *   xhat: optional out, the kept-mode input spectrum (complex64, uno_spectral_conv_xhat_elems()
 *         complex elements) that bwd needs; pass NULL for inference.
 * bwd is the autograd backward of the same lines:
 *   gy [B,Co,*out], xhat from fwd -> gx [B,Ci,*in] (overwritten, or += when accumulate_gx),
 *   gw[nW] like w (overwritten; torch convention dL/dRe + i dL/dIm).  gx or gw may be NULL to skip. */
int uno_spectral_conv_check(const uno_conv_desc* d);
size_t uno_spectral_conv_workspace_bytes(const uno_conv_desc* d);
size_t uno_spectral_conv_xhat_elems(const uno_conv_desc* d);
int uno_spectral_conv_fwd(const uno_conv_desc* d, const float* x, const float* const* w, float* y,
                          float* xhat, void* ws, size_t ws_bytes, void* stream);
int uno_spectral_conv_bwd(const uno_conv_desc* d, const float* gy, const float* xhat,
                          const float* const* w, float* gx, float* const* gw, int accumulate_gx,
                          void* ws, size_t ws_bytes, void* stream);

/* ---- pointwise_op_2D / pointwise_op_3D ----------------------------------------------------------
 * fwd replaces integral_operators.py:224-243 (2-D: Conv2d(k=1) + bicubic anti-aliased resample,
 * align_corners=True) and :438-468 (3-D: Conv3d(k=1) + rfftn / corner copy / irfftn(s=out)).
 *   saved: optional out of uno_pointwise_saved_elems() floats (the resampled input when the resample
 *          runs before the channel mix); bwd recomputes it when NULL.
 * bwd: gz [B,Co,*out] -> gx [B,Ci,*in] (overwritten), gconv_w [Co,Ci], gconv_b [Co] (overwritten). */
size_t uno_pointwise_workspace_bytes(const uno_conv_desc* d);
size_t uno_pointwise_saved_elems(const uno_conv_desc* d);
int uno_pointwise_fwd(const uno_conv_desc* d, const float* x, const float* conv_w,
                      const float* conv_b, float* z, float* saved, void* ws, size_t ws_bytes,
                      void* stream);
int uno_pointwise_bwd(const uno_conv_desc* d, const float* gz, const float* x, const float* saved,
                      const float* conv_w, float* gx, float* gconv_w, float* gconv_b, void* ws,
                      size_t ws_bytes, void* stream);

/* ---- OperatorBlock_2D / OperatorBlock_3D ---------------------------------------------------------
 * fwd replaces integral_operators.py:272-284 / :501-513:  y = gelu?( IN?( conv(x) + w(x) ) ).
 *   pre  : optional out [B,Co,*out]: the tensor bwd needs -- conv(x)+w(x) (pre-norm sum when
 *          normalize, else the pre-activation).  NULL for inference.
 *   stats: [B*Co, 2] (mean, rstd) out, required when normalize and pre != NULL.
 * bwd: gy -> gx, gw[nW], gconv_w, gconv_b, ggamma, gbeta (all overwritten).                          */
size_t uno_operator_block_workspace_bytes(const uno_block_desc* d);
int uno_operator_block_fwd(const uno_block_desc* d, const float* x, const float* const* w,
                           const float* conv_w, const float* conv_b, const float* gamma,
                           const float* beta, float* y, float* xhat, float* pw_saved, float* pre,
                           float* stats, void* ws, size_t ws_bytes, void* stream);
int uno_operator_block_bwd(const uno_block_desc* d, const float* gy, const float* x,
                           const float* xhat, const float* pw_saved, const float* pre,
                           const float* stats, const float* const* w, const float* conv_w,
                           const float* gamma, const float* beta, float* gx, float* const* gw,
                           float* gconv_w, float* gconv_b, float* ggamma, float* gbeta, void* ws,
                           size_t ws_bytes, void* stream);
/* Same, for an upstream gradient that is NOT one contiguous tensor: the sum of gy and (optional) gy2, each [B, out_ch, out_dim..]
 * with its own batch stride in floats (0 = contiguous).  This is what autograd hands a block whose output was used twice
 * (two gradients to add) and / or concatenated with a skip tensor along the channels (a channel slice of the concatenation's
 * gradient): the first kernel of the backward reads the sources in place, so neither the sum nor the slice copy
 * (the reference's AddBackward / `.contiguous()`) touches HBM. */
int uno_operator_block_bwd2(const uno_block_desc* d, const float* gy, long gy_batch_stride, const float* gy2,
                            long gy2_batch_stride, const float* x, const float* xhat, const float* pw_saved,
                            const float* pre, const float* stats, const float* const* w, const float* conv_w,
                            const float* gamma, const float* beta, float* gx, float* const* gw, float* gconv_w,
                            float* gconv_b, float* ggamma, float* gbeta, void* ws, size_t ws_bytes, void* stream);

/* ---- model glue around the blocks ------------------------------------
 * The reference models wrap the operator blocks in two per-pixel MLPs whose permute / pad / cat / crop
 * traffic dominates once the blocks are fused.  These entry points run each MLP as ONE kernel.
 *
 * lift   replaces  cat(a, grid) -> fc_n1/fc -> gelu -> fc0 -> gelu -> permute(0,3,1,2) -> F.pad
 *        darcy_flow_uno2d.py:96-107, navier_stokes_uno2d.py:191-201, navier_stokes_uno3d.py:497-511
 *   a    [B, *dim, raw_ch] channels-last      grid [*dim, grid_ch] (the positional features, built once)
 *   w_a  [hidden, raw_ch+grid_ch], b_a [hidden], w_b [out_ch, hidden], b_b [out_ch]   (nn.Linear layout)
 *   h    [B, out_ch, *(dim+pad_lo+pad_hi)] channels-first, zero in the padding
 * project replaces torch.cat(src..., dim=1) -> crop -> permute(0,2,3,1) -> fc1 -> gelu -> fc2
 *        darcy_flow_uno2d.py:121-131, navier_stokes_uno2d.py:215-225, navier_stokes_uno3d.py:551-575
 *   src[s] [B, src_ch[s], *(dim+pad_lo+pad_hi)]   w1 [hidden, sum src_ch], w2 [out_ch, hidden]
 *   out  [B, *dim, out_ch]
 * The backward entry points recompute the hidden activations (forward saves nothing) and overwrite every
 * gradient output; ga / gsrc[s] may be NULL to skip.  gsrc[s] is written over the whole padded grid
 * (zero outside the crop).  Limits: raw_ch+grid_ch <= 16, lift hidden <= 32, lift out_ch <= 64,
 * sum src_ch <= 64, project hidden <= 128, project out_ch <= 4 (uno_*_check reports UNO_EINVAL beyond). */
typedef struct uno_pixel_desc {
    int ndim;        /* 2 or 3 spatial axes */
    int batch;
    int dim[3];      /* un-padded / cropped grid (unused trailing entries 1) */
    int pad_lo[3];   /* padding before / after each axis (unused trailing entries 0) */
    int pad_hi[3];
} uno_pixel_desc;

typedef struct uno_lift_desc {
    uno_pixel_desc px;
    int raw_ch, grid_ch, hidden, out_ch;
} uno_lift_desc;

typedef struct uno_project_desc {
    uno_pixel_desc px;
    int nsrc;        /* 1..4 channel-first sources, concatenated along channels */
    int src_ch[4];
    int hidden, out_ch;
} uno_project_desc;

int uno_lift_check(const uno_lift_desc* d);
int uno_lift_fwd(const uno_lift_desc* d, const float* a, const float* grid, const float* w_a,
                 const float* b_a, const float* w_b, const float* b_b, float* h, void* stream);
int uno_lift_bwd(const uno_lift_desc* d, const float* gh, const float* a, const float* grid,
                 const float* w_a, const float* b_a, const float* w_b, const float* b_b, float* ga,
                 float* gw_a, float* gb_a, float* gw_b, float* gb_b, void* stream);
/* lift backward with a second upstream gradient gh2 (same shape, contiguous, may be NULL) added on the fly: the lifted input
 * feeds both the first block and the projection (darcy_flow_uno2d.py:121). */
int uno_lift_bwd2(const uno_lift_desc* d, const float* gh, const float* gh2, const float* a, const float* grid,
                  const float* w_a, const float* b_a, const float* w_b, const float* b_b, float* ga, float* gw_a, float* gb_a,
                  float* gw_b, float* gb_b, void* stream);
int uno_project_check(const uno_project_desc* d);
/* hidden_pre: optional [hidden, B * prod(dim)] floats -- the fc1 pre-activations.  When fwd writes them and bwd gets
 * them back, the backward skips recomputing fc1 (a third of its arithmetic); NULL on either side = recompute. */
int uno_project_fwd(const uno_project_desc* d, const float* const* src, const float* w1,
                    const float* b1, const float* w2, const float* b2, float* out, float* hidden_pre,
                    void* stream);
int uno_project_bwd(const uno_project_desc* d, const float* gout, const float* const* src,
                    const float* hidden_pre, const float* w1, const float* b1, const float* w2,
                    float* const* gsrc, float* gw1, float* gb1, float* gw2, float* gb2, void* stream);

/* ---- training-step ops next to the path ------------------------------------------------
 * uno_adam_step replaces the per-tensor Python loop of Adam.py:23-52 (functional `adam`) with one kernel launch per 24
 * tensors.  Semantics are the reference's, not torch.optim.Adam's: for a complex tensor the second moment is the running
 * average of g*conj(g) = |g|^2, stored -- like upstream -- in a complex tensor as (v, 0), so the real and imaginary part
 * of a weight share one denominator.  `numel` counts FLOATS (2 per complex element).  amsgrad on a complex tensor is
 * rejected (upstream raises from torch.maximum).  All tensors of a call share the 1-based `step`.
 * uno_lp_loss_* is utilities3.LpLoss.rel (utilities3.py:86-100) for p = 2 on x, y [B, N]:
 *   reduction 0: loss[B] = ||x_b - y_b|| / ||y_b||;  1: their sum (size_average=False);  2: their mean.
 *   norms [B, 2] (||x-y||, ||y||) is written by fwd and read by bwd; ws = 16*B bytes of scratch.                     */
typedef struct uno_adam_tensor {
    float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
    float* max_exp_avg_sq;   /* NULL unless amsgrad */
    long numel;
    int is_complex;
} uno_adam_tensor;

typedef struct uno_adam_hyper {
    double lr, beta1, beta2, eps, weight_decay;   /* doubles: 1 - beta2 must be formed before rounding to fp32, as upstream does */
    int amsgrad;
    int step;
} uno_adam_hyper;

int uno_adam_step(const uno_adam_tensor* tensors, int n, const uno_adam_hyper* h, void* stream);
int uno_lp_loss_fwd(const float* x, const float* y, int batch, long n, int reduction, float* loss, float* norms,
                    void* ws, size_t ws_bytes, void* stream);
int uno_lp_loss_bwd(const float* x, const float* y, const float* norms, const float* gloss, int batch, long n,
                    int reduction, float* gx, void* stream);

/* ---- run-time switches -------------------------------------------------------------------------------
 * Kernel-selection switches (uno_b200/csrc/config.h: "tc", "mid_tc", "cmm_tc", "kpipe_align", "kpipe_lw16",
 * "rowgemm_epi16", "rowgemm_parity", "norm_big_cluster", "overlap", "pointwise3d_fixed", "proj_simt", "nvtx", ...).
 * Their initial values come from the environment ONCE, at first use (UNO_B200_<NAME>); afterwards only these calls change
 * them -- no call path reads the environment.  Every setting computes the same function (parity-tested on B200) except
 * "pointwise3d_fixed", which is documented as deviating from the reference.  uno_config_name(i) enumerates the names
 * (NULL past the end).  Unknown name -> UNO_EINVAL.                                                             */
int uno_config_set(const char* name, int value);
int uno_config_get(const char* name, int* value);
const char* uno_config_name(int index);

/* ---- measurement hooks (bench.py) -----------------------------------------------------------------
 * uno_launch_count: kernels this library has launched since it was loaded.
 * uno_profile_enable(1) brackets every kernel launch with a CUDA-event pair on the launching stream;
 * uno_profile_report synchronises the device and writes a JSON object
 *   {"<kernel role>": {"launches": n, "ms": total, "bytes": algorithmic, "flops": algorithmic}, ...}
 * into buf (truncated to cap) and returns the untruncated length.  Off by default.
 * uno_profile_report_levels does the same per fused spectral-convolution call shape (one entry per U-level and
 * direction, "spectral fwd|bwd B=.. Ci->Co [in]->[out] modes=[..]"): its launches' summed time against the call's
 * ALGORITHMIC bytes (x, y, the spectral weights, + the block epilogue's tensors) and contraction flops
 *   {"<label>": {"calls": c, "launches": n, "ms": total, "bytes": total, "flops": total}, ...}                */
long uno_launch_count(void);
void uno_profile_enable(int on);
size_t uno_profile_report(char* buf, size_t cap);
size_t uno_profile_report_levels(char* buf, size_t cap);

/* ---- host-only planning helpers (no GPU touched; exercised by the CPU test-suite) ----------------
 * Each writes the dense fp32 matrix the kernels multiply by.  Sizes: see uno_b200/csrc/plan.h.      */
int uno_plan_dft_last_analysis(int n, int m, double scale, float* out /* [n, 2m] */);
int uno_plan_dft_last_synthesis(int n, int m, double scale, int hermitian, float* out /* [2m, n] */);
int uno_plan_dft_mid_analysis(int n, int m, float* out /* [2m, n] complex */);
int uno_plan_dft_mid_synthesis(int n, int m, float* out /* [n, 2m] complex */);
int uno_plan_sr_mid(int n_in, int n_out, float* out /* [n_out, n_in] complex */);
int uno_plan_sr_last_modes(int n_in, int n_out);
/* opt-in band-limited variant of the 3-D pointwise resample (UNO_B200_POINTWISE3D_FIXED=1; not the reference's behaviour) */
int uno_plan_sr_mid_fixed(int n_in, int n_out, float* out /* [n_out, n_in] complex */);
int uno_plan_sr_last_modes_fixed(int n_in, int n_out);
int uno_plan_bicubic_aa(int n_in, int n_out, int transpose, float* out /* dense [n_out,n_in] or its transpose */);
/* the register-blocked image of the same band that the fused 2-D resample kernel consumes, expanded to dense;
 * gw[0], gw[1] receive the chosen (outputs per group, window) or (0,0) when the band does not fit (generic kernel) */
int uno_plan_band_groups(int n_in, int n_out, int transpose, float* out, int* gw);

#ifdef __cplusplus
}
#endif
#endif /* UNO_B200_H */
