// Host-side planning for the U-NO spectral-convolution path: every constant matrix the kernels
// multiply by (truncated DFT twiddles with the reference's mode/corner maps baked in, the
// anti-aliased bicubic resample bands, the 3-D "spectral resample" operator of pointwise_op_3D).
// Pure C++ (no CUDA): built in double precision, rounded once to fp32.  Exposed for testing through
// the host-only uno_plan_* entry points in include/uno_b200.h.
//
// Reference semantics restated here (paths relative to /root/reference):
//   integral_operators.py:47-72, :181-207, :385-427   SpectralConv{1,2,3}d_Uno.forward
//   integral_operators.py:224-243                      pointwise_op_2D.forward (bicubic, antialias)
//   integral_operators.py:438-468                      pointwise_op_3D.forward (rfftn / corner copy / irfftn)
#pragma once
#include <vector>

namespace uno {

// ---- truncated DFT along the LAST (half-spectrum) axis -------------------------------------------
// analysis matrix  [n x 2m]  : col 2k = scale*cos(2 pi k j / n), col 2k+1 = -scale*sin(...)
//   real row of n samples  ->  m complex bins (interleaved re,im) of  scale * sum_j x_j e^{-2 pi i k j / n}
std::vector<float> dft_last_analysis(int n, int m, double scale);
// synthesis matrix [2m x n]  : row 2k = c_k*scale*cos(2 pi k j / n), row 2k+1 = -c_k*scale*sin(...)
//   m complex bins -> n real samples of an unnormalised C2R: c_k = 1 for DC / Nyquist, else 2; the
//   imaginary part of DC / Nyquist is dropped exactly as irfft does.  `hermitian` = false gives c_k = 1
//   for all k (adjoint of the analysis matrix, used by the backward pass).
std::vector<float> dft_last_synthesis(int n, int m, double scale, bool hermitian);

// ---- truncated DFT along a LEADING (full-spectrum) axis, complex -> complex -----------------------
// kept index kappa in [0,2m): kappa<m -> frequency kappa ; kappa>=m -> frequency n-m+(kappa-m)
// analysis  [2m x n] complex interleaved: e^{-2 pi i k(kappa) h / n}
std::vector<float> dft_mid_analysis(int n, int m);
// synthesis [n x 2m] complex interleaved: e^{+2 pi i k'(kappa) j / n}; when 2m > n the low block's
// overlapped entries (kappa >= n-m) are overwritten by the high block in the reference (write order
// weights1, weights2, ...) so their columns are ZERO here.
std::vector<float> dft_mid_synthesis(int n, int m);

// real [r x c] -> [c x r]
std::vector<float> transpose_real(const std::vector<float>& a, int r, int c);
// complex interleaved [r x c] -> conjugate transpose [c x r]
std::vector<float> conj_transpose(const std::vector<float>& a, int r, int c);

// ---- pointwise_op_3D spectral resample, one axis ---------------------------------------------------
// leading axes: dense complex [n_out x n_in]:  L[j,h] = sum_{k in keep, k < n_out} e^{2 pi i k j/n_out} e^{-2 pi i k h/n_in}
// keep = [0, min(n_out/2, n_in)) U [max(n_in - n_out/2, 0), n_in)
std::vector<float> sr_mid(int n_in, int n_out);
// last axis: number of kept half-spectrum bins
int sr_last_modes(int n_in, int n_out);
// Opt-in "fixed" variant (NOT the reference's behaviour): the band-limited Fourier resample the reference's
// operator approximates -- the symmetric band |k| <= K = (min(n_in, n_out) - 1) / 2 is kept on every axis, negative frequencies
// map to negative frequencies, nothing is aliased and a constant stays the same constant (gain 1 instead of N_in / N_out).
//   leading axes: L[j,h] = sum_{k=-K..K} e^{2 pi i k j/n_out} e^{-2 pi i k h/n_in}   (real: a Dirichlet kernel; stored complex)
std::vector<float> sr_mid_fixed(int n_in, int n_out);
// last axis: bins 0..K of the half spectrum
int sr_last_modes_fixed(int n_in, int n_out);

// ---- anti-aliased bicubic (align_corners=True) as a banded matrix ---------------------------------
struct Banded {
    int n_in = 0, n_out = 0, taps = 0;
    std::vector<int> start;    // [n_out] first input index of the band
    std::vector<float> w;      // [n_out x taps], zero padded
};
Banded bicubic_aa(int n_in, int n_out);
Banded banded_transpose(const Banded& b);   // [n_in x n_out] operator, also banded
std::vector<float> banded_dense(const Banded& b);

// Register-blocked image of a banded operator for the fused 2-D resample kernel: G consecutive output
// rows share one input window of W samples starting at gstart[g] (always inside [0, n_in - W]), and
// D[g][u][q] is the dense weight of window sample u for output 4g+q (zero outside the band, zero for
// outputs past n_out).  One of (G,W) = (8,8), (4,8), (4,16) is chosen -- the first whose windows fit;
// ok = false when none does (very large scale factors, or inputs shorter than the window).
struct BandGroups {
    bool ok = false;
    int G = 0, W = 0, ng = 0, n_in = 0, n_out = 0;
    std::vector<int> gstart;   // [ng]
    std::vector<float> D;      // [ng][W][G]
    int span(int groups_per_tile) const;   // largest input extent of a tile of that many groups
};
BandGroups band_groups(const Banded& b);
std::vector<float> band_groups_dense(const BandGroups& g);   // [n_out x n_in], for testing

}  // namespace uno
