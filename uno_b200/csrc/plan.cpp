// See plan.h.  All trigonometry is done in long double on an exactly reduced angle index
// (k*j mod n as integers), then rounded once to fp32.
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstdint>

namespace uno {

namespace {
const long double kTwoPi = 6.283185307179586476925286766559005768L;

inline void cis(int64_t k, int64_t j, int64_t n, long double& c, long double& s) {
    int64_t r = ((k % n) * (j % n)) % n;
    if (r < 0) r += n;
    // exact values on the axes keep DC / Nyquist rows free of rounding dust
    if (r == 0) { c = 1.0L; s = 0.0L; return; }
    if (2 * r == n) { c = -1.0L; s = 0.0L; return; }
    if (4 * r == n) { c = 0.0L; s = 1.0L; return; }
    if (4 * r == 3 * n) { c = 0.0L; s = -1.0L; return; }
    long double a = kTwoPi * (long double)r / (long double)n;
    c = cosl(a);
    s = sinl(a);
}

inline int mid_freq(int kappa, int n, int m) { return kappa < m ? kappa : n - m + (kappa - m); }
}  // namespace

std::vector<float> dft_last_analysis(int n, int m, double scale) {
    std::vector<float> a((size_t)n * 2 * m);
    for (int j = 0; j < n; ++j)
        for (int k = 0; k < m; ++k) {
            long double c, s;
            cis(k, j, n, c, s);
            a[(size_t)j * 2 * m + 2 * k] = (float)(scale * c);
            a[(size_t)j * 2 * m + 2 * k + 1] = (float)(-scale * s);
        }
    return a;
}

std::vector<float> dft_last_synthesis(int n, int m, double scale, bool hermitian) {
    std::vector<float> a((size_t)2 * m * n);
    for (int k = 0; k < m; ++k) {
        double ck = 1.0;
        if (hermitian) ck = (k == 0 || (n % 2 == 0 && 2 * k == n)) ? 1.0 : 2.0;
        for (int j = 0; j < n; ++j) {
            long double c, s;
            cis(k, j, n, c, s);
            a[(size_t)(2 * k) * n + j] = (float)(ck * scale * c);
            a[(size_t)(2 * k + 1) * n + j] = (float)(-ck * scale * s);
        }
    }
    return a;
}

std::vector<float> dft_mid_analysis(int n, int m) {
    const int J = 2 * m;
    std::vector<float> a((size_t)J * n * 2);
    for (int kap = 0; kap < J; ++kap) {
        const int k = mid_freq(kap, n, m);
        for (int h = 0; h < n; ++h) {
            long double c, s;
            cis(k, h, n, c, s);
            a[((size_t)kap * n + h) * 2] = (float)c;
            a[((size_t)kap * n + h) * 2 + 1] = (float)(-s);
        }
    }
    return a;
}

std::vector<float> dft_mid_synthesis(int n, int m) {
    const int J = 2 * m;
    std::vector<float> a((size_t)n * J * 2, 0.0f);
    for (int kap = 0; kap < J; ++kap) {
        // low-block entries landing on rows also written by the high block are overwritten
        if (kap < m && kap >= n - m) continue;
        const int k = mid_freq(kap, n, m);
        for (int j = 0; j < n; ++j) {
            long double c, s;
            cis(k, j, n, c, s);
            a[((size_t)j * J + kap) * 2] = (float)c;
            a[((size_t)j * J + kap) * 2 + 1] = (float)s;
        }
    }
    return a;
}

std::vector<float> transpose_real(const std::vector<float>& a, int r, int c) {
    std::vector<float> t((size_t)r * c);
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < c; ++j) t[(size_t)j * r + i] = a[(size_t)i * c + j];
    return t;
}

std::vector<float> conj_transpose(const std::vector<float>& a, int r, int c) {
    std::vector<float> t((size_t)r * c * 2);
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < c; ++j) {
            t[((size_t)j * r + i) * 2] = a[((size_t)i * c + j) * 2];
            t[((size_t)j * r + i) * 2 + 1] = -a[((size_t)i * c + j) * 2 + 1];
        }
    return t;
}

std::vector<float> sr_mid(int n_in, int n_out) {
    const int h = n_out / 2;
    std::vector<char> keep(n_in, 0);
    // python slices [:h] and [-h:] on an axis of length n_in ([-0:] is the whole axis)
    for (int k = 0; k < std::min(h, n_in); ++k) keep[k] = 1;
    for (int k = (h == 0 ? 0 : std::max(n_in - h, 0)); k < n_in; ++k) keep[k] = 1;
    std::vector<float> a((size_t)n_out * n_in * 2, 0.0f);
    for (int j = 0; j < n_out; ++j)
        for (int hh = 0; hh < n_in; ++hh) {
            long double re = 0, im = 0;
            for (int k = 0; k < n_in && k < n_out; ++k) {
                if (!keep[k]) continue;
                long double c1, s1, c2, s2;
                cis(k, j, n_out, c1, s1);   // e^{+i a1}
                cis(k, hh, n_in, c2, s2);   // e^{-i a2}
                re += c1 * c2 + s1 * s2;
                im += s1 * c2 - c1 * s2;
            }
            a[((size_t)j * n_in + hh) * 2] = (float)re;
            a[((size_t)j * n_in + hh) * 2 + 1] = (float)im;
        }
    return a;
}

int sr_last_modes(int n_in, int n_out) { return std::max(0, std::min(n_out / 2, n_in / 2 + 1)); }

std::vector<float> sr_mid_fixed(int n_in, int n_out) {
    const int K = (std::min(n_in, n_out) - 1) / 2;
    std::vector<float> a((size_t)n_out * n_in * 2, 0.0f);
    for (int j = 0; j < n_out; ++j)
        for (int hh = 0; hh < n_in; ++hh) {
            long double re = 0, im = 0;
            for (int k = -K; k <= K; ++k) {
                // frequency k on the input grid is bin (k mod n_in), on the output grid bin (k mod n_out)
                long double c1, s1, c2, s2;
                cis(((k % n_out) + n_out) % n_out, j, n_out, c1, s1);   // e^{+i a1}
                cis(((k % n_in) + n_in) % n_in, hh, n_in, c2, s2);      // e^{-i a2} (as in sr_mid)
                re += c1 * c2 + s1 * s2;
                im += s1 * c2 - c1 * s2;
            }
            a[((size_t)j * n_in + hh) * 2] = (float)re;
            a[((size_t)j * n_in + hh) * 2 + 1] = (float)im;
        }
    return a;
}

int sr_last_modes_fixed(int n_in, int n_out) { return (std::min(n_in, n_out) - 1) / 2 + 1; }

// ---------------------------------------------------------------------------------------------
// ATen upsample_bicubic2d_aa weights, fp32 arithmetic in ATen's order of operations
// ---------------------------------------------------------------------------------------------
namespace {
inline float cubic1(float x, float A) { return ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f; }
inline float cubic2(float x, float A) { return ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A; }
inline float aa_cubic(float x) {
    const float a = -0.5f;
    x = std::fabs(x);
    if (x < 1.0f) return cubic1(x, a);
    if (x < 2.0f) return cubic2(x, a);
    return 0.0f;
}
}  // namespace

Banded bicubic_aa(int n_in, int n_out) {
    Banded b;
    b.n_in = n_in;
    b.n_out = n_out;
    b.start.assign(n_out, 0);
    if (n_in == n_out) {  // ATen copies when sizes match
        b.taps = 1;
        b.w.assign(n_out, 1.0f);
        for (int i = 0; i < n_out; ++i) b.start[i] = i;
        return b;
    }
    const float scale = n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.0f;
    const float support = scale >= 1.0f ? 2.0f * scale : 2.0f;
    const float invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
    std::vector<std::vector<float>> rows(n_out);
    int taps = 1;
    for (int i = 0; i < n_out; ++i) {
        const float center = scale * ((float)i + 0.5f);
        int xmin = std::max((int)(int64_t)(center - support + 0.5f), 0);
        int xsize = std::min((int)(int64_t)(center + support + 0.5f), n_in) - xmin;
        if (xsize < 0) xsize = 0;
        std::vector<float>& w = rows[i];
        w.resize(xsize);
        float total = 0.0f;
        for (int j = 0; j < xsize; ++j) {
            w[j] = aa_cubic(((float)(j + xmin) - center + 0.5f) * invscale);
            total += w[j];
        }
        if (total != 0.0f)
            for (int j = 0; j < xsize; ++j) w[j] /= total;
        b.start[i] = xmin;
        taps = std::max(taps, xsize);
    }
    b.taps = taps;
    b.w.assign((size_t)n_out * taps, 0.0f);
    for (int i = 0; i < n_out; ++i) {
        // keep the band inside the input so padded taps never index out of range
        if (b.start[i] + taps > n_in) {
            const int shift = b.start[i] + taps - n_in;
            const int ns = std::max(b.start[i] - shift, 0);
            const int off = b.start[i] - ns;
            for (size_t j = 0; j < rows[i].size(); ++j)
                if ((int)j + off < taps) b.w[(size_t)i * taps + j + off] = rows[i][j];
            b.start[i] = ns;
        } else {
            for (size_t j = 0; j < rows[i].size(); ++j) b.w[(size_t)i * taps + j] = rows[i][j];
        }
    }
    return b;
}

std::vector<float> banded_dense(const Banded& b) {
    std::vector<float> d((size_t)b.n_out * b.n_in, 0.0f);
    for (int i = 0; i < b.n_out; ++i)
        for (int t = 0; t < b.taps; ++t) {
            const int j = b.start[i] + t;
            if (j >= 0 && j < b.n_in) d[(size_t)i * b.n_in + j] += b.w[(size_t)i * b.taps + t];
        }
    return d;
}

BandGroups band_groups(const Banded& b) {
    // nonzero extent of every output row
    std::vector<int> lo(b.n_out, 0), hi(b.n_out, 0);   // [lo, hi)
    for (int i = 0; i < b.n_out; ++i) {
        int f = -1, l = -1;
        for (int t = 0; t < b.taps; ++t)
            if (b.w[(size_t)i * b.taps + t] != 0.0f) { if (f < 0) f = t; l = t; }
        if (f < 0) { lo[i] = hi[i] = std::min(std::max(b.start[i], 0), b.n_in); }
        else { lo[i] = b.start[i] + f; hi[i] = b.start[i] + l + 1; }
    }
    static const int cfgs[3][2] = {{8, 8}, {4, 8}, {4, 16}};
    for (const auto& cfg : cfgs) {
        const int G = cfg[0], W = cfg[1];
        if (W > b.n_in) continue;
        const int ng = (b.n_out + G - 1) / G;
        bool fits = true;
        std::vector<int> gs(ng, 0);
        for (int g = 0; g < ng && fits; ++g) {
            int a = b.n_in, e = 0;
            bool any = false;
            for (int q = 0; q < G; ++q) {
                const int i = g * G + q;
                if (i >= b.n_out || hi[i] <= lo[i]) continue;
                a = std::min(a, lo[i]); e = std::max(e, hi[i]); any = true;
            }
            if (!any) { a = 0; e = 0; }
            if (e - a > W) { fits = false; break; }
            gs[g] = std::max(0, std::min(a, b.n_in - W));
            if (g > 0 && gs[g] < gs[g - 1]) fits = false;   // tiles assume monotone windows
        }
        if (!fits) continue;
        BandGroups r;
        r.ok = true; r.G = G; r.W = W; r.ng = ng; r.n_in = b.n_in; r.n_out = b.n_out;
        r.gstart = gs;
        r.D.assign((size_t)ng * W * G, 0.0f);
        for (int i = 0; i < b.n_out; ++i) {
            const int g = i / G, q = i % G;
            for (int t = 0; t < b.taps; ++t) {
                const float w = b.w[(size_t)i * b.taps + t];
                if (w == 0.0f) continue;
                const int u = b.start[i] + t - gs[g];   // in [0, W) by construction
                r.D[((size_t)g * W + u) * G + q] += w;
            }
        }
        return r;
    }
    return BandGroups();
}

int BandGroups::span(int groups_per_tile) const {
    int best = 0;
    for (int g0 = 0; g0 < ng; g0 += groups_per_tile) {
        const int g1 = std::min(g0 + groups_per_tile, ng) - 1;
        best = std::max(best, gstart[g1] + W - gstart[g0]);
    }
    return best;
}

std::vector<float> band_groups_dense(const BandGroups& g) {
    std::vector<float> d((size_t)g.n_out * g.n_in, 0.0f);
    for (int gi = 0; gi < g.ng; ++gi)
        for (int u = 0; u < g.W; ++u)
            for (int q = 0; q < g.G; ++q) {
                const int i = gi * g.G + q;
                if (i < g.n_out) d[(size_t)i * g.n_in + g.gstart[gi] + u] += g.D[((size_t)gi * g.W + u) * g.G + q];
            }
    return d;
}

Banded banded_transpose(const Banded& b) {
    Banded t;
    t.n_in = b.n_out;
    t.n_out = b.n_in;
    std::vector<int> lo(b.n_in, b.n_out), hi(b.n_in, -1);
    for (int i = 0; i < b.n_out; ++i)
        for (int k = 0; k < b.taps; ++k) {
            const int j = b.start[i] + k;
            if (j < 0 || j >= b.n_in || b.w[(size_t)i * b.taps + k] == 0.0f) continue;
            lo[j] = std::min(lo[j], i);
            hi[j] = std::max(hi[j], i);
        }
    int taps = 1;
    for (int j = 0; j < b.n_in; ++j)
        if (hi[j] >= lo[j]) taps = std::max(taps, hi[j] - lo[j] + 1);
    t.taps = taps;
    t.start.assign(b.n_in, 0);
    t.w.assign((size_t)b.n_in * taps, 0.0f);
    for (int j = 0; j < b.n_in; ++j) {
        int s = hi[j] >= lo[j] ? lo[j] : 0;
        if (s + taps > b.n_out) s = std::max(b.n_out - taps, 0);
        t.start[j] = s;
    }
    for (int i = 0; i < b.n_out; ++i)
        for (int k = 0; k < b.taps; ++k) {
            const int j = b.start[i] + k;
            const float w = b.w[(size_t)i * b.taps + k];
            if (j < 0 || j >= b.n_in || w == 0.0f) continue;
            const int pos = i - t.start[j];
            if (pos >= 0 && pos < taps) t.w[(size_t)j * taps + pos] += w;
        }
    return t;
}

}  // namespace uno
