// TEST INFRASTRUCTURE ONLY.  Plain-loop implementation of uno_b200/csrc/backend.h so that the
// orchestration in uno_api.cpp (plan matrices, strides, corner maps, adjoints, epilogue selection)
// can be exercised by the CPU test-suite on a machine with no GPU.  Built by tests/hostemu/build.py
// into tests/hostemu/libuno_hostemu.so; the product (uno_b200/_lib.py) never loads it and
// uno_backend_name() of this build returns "host-emulation".
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../uno_b200/csrc/backend.h"

namespace uno {

int be_upload(void** dptr, const void* host, size_t bytes) {
    *dptr = malloc(bytes ? bytes : 1);
    if (!*dptr) return 1;
    memcpy(*dptr, host, bytes);
    return 0;
}
void be_free(void* d) { free(d); }
int be_memset(void* d, int v, size_t bytes, stream_t) { memset(d, v, bytes); return 0; }
const char* be_name() { return "host-emulation"; }
int be_current_device() { return 0; }
void be_range_push(const char*) {}
void be_range_pop() {}
const char* be_error_string(int) { return "host emulation error"; }
stream_t be_side_stream(int) { return nullptr; }
int be_fork(stream_t, stream_t) { return 0; }
int be_join(stream_t, stream_t) { return 0; }
size_t be_profile_report(char* buf, size_t cap) { if (buf && cap > 2) { buf[0] = '{'; buf[1] = '}'; buf[2] = 0; } return 2; }
long be_launch_count() { return 0; }
// scopes are recorded (no timing here: ms = 0) so that the labels and the algorithmic-byte accounting of uno_api.cpp can be
// checked on the CPU
namespace {
struct HostScope { std::string label; long calls; double bytes, flops; };
std::vector<HostScope> g_scopes;
bool g_prof = false;
}
void be_profile_enable(int on) { if (on && !g_prof) g_scopes.clear(); g_prof = on != 0; }
int be_profile_enabled() { return g_prof ? 1 : 0; }
void be_profile_scope_begin(const char* label, double bytes, double flops) {
    if (!g_prof) return;
    for (auto& s : g_scopes)
        if (s.label == label) { s.calls += 1; s.bytes += bytes; s.flops += flops; return; }
    g_scopes.push_back(HostScope{label, 1, bytes, flops});
}
void be_profile_scope_end() {}
size_t be_profile_report_scopes(char* buf, size_t cap) {
    std::string out = "{";
    for (size_t i = 0; i < g_scopes.size(); ++i) {
        char line[384];
        snprintf(line, sizeof line, "%s\"%s\": {\"calls\": %ld, \"launches\": 0, \"ms\": 0.0, \"bytes\": %.0f, \"flops\": %.0f}",
                 i ? ", " : "", g_scopes[i].label.c_str(), g_scopes[i].calls, g_scopes[i].bytes, g_scopes[i].flops);
        out += line;
    }
    out += "}";
    if (buf && cap) {
        size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return out.size();
}

static inline float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
static inline float gelu_grad_f(float x) {
    return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * expf(-0.5f * x * x);
}

int be_gemm(const GemmArgs& a, stream_t) {
    for (int b = 0; b < a.batch; ++b) {
        const float* A = a.A + b * a.sA;
        const float* B = a.B + b * a.sB;
        float* C = a.C + b * a.sC;
        float* C2 = a.C2 ? a.C2 + b * a.sC : nullptr;
        for (int m = 0; m < a.M; ++m)
            for (int n = 0; n < a.N; ++n) {
                double acc = 0;
                for (int k = 0; k < a.K; ++k) acc += (double)A[m * a.a_rs + k * a.a_cs] * B[k * a.ldb + n];
                float v = (float)acc;
                float* c = &C[m * a.ldc + n];
                switch (a.epi) {
                    case EPI_STORE: *c = v + (a.bias ? a.bias[m] : 0.0f); break;
                    case EPI_ACCUM: *c += v; break;
                    case EPI_ACCUM_GELU: *c += v; C2[m * a.ldc + n] = gelu_f(*c); break;
                    case EPI_ACCUM_GELU_INPLACE: *c = gelu_f(*c + v); break;
                    default: return 1;
                }
            }
    }
    return 0;
}

int be_gemm_nt_atomic(const GemmNtArgs& a, stream_t) {
    for (int m = 0; m < a.M; ++m)
        for (int n = 0; n < a.N; ++n) {
            double acc = 0;
            for (int b = 0; b < a.batch; ++b) {
                const float* A = a.A + b * a.sA + m * a.lda;
                const float* B = a.B + b * a.sB + n * a.ldb;
                for (int k = 0; k < a.K; ++k) acc += (double)A[k] * B[k];
            }
            a.C[m * a.ldc + n] += (float)acc;
        }
    return 0;
}

int be_mid(const MidArgs& a, stream_t) {
    for (long o = 0; o < a.O; ++o)
        for (int j = 0; j < a.J; ++j)
            for (int i = 0; i < a.I; ++i) {
                double re = 0, im = 0;
                for (int h = 0; h < a.H; ++h) {
                    const float mr = a.Mat[((long)j * a.H + h) * 2], mi = a.Mat[((long)j * a.H + h) * 2 + 1];
                    const float* x = a.X + ((o * a.H + h) * a.I + i) * 2;
                    re += (double)mr * x[0] - (double)mi * x[1];
                    im += (double)mr * x[1] + (double)mi * x[0];
                }
                float* y = a.Y + ((o * a.J + j) * a.I + i) * 2;
                y[0] = (float)re;
                y[1] = (float)im;
            }
    return 0;
}

int be_cmm(const CmmArgs& a, stream_t) {
  for (int cr = 0; cr < a.ncorner; ++cr)
    for (int m = 0; m < a.M; ++m)
        for (int n = 0; n < a.N; ++n)
            for (int qo = 0; qo < a.q_outer; ++qo)
                for (int qi = 0; qi < a.q_inner; ++qi) {
                    double re = 0, im = 0;
                    for (int k = 0; k < a.K; ++k) {
                        const float* pa = a.A[cr] + 2 * (m * a.a_sm + k * a.a_sk + qo * a.a_sqo + qi);
                        const float* pb = a.B[cr] + 2 * (k * a.b_sk + n * a.b_sn + qo * a.b_sqo + qi);
                        const double ar = pa[0], ai = a.conjA ? -pa[1] : pa[1];
                        const double br = pb[0], bi = a.conjB ? -pb[1] : pb[1];
                        re += ar * br - ai * bi;
                        im += ar * bi + ai * br;
                    }
                    float* pc = a.C[cr] + 2 * (m * a.c_sm + n * a.c_sn + qo * a.c_sqo + qi);
                    pc[0] = (float)re;
                    pc[1] = (float)im;
                }
    return 0;
}

int be_banded(const BandedArgs& a, stream_t) {
    for (long o = 0; o < a.outer; ++o)
        for (int j = 0; j < a.n_out; ++j)
            for (int i = 0; i < a.inner; ++i) {
                double acc = 0;
                for (int t = 0; t < a.taps; ++t)
                    acc += (double)a.w[(long)j * a.taps + t] * a.x[(o * a.n_in + a.start[j] + t) * a.inner + i];
                a.y[(o * a.n_out + j) * a.inner + i] = (float)acc;
            }
    return 0;
}

int be_banded2d(const Banded2DArgs& a, stream_t) {
    if (a.D0 && a.D1) {
        // the register-blocked images (what the CUDA kernel consumes): window of W inputs per group of G outputs.
        // Walks the same tiles as the kernel so that the tile geometry handed over by the plan is exercised too.
        const int TW = a.G1 == 8 ? 128 : 64;
        const int TH = a.tile_groups0 * a.G0, NG1 = TW / a.G1;
        for (long p = 0; p < a.planes; ++p)
            for (int ga0 = 0; ga0 < a.ng0; ga0 += a.tile_groups0)
                for (int ga1 = 0; ga1 < a.ng1; ga1 += NG1) {
                    const int ngh = std::min(a.tile_groups0, a.ng0 - ga0), ngw = std::min(NG1, a.ng1 - ga1);
                    const int r0 = a.gs0[ga0], c0 = a.gs1[ga1];
                    const int rin = a.gs0[ga0 + ngh - 1] + a.W0 - r0, cin = a.gs1[ga1 + ngw - 1] + a.W1 - c0;
                    if (rin > a.tile_span0 || cin > a.tile_span1 || r0 + rin > a.n_in0 || c0 + cin > a.n_in1 || TH < 1) return 1;
                    std::vector<double> mid((size_t)rin * TW, 0.0);
                    for (int r = 0; r < rin; ++r)
                        for (int g = 0; g < ngw; ++g)
                            for (int q = 0; q < a.G1; ++q) {
                                double acc = 0;
                                const int cs = a.gs1[ga1 + g] - c0;
                                if (cs < 0) return 1;
                                for (int u = 0; u < a.W1; ++u)
                                    acc += (double)a.D1[((long)(ga1 + g) * a.W1 + u) * a.G1 + q] * a.x[(p * a.n_in0 + r0 + r) * a.n_in1 + c0 + cs + u];
                                mid[(size_t)r * TW + g * a.G1 + q] = acc;
                            }
                    for (int g = 0; g < ngh; ++g)
                        for (int q = 0; q < a.G0; ++q) {
                            const int i = (ga0 + g) * a.G0 + q;
                            if (i >= a.n_out0) continue;
                            const int rs = a.gs0[ga0 + g] - r0;
                            if (rs < 0) return 1;
                            for (int jc = 0; jc < TW; ++jc) {
                                const int j = ga1 * a.G1 + jc;
                                if (j >= a.n_out1) continue;
                                double acc = 0;
                                for (int u = 0; u < a.W0; ++u) acc += (double)a.D0[((long)(ga0 + g) * a.W0 + u) * a.G0 + q] * mid[(size_t)(rs + u) * TW + jc];
                                a.y[(p * a.n_out0 + i) * a.n_out1 + j] = (float)acc;
                            }
                        }
                }
        return 0;
    }
    for (long p = 0; p < a.planes; ++p)
        for (int i = 0; i < a.n_out0; ++i)
            for (int j = 0; j < a.n_out1; ++j) {
                double acc = 0;
                for (int t = 0; t < a.taps0; ++t) {
                    double row = 0;
                    const float* xr = a.x + (p * a.n_in0 + a.start0[i] + t) * a.n_in1 + a.start1[j];
                    for (int u = 0; u < a.taps1; ++u) row += (double)a.w1[(long)j * a.taps1 + u] * xr[u];
                    acc += (double)a.w0[(long)i * a.taps0 + t] * row;
                }
                a.y[(p * a.n_out0 + i) * a.n_out1 + j] = (float)acc;
            }
    return 0;
}

int be_gelu_fwd(const float* pre, float* y, size_t n, stream_t) {
    for (size_t i = 0; i < n; ++i) y[i] = gelu_f(pre[i]);
    return 0;
}
// element i of plane p of the upstream gradient: the sum of its (up to two) strided sources
static inline float upgrad_at(const UpGrad& gy, long p, int C, long L, long i) {
    const long b = p / C, c = p % C;
    float v = gy.p0[b * gy.bs0 + c * L + i];
    if (gy.p1) v += gy.p1[b * gy.bs1 + c * L + i];
    return v;
}
int be_gelu_bwd_bias(const UpGrad& gy, const float* pre, float* g, long planes, int C, long L, float* gbias, float alpha, stream_t) {
    for (long p = 0; p < planes; ++p) {
        double s = 0;
        for (long i = 0; i < L; ++i) {
            const float u = upgrad_at(gy, p, C, L, i);
            g[p * L + i] = pre ? u * gelu_grad_f(pre[p * L + i]) : u;
            s += g[p * L + i];
        }
        if (gbias) gbias[p % C] += alpha * (float)s;
    }
    return 0;
}

int be_plane_stats(const float* x, float* stats, long planes, long L, float eps, stream_t) {
    for (long p = 0; p < planes; ++p) {
        double s = 0;
        for (long i = 0; i < L; ++i) s += x[p * L + i];
        const double mu = s / L;
        double v = 0;
        for (long i = 0; i < L; ++i) { double d = x[p * L + i] - mu; v += d * d; }
        stats[2 * p] = (float)mu;
        stats[2 * p + 1] = (float)(1.0 / std::sqrt(v / L + eps));
    }
    return 0;
}

int be_norm_act_fwd(const float* x, const float* stats, const float* gamma, const float* beta,
                    float* y, long planes, int C, long L, int non_lin, stream_t) {
    for (long p = 0; p < planes; ++p) {
        const int c = (int)(p % C);
        for (long i = 0; i < L; ++i) {
            float n = (x[p * L + i] - stats[2 * p]) * stats[2 * p + 1] * gamma[c] + beta[c];
            y[p * L + i] = non_lin ? gelu_f(n) : n;
        }
    }
    return 0;
}

int be_norm_fused_fwd(const float* x, float* stats, const float* gamma, const float* beta, float* y, long planes, int C,
                      long L, float eps, int non_lin, stream_t s) {
    int rc = be_plane_stats(x, stats, planes, L, eps, s);
    return rc ? rc : be_norm_act_fwd(x, stats, gamma, beta, y, planes, C, L, non_lin, s);
}

int be_norm_act_bwd(const UpGrad& gy, const float* x, const float* stats, const float* gamma,
                    const float* beta, float* g, float* ggamma, float* gbeta, long planes, int C,
                    long L, int non_lin, stream_t) {
    for (long p = 0; p < planes; ++p) {
        const int c = (int)(p % C);
        const float mu = stats[2 * p], rstd = stats[2 * p + 1];
        double s1 = 0, s2 = 0;
        for (long i = 0; i < L; ++i) {
            const float xh = (x[p * L + i] - mu) * rstd;
            const float up = upgrad_at(gy, p, C, L, i);
            const float gn = non_lin ? up * gelu_grad_f(xh * gamma[c] + beta[c]) : up;
            s1 += gn;
            s2 += (double)gn * xh;
        }
        ggamma[c] += (float)s2;
        gbeta[c] += (float)s1;
        const float m1 = (float)(s1 / L), m2 = (float)(s2 / L);
        for (long i = 0; i < L; ++i) {
            const float xh = (x[p * L + i] - mu) * rstd;
            const float up = upgrad_at(gy, p, C, L, i);
            const float gn = non_lin ? up * gelu_grad_f(xh * gamma[c] + beta[c]) : up;
            g[p * L + i] = gamma[c] * rstd * (gn - m1 - xh * m2);
        }
    }
    return 0;
}

int be_channel_sum(const float* x, float* out, long planes, int C, long L, float alpha, stream_t) {
    for (long p = 0; p < planes; ++p) {
        double s = 0;
        for (long i = 0; i < L; ++i) s += x[p * L + i];
        out[p % C] += alpha * (float)s;
    }
    return 0;
}

int be_add_channel_const(float* y, const float* v, float alpha, long planes, int C, long L, stream_t) {
    for (long p = 0; p < planes; ++p)
        for (long i = 0; i < L; ++i) y[p * L + i] += v[p % C] * alpha;
    return 0;
}

// ---- model glue: lift / project as plain loops (double accumulation) ------------------------------------
namespace {
struct Geo { long nraw, npad; };
inline Geo geo(const int* n, const int* N) { return {(long)n[0] * n[1] * n[2], (long)N[0] * N[1] * N[2]}; }
inline long padded_index(const int* n, const int* N, const int* lo, long rp) {
    const int r2 = (int)(rp % n[2]);
    const long t = rp / n[2];
    const int r1 = (int)(t % n[1]), r0 = (int)(t / n[1]);
    return ((long)(r0 + lo[0]) * N[1] + (r1 + lo[1])) * N[2] + (r2 + lo[2]);
}
}  // namespace

int be_lift_supported(const LiftArgs& a) {
    return a.raw_ch + a.grid_ch <= 16 && a.hid <= 32 && a.out_ch <= 64;
}

int be_lift_fwd(const LiftArgs& a, stream_t) {
    const Geo g = geo(a.n, a.N);
    const int cin = a.raw_ch + a.grid_ch;
    memset(a.h, 0, sizeof(float) * a.batch * a.out_ch * g.npad);
    for (long b = 0; b < a.batch; ++b)
        for (long rp = 0; rp < g.nraw; ++rp) {
            double in[64], a0[64];
            for (int c = 0; c < cin; ++c)
                in[c] = c < a.raw_ch ? a.a[(b * g.nraw + rp) * a.raw_ch + c] : a.grid[rp * a.grid_ch + (c - a.raw_ch)];
            for (int k = 0; k < a.hid; ++k) {
                double s = a.b_a[k];
                for (int c = 0; c < cin; ++c) s += (double)a.w_a[k * cin + c] * in[c];
                a0[k] = gelu_f((float)s);
            }
            const long pp = padded_index(a.n, a.N, a.lo, rp);
            for (int o = 0; o < a.out_ch; ++o) {
                double s = a.b_b[o];
                for (int k = 0; k < a.hid; ++k) s += (double)a.w_b[o * a.hid + k] * a0[k];
                a.h[(b * a.out_ch + o) * g.npad + pp] = gelu_f((float)s);
            }
        }
    return 0;
}

int be_lift_bwd(const LiftArgs& a, stream_t) {
    const Geo g = geo(a.n, a.N);
    const int cin = a.raw_ch + a.grid_ch;
    std::vector<double> gwa((size_t)a.hid * cin, 0.0), gba(a.hid, 0.0), gwb((size_t)a.out_ch * a.hid, 0.0), gbb(a.out_ch, 0.0);
    for (long b = 0; b < a.batch; ++b)
        for (long rp = 0; rp < g.nraw; ++rp) {
            double in[64], pre0[64], a0[64], da0[64];
            for (int c = 0; c < cin; ++c)
                in[c] = c < a.raw_ch ? a.a[(b * g.nraw + rp) * a.raw_ch + c] : a.grid[rp * a.grid_ch + (c - a.raw_ch)];
            for (int k = 0; k < a.hid; ++k) {
                double s = a.b_a[k];
                for (int c = 0; c < cin; ++c) s += (double)a.w_a[k * cin + c] * in[c];
                pre0[k] = s; a0[k] = gelu_f((float)s); da0[k] = 0;
            }
            const long pp = padded_index(a.n, a.N, a.lo, rp);
            for (int o = 0; o < a.out_ch; ++o) {
                double s = a.b_b[o];
                for (int k = 0; k < a.hid; ++k) s += (double)a.w_b[o * a.hid + k] * a0[k];
                const double d1 = ((double)a.gh[(b * a.out_ch + o) * g.npad + pp] + (a.gh2 ? (double)a.gh2[(b * a.out_ch + o) * g.npad + pp] : 0.0)) *
                                  gelu_grad_f((float)s);
                gbb[o] += d1;
                for (int k = 0; k < a.hid; ++k) { gwb[o * a.hid + k] += d1 * a0[k]; da0[k] += d1 * a.w_b[o * a.hid + k]; }
            }
            for (int k = 0; k < a.hid; ++k) {
                const double d0 = da0[k] * gelu_grad_f((float)pre0[k]);
                da0[k] = d0;
                gba[k] += d0;
                for (int c = 0; c < cin; ++c) gwa[k * cin + c] += d0 * in[c];
            }
            if (a.ga)
                for (int c = 0; c < a.raw_ch; ++c) {
                    double s = 0;
                    for (int k = 0; k < a.hid; ++k) s += da0[k] * a.w_a[k * cin + c];
                    a.ga[(b * g.nraw + rp) * a.raw_ch + c] = (float)s;
                }
        }
    for (size_t i = 0; i < gwa.size(); ++i) a.gw_a[i] += (float)gwa[i];
    for (size_t i = 0; i < gba.size(); ++i) a.gb_a[i] += (float)gba[i];
    for (size_t i = 0; i < gwb.size(); ++i) a.gw_b[i] += (float)gwb[i];
    for (size_t i = 0; i < gbb.size(); ++i) a.gb_b[i] += (float)gbb[i];
    return 0;
}

int be_proj_supported(const ProjArgs& a) {
    int ctot = 0;
    for (int s = 0; s < a.nsrc; ++s) ctot += a.src_ch[s];
    return a.nsrc >= 1 && a.nsrc <= 4 && ctot <= 64 && a.hid <= 128 && a.out_ch <= 4;
}

int be_proj_fwd(const ProjArgs& a, stream_t) {
    const Geo g = geo(a.n, a.N);
    int ctot = 0;
    for (int s = 0; s < a.nsrc; ++s) ctot += a.src_ch[s];
    for (long b = 0; b < a.batch; ++b)
        for (long rp = 0; rp < g.nraw; ++rp) {
            const long pp = padded_index(a.n, a.N, a.lo, rp);
            double in[64], o[4] = {0, 0, 0, 0};
            int c = 0;
            for (int s = 0; s < a.nsrc; ++s)
                for (int cl = 0; cl < a.src_ch[s]; ++cl) in[c++] = a.src[s][(b * a.src_ch[s] + cl) * g.npad + pp];
            for (int n = 0; n < a.hid; ++n) {
                double s = a.b1[n];
                for (int cc = 0; cc < ctot; ++cc) s += (double)a.w1[n * ctot + cc] * in[cc];
                if (a.pre_out) a.pre_out[(size_t)n * a.batch * g.nraw + b * g.nraw + rp] = (float)s;
                const double act = gelu_f((float)s);
                for (int q = 0; q < a.out_ch; ++q) o[q] += (double)a.w2[q * a.hid + n] * act;
            }
            for (int q = 0; q < a.out_ch; ++q) a.out[(b * g.nraw + rp) * a.out_ch + q] = (float)(o[q] + a.b2[q]);
        }
    return 0;
}

int be_proj_bwd(const ProjArgs& a, stream_t) {
    const Geo g = geo(a.n, a.N);
    int ctot = 0;
    for (int s = 0; s < a.nsrc; ++s) ctot += a.src_ch[s];
    for (int s = 0; s < a.nsrc; ++s)
        if (a.gsrc[s]) memset(a.gsrc[s], 0, sizeof(float) * a.batch * a.src_ch[s] * g.npad);
    std::vector<double> gw1((size_t)a.hid * ctot, 0.0), gb1(a.hid, 0.0), gw2((size_t)a.out_ch * a.hid, 0.0), gb2(a.out_ch, 0.0);
    for (long b = 0; b < a.batch; ++b)
        for (long rp = 0; rp < g.nraw; ++rp) {
            const long pp = padded_index(a.n, a.N, a.lo, rp);
            double in[64], din[64];
            int c = 0;
            for (int s = 0; s < a.nsrc; ++s)
                for (int cl = 0; cl < a.src_ch[s]; ++cl) { in[c] = a.src[s][(b * a.src_ch[s] + cl) * g.npad + pp]; din[c++] = 0; }
            const float* go = a.gout + (b * g.nraw + rp) * a.out_ch;
            for (int q = 0; q < a.out_ch; ++q) gb2[q] += go[q];
            for (int n = 0; n < a.hid; ++n) {
                double s = a.b1[n];
                for (int cc = 0; cc < ctot; ++cc) s += (double)a.w1[n * ctot + cc] * in[cc];
                if (a.pre_in) s = a.pre_in[(size_t)n * a.batch * g.nraw + b * g.nraw + rp];   // what the forward pass saved
                const double act = gelu_f((float)s);
                double t = 0;
                for (int q = 0; q < a.out_ch; ++q) { t += (double)go[q] * a.w2[q * a.hid + n]; gw2[q * a.hid + n] += go[q] * act; }
                const double dp = t * gelu_grad_f((float)s);
                gb1[n] += dp;
                for (int cc = 0; cc < ctot; ++cc) { gw1[n * ctot + cc] += dp * in[cc]; din[cc] += dp * a.w1[n * ctot + cc]; }
            }
            c = 0;
            for (int s = 0; s < a.nsrc; ++s)
                for (int cl = 0; cl < a.src_ch[s]; ++cl, ++c)
                    if (a.gsrc[s]) a.gsrc[s][(b * a.src_ch[s] + cl) * g.npad + pp] = (float)din[c];
        }
    for (size_t i = 0; i < gw1.size(); ++i) a.gw1[i] += (float)gw1[i];
    for (size_t i = 0; i < gb1.size(); ++i) a.gb1[i] += (float)gb1[i];
    for (size_t i = 0; i < gw2.size(); ++i) a.gw2[i] += (float)gw2[i];
    for (size_t i = 0; i < gb2.size(); ++i) a.gb2[i] += (float)gb2[i];
    return 0;
}

// ---- training-step ops: plain loops following Adam.py:23-52 and utilities3.py:86-100 -------------------------------------
int be_adam_step(const AdamTensor* t, int n, const AdamHyper& h, stream_t) {
    const double bc1 = 1.0 - pow(h.beta1, (double)h.step), bc2 = 1.0 - pow(h.beta2, (double)h.step);
    const float step_size = (float)(h.lr / bc1), s2 = (float)sqrt(bc2);
    const float omb1 = (float)(1.0 - h.beta1), omb2 = (float)(1.0 - h.beta2);
    const float beta1 = (float)h.beta1, beta2 = (float)h.beta2, eps = (float)h.eps, wd = (float)h.weight_decay;
    for (int k = 0; k < n; ++k) {
        const AdamTensor& a = t[k];
        for (long i = 0; i < a.numel; i += (a.is_complex ? 2 : 1)) {
            const int w = a.is_complex ? 2 : 1;
            float g[2] = {0, 0};
            float sq = 0;
            for (int c = 0; c < w; ++c) {
                g[c] = a.grad[i + c] + wd * a.param[i + c];
                a.exp_avg[i + c] = a.exp_avg[i + c] * beta1 + omb1 * g[c];
                sq += g[c] * g[c];
            }
            a.exp_avg_sq[i] = a.exp_avg_sq[i] * beta2 + omb2 * sq;
            if (a.is_complex) a.exp_avg_sq[i + 1] *= beta2;
            float vhat = a.exp_avg_sq[i];
            if (h.amsgrad && !a.is_complex) { vhat = std::max(a.max_exp_avg_sq[i], vhat); a.max_exp_avg_sq[i] = vhat; }
            const float denom = sqrtf(vhat) / s2 + eps;
            for (int c = 0; c < w; ++c) a.param[i + c] -= step_size * (a.exp_avg[i + c] / denom);
        }
    }
    return 0;
}

int be_lp_loss_fwd(const float* x, const float* y, int B, long N, int reduction, float* loss, float* norms, double*, stream_t) {
    double tot = 0;
    for (int b = 0; b < B; ++b) {
        double d2 = 0, y2 = 0;
        for (long i = 0; i < N; ++i) { const double d = (double)x[b * N + i] - y[b * N + i]; d2 += d * d; y2 += (double)y[b * N + i] * y[b * N + i]; }
        norms[2 * b] = (float)sqrt(d2);
        norms[2 * b + 1] = (float)sqrt(y2);
        const float r = norms[2 * b] / norms[2 * b + 1];
        if (reduction == 0) loss[b] = r;
        tot += r;
    }
    if (reduction == 1) loss[0] = (float)tot;
    if (reduction == 2) loss[0] = (float)(tot / B);
    return 0;
}
int be_lp_loss_bwd(const float* x, const float* y, const float* norms, const float* gl, int B, long N, int reduction, float* gx,
                   stream_t) {
    for (int b = 0; b < B; ++b) {
        float scale = reduction == 0 ? gl[b] : gl[0];
        if (reduction == 2) scale /= (float)B;
        scale = norms[2 * b] > 0.0f ? scale / (norms[2 * b] * norms[2 * b + 1]) : 0.0f;   // torch masks a zero residual norm
        for (long i = 0; i < N; ++i) gx[b * N + i] = (x[b * N + i] - y[b * N + i]) * scale;
    }
    return 0;
}

}  // namespace uno
