// C-ABI entry points (include/uno_b200.h) and the orchestration of the primitive kernels.
//
// One SpectralConv call is   analyse (real grid -> kept modes)  ->  per-mode channel contraction  ->
// synthesise (kept modes -> real grid).  Every linear stage is multiplication by a small constant
// matrix from plan.cpp in which the reference's semantics are baked (1/N scaling, corner/mode maps,
// last-writer-wins on overlapping output corners, C2R dropping Im of DC/Nyquist); the backward pass
// runs the same kernels with the (conjugate-)transposed matrices.
#include "../../include/uno_b200.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "backend.h"
#include "plan.h"

using namespace uno;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define UNO_TRY(expr)                 \
    do {                              \
        int _rc = (expr);             \
        if (_rc != 0) return _rc;     \
    } while (0)

// backend (kernel-launch) calls return 0 or a backend error code
#define BE_TRY(expr)                                                                  \
    do {                                                                              \
        int _rc = (expr);                                                             \
        if (_rc != 0) return fail(UNO_ECUDA, "%s: %s", #expr, be_error_string(_rc));  \
    } while (0)

inline size_t align64(size_t n) { return (n + 63) & ~size_t(63); }

// bump allocator over the caller's workspace (floats, 256-byte aligned regions)
struct Arena {
    float* base;
    size_t cap, off = 0;
    bool ok = true;
    Arena(void* p, size_t bytes) : base((float*)p), cap(bytes / sizeof(float)) {}
    float* take(size_t n) {
        n = align64(n);
        if (off + n > cap) { ok = false; return nullptr; }
        float* r = base + off;
        off += n;
        return r;
    }
};

struct DevMat {
    float* d = nullptr;
    ~DevMat() { if (d) be_free(d); }
    int upload(const std::vector<float>& h) {
        void* p = nullptr;
        int rc = be_upload(&p, h.data(), h.size() * sizeof(float));
        d = (float*)p;
        return rc;
    }
};

struct DevBand {
    int n_in = 0, n_out = 0, taps = 0;
    int span32 = 0, span64 = 0;   // max input extent of a 32- / 64-output tile
    int* start = nullptr;
    float* w = nullptr;
    // register-blocked image (plan.h BandGroups); G == 0 when the band does not fit one
    int G = 0, W = 0, ng = 0;
    int tile_groups_rows = 0, tile_span_rows = 0;   // tile height in groups when this band acts on rows, and its window
    int tile_span_cols = 0;                         // window of a 64-column tile when it acts on columns
    int* gstart = nullptr;
    float* D = nullptr;
    ~DevBand() {
        if (start) be_free(start);
        if (w) be_free(w);
        if (gstart) be_free(gstart);
        if (D) be_free(D);
    }
    int upload_groups(const Banded& b) {
        const BandGroups g = band_groups(b);
        if (!g.ok) return 0;
        G = g.G; W = g.W; ng = g.ng;
        // rows: the tile height (in groups, <= 32, window <= 128 rows so that two CTAs still share an SM) that wastes the fewest lanes
        // when a warp sweeps the window rows.  (Up to 8 groups until round 2: 72-row windows for the 2x down-sampling levels, a
        // third of the lanes of pass A idle in the last sweep; 14 groups give 120-row windows.)
        double best = -1.0;
        for (int n = std::min(32, ng); n >= 1; --n) {
            const int sp = g.span(n);
            if (sp > 128 && n > 1) continue;
            const double eff = (double)(n * G) / (32.0 * ((sp + 31) / 32));
            if (eff > best + 1e-9) { best = eff; tile_groups_rows = n; tile_span_rows = sp; }
        }
        tile_span_cols = g.span((G == 8 ? 128 : 64) / G);   // columns per tile: resample2d.cuh rs_tile_w
        void* p = nullptr;
        int rc = be_upload(&p, g.gstart.data(), g.gstart.size() * sizeof(int));
        gstart = (int*)p;
        if (rc) return rc;
        rc = be_upload(&p, g.D.data(), g.D.size() * sizeof(float));
        D = (float*)p;
        return rc;
    }
    int upload(const Banded& b) {
        n_in = b.n_in; n_out = b.n_out; taps = b.taps;
        auto span = [&](int tile) {
            int best = 0;
            for (int i0 = 0; i0 < b.n_out; i0 += tile) {
                const int last = std::min(i0 + tile, b.n_out) - 1;
                best = std::max(best, b.start[last] + b.taps - b.start[i0]);
            }
            return best;
        };
        span32 = span(32);
        span64 = span(64);
        void* p = nullptr;
        int rc = be_upload(&p, b.start.data(), b.start.size() * sizeof(int));
        start = (int*)p;
        if (rc) return rc;
        rc = be_upload(&p, b.w.data(), b.w.size() * sizeof(float));
        w = (float*)p;
        if (rc) return rc;
        return upload_groups(b);
    }
};

struct MidOp { const float* mat; int n_in, n_out; };   // complex [n_out x n_in]

// ---------------------------------------------------------------------------------------------------
// analyse: real x [P, n_0..n_{k-1}, n_last] -> complex out [P, J_0..J_{k-1}, m]
// ---------------------------------------------------------------------------------------------------
size_t analyse_stage_floats(long P, int nmid, const MidOp* mids, int m) {
    size_t best = 0;
    long dims[2];
    for (int a = 0; a < nmid; ++a) dims[a] = mids[a].n_in;
    auto total = [&]() { size_t t = (size_t)P * 2 * m; for (int a = 0; a < nmid; ++a) t *= dims[a]; return t; };
    best = total();
    for (int a = nmid - 1; a >= 0; --a) { dims[a] = mids[a].n_out; best = std::max(best, total()); }
    return best;
}

int analyse(const float* x, long P, int nmid, const MidOp* mids, int n_last, const float* last_mat,
            int m, float* out, float* ws0, float* ws1, stream_t st) {
    long cur[2] = {1, 1};
    long R = P;
    for (int a = 0; a < nmid; ++a) { cur[a] = mids[a].n_in; R *= cur[a]; }
    float* dst = nmid == 0 ? out : ws0;
    GemmArgs g;
    g.A = x; g.a_rs = n_last; g.a_cs = 1;
    g.B = last_mat; g.ldb = 2 * m;
    g.C = dst; g.ldc = 2 * m;
    g.M = (int)R; g.N = 2 * m; g.K = n_last;
    g.tag = "dft_last_analysis";
    g.b_const = true;
    BE_TRY(be_gemm(g, st));
    const float* src = dst;
    for (int a = nmid - 1; a >= 0; --a) {
        long O = P, I = m;
        for (int b = 0; b < a; ++b) O *= cur[b];
        for (int b = a + 1; b < nmid; ++b) I *= cur[b];
        float* d2 = (a == 0) ? out : (src == ws0 ? ws1 : ws0);
        MidArgs ma;
        ma.X = src; ma.Mat = mids[a].mat; ma.Y = d2;
        ma.O = (int)O; ma.H = mids[a].n_in; ma.J = mids[a].n_out; ma.I = (int)I;
        BE_TRY(be_mid(ma, st));
        cur[a] = mids[a].n_out;
        src = d2;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// synthesise: complex in [P, J_0..J_{k-1}, m] -> real y [P, n_0..n_{k-1}, n_last]
// ---------------------------------------------------------------------------------------------------
size_t synth_stage_floats(long P, int nmid, const MidOp* mids, int m) {
    long dims[2];
    for (int a = 0; a < nmid; ++a) dims[a] = mids[a].n_in;
    auto total = [&]() { size_t t = (size_t)P * 2 * m; for (int a = 0; a < nmid; ++a) t *= dims[a]; return t; };
    size_t best = total();
    for (int a = 0; a < nmid; ++a) { dims[a] = mids[a].n_out; best = std::max(best, total()); }
    return best;
}

// `join_side`: a side stream whose work (it produced the tensor the epilogue accumulates onto) must be complete
// before the last stage runs; the leading-axis stages before it still overlap with that stream.
int synthesise(const float* in, long P, int nmid, const MidOp* mids, int m, const float* last_mat,
               int n_last, float* y, int epi, float* y2, float* ws0, float* ws1, stream_t st,
               stream_t join_side = nullptr) {
    long cur[2] = {1, 1};
    for (int a = 0; a < nmid; ++a) cur[a] = mids[a].n_in;
    const float* src = in;
    for (int a = 0; a < nmid; ++a) {
        long O = P, I = m;
        for (int b = 0; b < a; ++b) O *= cur[b];
        for (int b = a + 1; b < nmid; ++b) I *= cur[b];
        float* d2 = (src == ws0) ? ws1 : ws0;
        MidArgs ma;
        ma.X = src; ma.Mat = mids[a].mat; ma.Y = d2;
        ma.O = (int)O; ma.H = mids[a].n_in; ma.J = mids[a].n_out; ma.I = (int)I;
        BE_TRY(be_mid(ma, st));
        cur[a] = mids[a].n_out;
        src = d2;
    }
    long R = P;
    for (int a = 0; a < nmid; ++a) R *= cur[a];
    GemmArgs g;
    g.A = src; g.a_rs = 2 * m; g.a_cs = 1;
    g.B = last_mat; g.ldb = n_last;
    g.C = y; g.ldc = n_last; g.C2 = y2;
    g.M = (int)R; g.N = n_last; g.K = 2 * m;
    g.epi = epi;
    g.tag = "dft_last_synthesis";
    g.b_const = true;
    if (join_side) BE_TRY(be_join(st, join_side));
    BE_TRY(be_gemm(g, st));
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// SpectralConv plan
// ---------------------------------------------------------------------------------------------------
struct SpectralPlan {
    int d = 0, nmid = 0;
    int in[3], out[3], m[3];
    DevMat a_last, s_last, ga_last, gs_last;
    DevMat a_mid[2], s_mid[2], ga_mid[2], gs_mid[2];
    MidOp fa[2], fs[2], ba[2], bs[2];   // fwd analyse / fwd synth / bwd analyse / bwd synth
    long Q = 0;                         // kept complex modes per (b,c) plane
    int build() {
        nmid = d - 1;
        const int nl_in = in[d - 1], nl_out = out[d - 1], ml = m[d - 1];
        double n_total = 1.0;
        for (int a = 0; a < d; ++a) n_total *= in[a];
        auto al = dft_last_analysis(nl_in, ml, 1.0 / n_total);
        auto sl = dft_last_synthesis(nl_out, ml, 1.0, true);
        BE_TRY(a_last.upload(al));
        BE_TRY(s_last.upload(sl));
        BE_TRY(ga_last.upload(transpose_real(sl, 2 * ml, nl_out)));
        BE_TRY(gs_last.upload(transpose_real(al, nl_in, 2 * ml)));
        Q = ml;
        for (int a = 0; a < nmid; ++a) {
            const int J = 2 * m[a];
            auto am = dft_mid_analysis(in[a], m[a]);
            auto sm = dft_mid_synthesis(out[a], m[a]);
            BE_TRY(a_mid[a].upload(am));
            BE_TRY(s_mid[a].upload(sm));
            BE_TRY(ga_mid[a].upload(conj_transpose(sm, out[a], J)));
            BE_TRY(gs_mid[a].upload(conj_transpose(am, J, in[a])));
            fa[a] = {a_mid[a].d, in[a], J};
            fs[a] = {s_mid[a].d, J, out[a]};
            ba[a] = {ga_mid[a].d, out[a], J};
            bs[a] = {gs_mid[a].d, J, in[a]};
            Q *= J;
        }
        return 0;
    }
};

struct Key {
    int v[12];
    bool operator<(const Key& o) const { return memcmp(v, o.v, sizeof v) < 0; }
};

std::mutex g_mu;
std::map<Key, std::unique_ptr<SpectralPlan>> g_spectral;

int check_conv_desc(const uno_conv_desc* d, bool need_modes) {
    if (!d) return fail(UNO_EINVAL, "null descriptor");
    if (d->ndim < 1 || d->ndim > 3) return fail(UNO_EINVAL, "ndim must be 1, 2 or 3 (got %d)", d->ndim);
    if (d->batch < 1 || d->in_ch < 1 || d->out_ch < 1)
        return fail(UNO_EINVAL, "batch/in_ch/out_ch must be positive (got %d/%d/%d)", d->batch, d->in_ch, d->out_ch);
    for (int a = 0; a < d->ndim; ++a) {
        if (d->in_dim[a] < 1 || d->out_dim[a] < 1)
            return fail(UNO_EINVAL, "grid sizes must be positive (axis %d: in %d out %d)", a, d->in_dim[a], d->out_dim[a]);
        if (!need_modes) continue;
        const bool last = a == d->ndim - 1;
        const int lim_in = last ? d->in_dim[a] / 2 + 1 : d->in_dim[a];
        const int lim_out = last ? d->out_dim[a] / 2 + 1 : d->out_dim[a];
        if (d->modes[a] < 1) return fail(UNO_EINVAL, "modes%d must be positive (got %d)", a + 1, d->modes[a]);
        // same failures the reference raises from einsum / slice-assign
        if (d->modes[a] > lim_in)
            return fail(UNO_EINVAL, "modes%d=%d exceeds the input spectrum size %d along axis %d", a + 1, d->modes[a], lim_in, a);
        if (d->modes[a] > lim_out)
            return fail(UNO_EINVAL, "modes%d=%d exceeds the output spectrum size %d along axis %d", a + 1, d->modes[a], lim_out, a);
    }
    return 0;
}

int get_spectral_plan(const uno_conv_desc* d, SpectralPlan** out) {
    Key k;
    memset(&k, 0, sizeof k);
    k.v[0] = d->ndim;
    for (int a = 0; a < d->ndim; ++a) { k.v[1 + a] = d->in_dim[a]; k.v[4 + a] = d->out_dim[a]; k.v[7 + a] = d->modes[a]; }
    k.v[11] = be_current_device();   // the constants live in that device's memory
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_spectral.find(k);
    if (it == g_spectral.end()) {
        std::unique_ptr<SpectralPlan> p(new SpectralPlan());
        p->d = d->ndim;
        for (int a = 0; a < 3; ++a) { p->in[a] = d->in_dim[a]; p->out[a] = d->out_dim[a]; p->m[a] = d->modes[a]; }
        int rc = p->build();
        if (rc) return fail(UNO_ECUDA, "failed to upload plan constants (%s)", g_err.c_str());
        it = g_spectral.emplace(k, std::move(p)).first;
    }
    *out = it->second.get();
    return 0;
}

struct SpectralSizes { size_t S, xh, yh; };   // floats

SpectralSizes spectral_sizes(const uno_conv_desc* d) {
    const int nmid = d->ndim - 1;
    const int ml = d->modes[d->ndim - 1];
    MidOp fa[2], fs[2], ba[2], bs[2];
    long Q = ml;
    for (int a = 0; a < nmid; ++a) {
        const int J = 2 * d->modes[a];
        fa[a] = {nullptr, d->in_dim[a], J};
        fs[a] = {nullptr, J, d->out_dim[a]};
        ba[a] = {nullptr, d->out_dim[a], J};
        bs[a] = {nullptr, J, d->in_dim[a]};
        Q *= J;
    }
    const long Pin = (long)d->batch * d->in_ch, Pout = (long)d->batch * d->out_ch;
    size_t S = analyse_stage_floats(Pin, nmid, fa, ml);
    S = std::max(S, synth_stage_floats(Pout, nmid, fs, ml));
    S = std::max(S, analyse_stage_floats(Pout, nmid, ba, ml));
    S = std::max(S, synth_stage_floats(Pin, nmid, bs, ml));
    return {align64(S), align64((size_t)2 * Pin * Q), align64((size_t)2 * Pout * Q)};
}

size_t spectral_ws_floats(const uno_conv_desc* d) {
    SpectralSizes s = spectral_sizes(d);
    return 2 * s.S + 2 * s.xh + 2 * s.yh + 256;
}

// corner bookkeeping: corner c = sum_a h_a 2^a  (h_a = 1 -> high block of axis a) -> weightsN, N = c+1
struct Corner { long off_x; };

void corner_geometry(const SpectralPlan* p, int c, long* off, int* q_outer, int* q_inner, long* sq_x, long* sq_w) {
    const int d = p->d, nmid = d - 1, ml = p->m[d - 1];
    long stride[2] = {0, 0};
    // plane layout [J_0, J_1, m]
    if (nmid >= 1) stride[nmid - 1] = ml;
    if (nmid == 2) stride[0] = (long)2 * p->m[1] * ml;
    long o = 0;
    for (int a = 0; a < nmid; ++a)
        if ((c >> a) & 1) o += (long)p->m[a] * stride[a];
    *off = o;
    if (nmid <= 1) {
        *q_outer = 1;
        *q_inner = (nmid == 1 ? p->m[0] : 1) * ml;
        *sq_x = 0; *sq_w = 0;
    } else {
        *q_outer = p->m[0];
        *q_inner = p->m[1] * ml;
        *sq_x = stride[0];
        *sq_w = (long)p->m[1] * ml;
    }
}

// Profiling scope of one fused spectral-convolution call (bench.py's per-U-level roofline): algorithmic bytes and
// contraction flops as SURVEY.md 8(d) defines them.  `extra_out_tensors`: output-sized tensors the fused block epilogue
// reads / writes on top of the plain layer (forward: the pointwise sum it accumulates onto and, with GELU to a second
// tensor, that tensor; backward: the input gradient it accumulates onto, when it does).
struct SpectralProfScope {
    bool on = false, range = false;
    SpectralProfScope(const uno_conv_desc* d, bool backward, int extra_out_tensors) {
        range = cfg(CFG_NVTX) != 0;
        if (!be_profile_enabled() && !range) return;
        const int nd = d->ndim;
        double n_in = 1, n_out = 1, M = 1;
        for (int a = 0; a < nd; ++a) { n_in *= d->in_dim[a]; n_out *= d->out_dim[a]; M *= d->modes[a]; }
        const double nW = (double)(1 << (nd - 1)), B = d->batch, Ci = d->in_ch, Co = d->out_ch;
        double bytes = 4.0 * B * (Ci * n_in + Co * n_out);
        if (!backward) bytes += 8.0 * nW * Ci * Co * M;
        else bytes += 8.0 * B * Ci * nW * M + 16.0 * nW * Ci * Co * M;
        bytes += 4.0 * B * (backward ? Ci * n_in : Co * n_out) * extra_out_tensors;
        const double flops = 8.0 * B * Ci * Co * nW * M * (backward ? 2.0 : 1.0);
        char label[256];
        int o = snprintf(label, sizeof label, "spectral %s B=%d %d->%d [", backward ? "bwd" : "fwd", d->batch, d->in_ch, d->out_ch);
        for (int a = 0; a < nd; ++a) o += snprintf(label + o, sizeof label - o, "%s%d", a ? "," : "", d->in_dim[a]);
        o += snprintf(label + o, sizeof label - o, "]->[");
        for (int a = 0; a < nd; ++a) o += snprintf(label + o, sizeof label - o, "%s%d", a ? "," : "", d->out_dim[a]);
        o += snprintf(label + o, sizeof label - o, "] modes=[");
        for (int a = 0; a < nd; ++a) o += snprintf(label + o, sizeof label - o, "%s%d", a ? "," : "", d->modes[a]);
        snprintf(label + o, sizeof label - o, "]");
        if (range) be_range_push(label);          // one NVTX range per fused spectral convolution (U-level and direction)
        if (be_profile_enabled()) {
            be_profile_scope_begin(label, bytes, flops);
            on = true;
        }
    }
    ~SpectralProfScope() {
        if (on) be_profile_scope_end();
        if (range) be_range_pop();
    }
};

// NVTX range around one C-ABI call (switch `nvtx`)
struct ApiRange {
    bool on;
    explicit ApiRange(const char* name) : on(cfg(CFG_NVTX) != 0) { if (on) be_range_push(name); }
    ~ApiRange() { if (on) be_range_pop(); }
};

int spectral_fwd_impl(const uno_conv_desc* d, SpectralPlan* p, const float* x, const float* const* w,
                      float* y, int epi, float* y2, float* xhat, Arena& ar, stream_t st, stream_t join_side = nullptr) {
    const int nd = d->ndim, nmid = nd - 1, ml = d->modes[nd - 1];
    const long Pin = (long)d->batch * d->in_ch, Pout = (long)d->batch * d->out_ch;
    SpectralSizes sz = spectral_sizes(d);
    float* ws0 = ar.take(sz.S);
    float* ws1 = ar.take(sz.S);
    float* yhat = ar.take(sz.yh);
    if (!xhat) xhat = ar.take(sz.xh);
    if (!ar.ok) return fail(UNO_EWORKSPACE, "workspace too small for spectral conv forward");
    SpectralProfScope prof_scope(d, false, epi == EPI_STORE ? 0 : (epi == EPI_ACCUM_GELU ? 2 : 1));
    UNO_TRY(analyse(x, Pin, nmid, p->fa, d->in_dim[nd - 1], p->a_last.d, ml, xhat, ws0, ws1, st));
    const long Q = p->Q;
    const int ncorner = 1 << nmid;
    long Qw = 1;
    for (int a = 0; a < nd; ++a) Qw *= d->modes[a];
    {   // all corners (weights1..) in one launch: same shapes and strides, different offsets into the kept-mode planes
        CmmArgs ca;
        ca.ncorner = ncorner;
        long off = 0, sqx = 0, sqw = 0; int qo = 1, qi = 0;
        for (int c = 0; c < ncorner; ++c) {
            corner_geometry(p, c, &off, &qo, &qi, &sqx, &sqw);
            ca.A[c] = xhat + 2 * off;
            ca.B[c] = w[c];
            ca.C[c] = yhat + 2 * off;
        }
        ca.a_sm = (long)d->in_ch * Q; ca.a_sk = Q; ca.a_sqo = sqx;
        ca.b_sk = (long)d->out_ch * Qw; ca.b_sn = Qw; ca.b_sqo = sqw;
        ca.c_sm = (long)d->out_ch * Q; ca.c_sn = Q; ca.c_sqo = sqx;
        ca.M = d->batch; ca.N = d->out_ch; ca.K = d->in_ch; ca.q_outer = qo; ca.q_inner = qi;
        ca.deterministic = 1;            // forward: repeated calls give bit-identical outputs
        BE_TRY(be_cmm(ca, st));
    }
    UNO_TRY(synthesise(yhat, Pout, nmid, p->fs, ml, p->s_last.d, d->out_dim[nd - 1], y, epi, y2, ws0, ws1, st, join_side));
    return 0;
}

// `wst`: stream for the weight-gradient contraction dW (independent of the dX chain once ghat exists); the caller joins it.
int spectral_bwd_impl(const uno_conv_desc* d, SpectralPlan* p, const float* gy, const float* xhat,
                      const float* const* w, float* gx, float* const* gw, int accumulate_gx, Arena& ar,
                      stream_t st, stream_t join_side = nullptr, stream_t wst = nullptr) {
    const int nd = d->ndim, nmid = nd - 1, ml = d->modes[nd - 1];
    const long Pin = (long)d->batch * d->in_ch, Pout = (long)d->batch * d->out_ch;
    SpectralSizes sz = spectral_sizes(d);
    float* ws0 = ar.take(sz.S);
    float* ws1 = ar.take(sz.S);
    float* ghat = ar.take(sz.yh);
    float* dxhat = gx ? ar.take(sz.xh) : nullptr;
    if (!ar.ok) return fail(UNO_EWORKSPACE, "workspace too small for spectral conv backward");
    SpectralProfScope prof_scope(d, true, (gx && accumulate_gx) ? 1 : 0);
    UNO_TRY(analyse(gy, Pout, nmid, p->ba, d->out_dim[nd - 1], p->ga_last.d, ml, ghat, ws0, ws1, st));
    const long Q = p->Q;
    long Qw = 1;
    for (int a = 0; a < nd; ++a) Qw *= d->modes[a];
    const int ncorner = 1 << nmid;
    long off[4], sqx = 0, sqw = 0; int qo = 1, qi = 0;
    for (int c = 0; c < ncorner; ++c) corner_geometry(p, c, &off[c], &qo, &qi, &sqx, &sqw);
    if (gw) {   // dW[i,o,q] = sum_b conj(xhat[b,i,q]) ghat[b,o,q], corners with a gradient buffer, one launch
        CmmArgs ca;
        int n = 0;
        for (int c = 0; c < ncorner; ++c) {
            if (!gw[c]) continue;
            ca.A[n] = xhat + 2 * off[c];
            ca.B[n] = ghat + 2 * off[c];
            ca.C[n] = gw[c];
            ++n;
        }
        ca.ncorner = n;
        ca.a_sm = Q; ca.a_sk = (long)d->in_ch * Q; ca.a_sqo = sqx; ca.conjA = 1;
        ca.b_sk = (long)d->out_ch * Q; ca.b_sn = Q; ca.b_sqo = sqx;
        ca.c_sm = (long)d->out_ch * Qw; ca.c_sn = Qw; ca.c_sqo = sqw;
        ca.M = d->in_ch; ca.N = d->out_ch; ca.K = d->batch; ca.q_outer = qo; ca.q_inner = qi;
        if (n > 0) {
            if (wst && wst != st) {
                BE_TRY(be_fork(st, wst));
                BE_TRY(be_cmm(ca, wst));
            } else {
                BE_TRY(be_cmm(ca, st));
            }
        }
    }
    if (gx) {   // dxhat[b,i,q] = sum_o ghat[b,o,q] conj(w[i,o,q])
        CmmArgs ca;
        ca.ncorner = ncorner;
        for (int c = 0; c < ncorner; ++c) {
            ca.A[c] = ghat + 2 * off[c];
            ca.B[c] = w[c];
            ca.C[c] = dxhat + 2 * off[c];
        }
        ca.a_sm = (long)d->out_ch * Q; ca.a_sk = Q; ca.a_sqo = sqx;
        ca.b_sk = Qw; ca.b_sn = (long)d->out_ch * Qw; ca.b_sqo = sqw; ca.conjB = 1;
        ca.c_sm = (long)d->in_ch * Q; ca.c_sn = Q; ca.c_sqo = sqx;
        ca.M = d->batch; ca.N = d->in_ch; ca.K = d->out_ch; ca.q_outer = qo; ca.q_inner = qi;
        BE_TRY(be_cmm(ca, st));
    }
    if (gx)
        UNO_TRY(synthesise(dxhat, Pin, nmid, p->bs, ml, p->gs_last.d, d->in_dim[nd - 1], gx,
                           accumulate_gx ? EPI_ACCUM : EPI_STORE, nullptr, ws0, ws1, st, join_side));
    else if (join_side)
        BE_TRY(be_join(st, join_side));
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// pointwise branch: channel mix (1x1 conv) + spatial resample R
//   2-D: R = anti-aliased bicubic, separable banded            (integral_operators.py:240-242)
//   3-D: R = rfftn / corner copy / irfftn(s=out) "spectral resample" (integral_operators.py:448-467)
// R is linear and acts per channel, so it commutes with the channel mix; it is applied on whichever
// side has fewer channels.  R(1) = gain (1 for bicubic; N_in/N_out for the 3-D operator).
// ---------------------------------------------------------------------------------------------------
// Opt-in, NOT the reference's behaviour: the switch pointwise3d_fixed replaces pointwise_op_3D's quirky
// "spectral resample" by the band-limited Fourier resample it approximates (plan.h sr_mid_fixed).  Read when a call looks its
// plan up, so set it before the first forward and keep it for the matching backward.
inline bool pointwise3d_fixed() { return cfg(CFG_POINTWISE3D_FIXED) != 0; }

struct ResamplePlan {
    int d = 0;
    bool fixed3d = false;
    int in[3], out[3];
    bool identity = false;
    double gain = 1.0;
    long n_in = 1, n_out = 1;
    // 2-D
    DevBand r[2], rt[2];
    // 3-D
    int m3 = 0;
    DevMat sa_last, ss_last, sga_last, sgs_last;   // analysis [I3 x 2m3], synthesis [2m3 x d3], and transposes
    DevMat l_mid[2], lh_mid[2];                    // [d_a x I_a] and conj transposes [I_a x d_a]

    int build() {
        n_in = n_out = 1;
        for (int a = 0; a < d; ++a) { n_in *= in[a]; n_out *= out[a]; }
        if (d == 2) {
            identity = in[0] == out[0] && in[1] == out[1];
            gain = 1.0;
            if (identity) return 0;
            for (int a = 0; a < 2; ++a) {
                Banded b = bicubic_aa(in[a], out[a]);
                BE_TRY(r[a].upload(b));
                BE_TRY(rt[a].upload(banded_transpose(b)));
            }
            return 0;
        }
        // 3-D spectral resample (never the identity: even same-size drops the top half-axis bins)
        identity = false;
        m3 = fixed3d ? sr_last_modes_fixed(in[2], out[2]) : sr_last_modes(in[2], out[2]);
        gain = fixed3d ? 1.0 : (m3 >= 1 ? (double)n_in / (double)n_out : 0.0);
        if (m3 == 0) return 0;
        auto al = dft_last_analysis(in[2], m3, 1.0);
        auto sl = dft_last_synthesis(out[2], m3, 1.0 / (double)(fixed3d ? n_in : n_out), true);
        BE_TRY(sa_last.upload(al));
        BE_TRY(ss_last.upload(sl));
        BE_TRY(sga_last.upload(transpose_real(sl, 2 * m3, out[2])));
        BE_TRY(sgs_last.upload(transpose_real(al, in[2], 2 * m3)));
        for (int a = 0; a < 2; ++a) {
            auto L = fixed3d ? sr_mid_fixed(in[a], out[a]) : sr_mid(in[a], out[a]);
            BE_TRY(l_mid[a].upload(L));
            BE_TRY(lh_mid[a].upload(conj_transpose(L, out[a], in[a])));
        }
        return 0;
    }

    // scratch floats needed by apply()/applyT() on P planes
    size_t ws_floats(long P) const {
        if (identity) return 0;
        if (d == 2) return align64((size_t)P * std::max(in[0], out[0]) * std::max(in[1], out[1]));
        if (m3 == 0) return 0;
        MidOp f[2] = {{nullptr, in[0], out[0]}, {nullptr, in[1], out[1]}};
        MidOp b[2] = {{nullptr, out[0], in[0]}, {nullptr, out[1], in[1]}};
        size_t S = std::max(analyse_stage_floats(P, 2, f, m3), analyse_stage_floats(P, 2, b, m3));
        return 3 * align64(S);
    }

    int banded2(const float* x, long P, float* y, float* tmp, const DevBand* b, stream_t st) const {
        // b[0] acts on axis 0 (rows), b[1] on axis 1 (last); one fused, shared-memory tiled kernel
        Banded2DArgs a;
        a.x = x; a.y = y; a.planes = P; a.tmp = tmp;
        a.start0 = b[0].start; a.w0 = b[0].w; a.n_in0 = b[0].n_in; a.n_out0 = b[0].n_out; a.taps0 = b[0].taps; a.span0 = b[0].span32;
        a.start1 = b[1].start; a.w1 = b[1].w; a.n_in1 = b[1].n_in; a.n_out1 = b[1].n_out; a.taps1 = b[1].taps; a.span1 = b[1].span64;
        if (b[0].G && b[1].G) {
            a.gs0 = b[0].gstart; a.D0 = b[0].D; a.G0 = b[0].G; a.W0 = b[0].W; a.ng0 = b[0].ng;
            a.tile_groups0 = b[0].tile_groups_rows; a.tile_span0 = b[0].tile_span_rows;
            a.gs1 = b[1].gstart; a.D1 = b[1].D; a.G1 = b[1].G; a.W1 = b[1].W; a.ng1 = b[1].ng; a.tile_span1 = b[1].tile_span_cols;
        }
        BE_TRY(be_banded2d(a, st));
        return 0;
    }

    // y [P, *out] = R x [P, *in]
    int apply(const float* x, long P, float* y, float* ws, stream_t st) const {
        if (d == 2) return banded2(x, P, y, ws, r, st);
        if (m3 == 0) { BE_TRY(be_memset(y, 0, (size_t)P * n_out * sizeof(float), st)); return 0; }
        MidOp f[2] = {{l_mid[0].d, in[0], out[0]}, {l_mid[1].d, in[1], out[1]}};
        size_t S = ws_floats(P) / 3;
        float *w0 = ws, *w1 = ws + S, *mid = ws + 2 * S;
        UNO_TRY(analyse(x, P, 2, f, in[2], sa_last.d, m3, mid, w0, w1, st));
        return synthesise(mid, P * out[0] * out[1], 0, nullptr, m3, ss_last.d, out[2], y, EPI_STORE, nullptr, w0, w1, st);
    }

    // gx [P, *in] = R^T g [P, *out]
    int applyT(const float* g, long P, float* gx, float* ws, stream_t st) const {
        if (d == 2) return banded2(g, P, gx, ws, rt, st);
        if (m3 == 0) { BE_TRY(be_memset(gx, 0, (size_t)P * n_in * sizeof(float), st)); return 0; }
        MidOp b[2] = {{lh_mid[0].d, out[0], in[0]}, {lh_mid[1].d, out[1], in[1]}};
        size_t S = ws_floats(P) / 3;
        float *w0 = ws, *w1 = ws + S, *mid = ws + 2 * S;
        UNO_TRY(analyse(g, P, 2, b, out[2], sga_last.d, m3, mid, w0, w1, st));
        return synthesise(mid, P * in[0] * in[1], 0, nullptr, m3, sgs_last.d, in[2], gx, EPI_STORE, nullptr, w0, w1, st);
    }
};

std::map<Key, std::unique_ptr<ResamplePlan>> g_resample;

int get_resample_plan(const uno_conv_desc* d, ResamplePlan** out) {
    if (d->ndim != 2 && d->ndim != 3)
        return fail(UNO_EINVAL, "pointwise_op_1D is unsupported: the reference itself raises ValueError "
                                "(linear + antialias) on torch >= 1.11");
    Key k;
    memset(&k, 0, sizeof k);
    k.v[0] = d->ndim;
    for (int a = 0; a < d->ndim; ++a) { k.v[1 + a] = d->in_dim[a]; k.v[4 + a] = d->out_dim[a]; }
    const bool fixed3d = d->ndim == 3 && pointwise3d_fixed();
    k.v[7] = fixed3d ? 1 : 0;
    k.v[11] = be_current_device();
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_resample.find(k);
    if (it == g_resample.end()) {
        std::unique_ptr<ResamplePlan> p(new ResamplePlan());
        p->d = d->ndim;
        p->fixed3d = fixed3d;
        for (int a = 0; a < 3; ++a) { p->in[a] = d->in_dim[a]; p->out[a] = d->out_dim[a]; }
        int rc = p->build();
        if (rc) return fail(UNO_ECUDA, "failed to upload resample plan (%s)", g_err.c_str());
        it = g_resample.emplace(k, std::move(p)).first;
    }
    *out = it->second.get();
    return 0;
}

inline bool resample_first(const uno_conv_desc* d) { return d->in_ch <= d->out_ch; }

struct PwGeom { long n_in, n_out; bool identity; };

PwGeom pw_geom(const uno_conv_desc* d) {
    PwGeom g{1, 1, false};
    for (int a = 0; a < d->ndim; ++a) { g.n_in *= d->in_dim[a]; g.n_out *= d->out_dim[a]; }
    g.identity = d->ndim == 2 && d->in_dim[0] == d->out_dim[0] && d->in_dim[1] == d->out_dim[1];
    return g;
}

size_t pw_resample_ws(const uno_conv_desc* d, long P) {
    // mirrors ResamplePlan::ws_floats without needing device uploads
    ResamplePlan tmp;
    tmp.d = d->ndim;
    for (int a = 0; a < 3; ++a) { tmp.in[a] = d->in_dim[a]; tmp.out[a] = d->out_dim[a]; }
    tmp.identity = pw_geom(d).identity;
    // the larger of the two modes' kept-bin counts: the workspace query must not depend on an environment switch
    tmp.m3 = d->ndim == 3 ? std::max(sr_last_modes(d->in_dim[2], d->out_dim[2]), sr_last_modes_fixed(d->in_dim[2], d->out_dim[2])) : 0;
    return tmp.ws_floats(P);
}

size_t pw_ws_floats(const uno_conv_desc* d) {
    PwGeom g = pw_geom(d);
    if (g.identity) return 256;
    if (resample_first(d)) {
        const long P = (long)d->batch * d->in_ch;
        return 2 * align64((size_t)P * g.n_out) + pw_resample_ws(d, P) + 256;   // r (recomputed) + gr + scratch
    }
    const long P = (long)d->batch * d->out_ch;
    return align64((size_t)P * g.n_in) + pw_resample_ws(d, P) + 256;            // t / gt + scratch
}

int conv1x1(const float* wmat, long w_rs, long w_cs, const float* bias, const float* x, float* z,
            int Co, int Ci, long npix, int batch, int epi, stream_t st) {
    GemmArgs g;
    g.A = wmat; g.a_rs = w_rs; g.a_cs = w_cs; g.sA = 0;
    g.B = x; g.ldb = npix; g.sB = (long)Ci * npix;
    g.C = z; g.ldc = npix; g.sC = (long)Co * npix;
    g.bias = bias;
    g.M = Co; g.N = (int)npix; g.K = Ci; g.batch = batch; g.epi = epi;
    g.tag = "conv1x1";
    g.channel_mix = true;
    return be_gemm(g, st);
}

int pointwise_fwd_impl(const uno_conv_desc* d, ResamplePlan* rp, const float* x, const float* conv_w,
                       const float* conv_b, float* z, float* saved, Arena& ar, stream_t st) {
    PwGeom g = pw_geom(d);
    const int B = d->batch, Ci = d->in_ch, Co = d->out_ch;
    if (g.identity) { BE_TRY(conv1x1(conv_w, Ci, 1, conv_b, x, z, Co, Ci, g.n_in, B, EPI_STORE, st)); return 0; }
    if (resample_first(d)) {
        const long P = (long)B * Ci;
        float* r = saved ? saved : ar.take((size_t)P * g.n_out);
        float* scratch = ar.take(rp->ws_floats(P));
        if (!ar.ok) return fail(UNO_EWORKSPACE, "workspace too small for pointwise forward");
        UNO_TRY(rp->apply(x, P, r, scratch, st));
        if (rp->gain == 1.0) { BE_TRY(conv1x1(conv_w, Ci, 1, conv_b, r, z, Co, Ci, g.n_out, B, EPI_STORE, st)); return 0; }
        BE_TRY(conv1x1(conv_w, Ci, 1, nullptr, r, z, Co, Ci, g.n_out, B, EPI_STORE, st));
        BE_TRY(be_add_channel_const(z, conv_b, (float)rp->gain, (long)B * Co, Co, g.n_out, st));
        return 0;
    }
    const long P = (long)B * Co;
    float* t = ar.take((size_t)P * g.n_in);
    float* scratch = ar.take(rp->ws_floats(P));
    if (!ar.ok) return fail(UNO_EWORKSPACE, "workspace too small for pointwise forward");
    BE_TRY(conv1x1(conv_w, Ci, 1, conv_b, x, t, Co, Ci, g.n_in, B, EPI_STORE, st));
    return rp->apply(t, P, z, scratch, st);
}

// gconv_b == NULL: the caller has already produced the bias gradient (fused into the activation backward)
// `wst`: stream for the weight-gradient reduction (reads gz and the saved / recomputed resampled input, writes gconv_w); the
// caller forks it from `st` before this call and joins it afterwards.  NULL: everything on `st`.
int pointwise_bwd_impl(const uno_conv_desc* d, ResamplePlan* rp, const float* gz, const float* x,
                       const float* saved, const float* conv_w, float* gx, float* gconv_w,
                       float* gconv_b, Arena& ar, stream_t st, stream_t wst = nullptr) {
    if (!wst || (!saved && resample_first(d) && !pw_geom(d).identity)) wst = st;   // (a recomputed resample feeds the reduction: keep one stream)
    if (wst != st) BE_TRY(be_fork(st, wst));
    PwGeom g = pw_geom(d);
    const int B = d->batch, Ci = d->in_ch, Co = d->out_ch;
    const float gain = g.identity ? 1.0f : (float)rp->gain;
    if (gconv_b) {
        BE_TRY(be_memset(gconv_b, 0, Co * sizeof(float), st));
        BE_TRY(be_channel_sum(gz, gconv_b, (long)B * Co, Co, g.n_out, gain, st));
    }
    if (gconv_w) BE_TRY(be_memset(gconv_w, 0, (size_t)Co * Ci * sizeof(float), wst));
    auto wgrad = [&](const float* gmat, const float* act, long npix) {
        GemmNtArgs a;
        a.A = gmat; a.lda = npix; a.sA = (long)Co * npix;
        a.B = act; a.ldb = npix; a.sB = (long)Ci * npix;
        a.C = gconv_w; a.ldc = Ci;
        a.M = Co; a.N = Ci; a.K = (int)npix; a.batch = B;
        if (wst != st) {            // the reduction's inputs were produced on `st`
            const int rc = be_fork(st, wst);
            if (rc) return rc;
        }
        return be_gemm_nt_atomic(a, wst);
    };
    if (g.identity) {
        if (gconv_w) BE_TRY(wgrad(gz, x, g.n_in));
        if (gx) BE_TRY(conv1x1(conv_w, 1, Ci, nullptr, gz, gx, Ci, Co, g.n_in, B, EPI_STORE, st));
        return 0;
    }
    if (resample_first(d)) {
        const long P = (long)B * Ci;
        float* gr = ar.take((size_t)P * g.n_out);
        float* rbuf = saved ? nullptr : ar.take((size_t)P * g.n_out);
        float* scratch = ar.take(rp->ws_floats(P));
        if (!ar.ok) return fail(UNO_EWORKSPACE, "workspace too small for pointwise backward");
        const float* r = saved;
        if (gconv_w) {
            if (!r) { UNO_TRY(rp->apply(x, P, rbuf, scratch, st)); r = rbuf; }
            BE_TRY(wgrad(gz, r, g.n_out));
        }
        if (gx) {
            BE_TRY(conv1x1(conv_w, 1, Ci, nullptr, gz, gr, Ci, Co, g.n_out, B, EPI_STORE, st));
            UNO_TRY(rp->applyT(gr, P, gx, scratch, st));
        }
        return 0;
    }
    const long P = (long)B * Co;
    float* gt = ar.take((size_t)P * g.n_in);
    float* scratch = ar.take(rp->ws_floats(P));
    if (!ar.ok) return fail(UNO_EWORKSPACE, "workspace too small for pointwise backward");
    UNO_TRY(rp->applyT(gz, P, gt, scratch, st));
    if (gconv_w) BE_TRY(wgrad(gt, x, g.n_in));
    if (gx) BE_TRY(conv1x1(conv_w, 1, Ci, nullptr, gt, gx, Ci, Co, g.n_in, B, EPI_STORE, st));
    return 0;
}

size_t block_ws_floats(const uno_block_desc* bd) {
    const uno_conv_desc* d = &bd->conv;
    PwGeom g = pw_geom(d);
    const size_t act = align64((size_t)d->batch * d->out_ch * g.n_out);
    return spectral_ws_floats(d) + pw_ws_floats(d) + act + align64((size_t)2 * d->batch * d->out_ch) + 256;
}

// ---- model glue: lift / project ------------------------------------------------------------------------
int check_pixel_desc(const uno_pixel_desc* p) {
    if (p->ndim != 2 && p->ndim != 3) return fail(UNO_EINVAL, "pixel desc: ndim must be 2 or 3 (got %d)", p->ndim);
    if (p->batch < 1) return fail(UNO_EINVAL, "pixel desc: batch must be positive (got %d)", p->batch);
    for (int a = 0; a < p->ndim; ++a)
        if (p->dim[a] < 1 || p->pad_lo[a] < 0 || p->pad_hi[a] < 0)
            return fail(UNO_EINVAL, "pixel desc: axis %d has dim %d pad (%d,%d)", a, p->dim[a], p->pad_lo[a], p->pad_hi[a]);
    return 0;
}

// spatial axes right-aligned into three slots (2-D: slot 0 has extent 1)
void pixel_geometry(const uno_pixel_desc* p, int* n, int* N, int* lo) {
    const int shift = 3 - p->ndim;
    for (int a = 0; a < 3; ++a) { n[a] = N[a] = 1; lo[a] = 0; }
    for (int a = 0; a < p->ndim; ++a) {
        n[a + shift] = p->dim[a];
        N[a + shift] = p->dim[a] + p->pad_lo[a] + p->pad_hi[a];
        lo[a + shift] = p->pad_lo[a];
    }
}

int make_lift_args(const uno_lift_desc* d, LiftArgs* a) {
    if (!d) return fail(UNO_EINVAL, "null descriptor");
    UNO_TRY(check_pixel_desc(&d->px));
    a->batch = d->px.batch;
    pixel_geometry(&d->px, a->n, a->N, a->lo);
    a->raw_ch = d->raw_ch; a->grid_ch = d->grid_ch; a->hid = d->hidden; a->out_ch = d->out_ch;
    if (d->raw_ch < 1 || d->grid_ch < 0 || d->hidden < 1 || d->out_ch < 1)
        return fail(UNO_EINVAL, "lift: raw_ch/grid_ch/hidden/out_ch = %d/%d/%d/%d", d->raw_ch, d->grid_ch, d->hidden, d->out_ch);
    if (!be_lift_supported(*a))
        return fail(UNO_EINVAL, "lift: unsupported widths (raw_ch+grid_ch=%d <= 16, hidden=%d <= 32, out_ch=%d <= 64 required)",
                    d->raw_ch + d->grid_ch, d->hidden, d->out_ch);
    return 0;
}

int make_proj_args(const uno_project_desc* d, ProjArgs* a) {
    if (!d) return fail(UNO_EINVAL, "null descriptor");
    UNO_TRY(check_pixel_desc(&d->px));
    a->batch = d->px.batch;
    pixel_geometry(&d->px, a->n, a->N, a->lo);
    a->nsrc = d->nsrc; a->hid = d->hidden; a->out_ch = d->out_ch;
    if (d->nsrc < 1 || d->nsrc > 4) return fail(UNO_EINVAL, "project: nsrc must be 1..4 (got %d)", d->nsrc);
    int ctot = 0;
    for (int s = 0; s < d->nsrc; ++s) { a->src_ch[s] = d->src_ch[s]; ctot += d->src_ch[s]; }
    if (!be_proj_supported(*a))
        return fail(UNO_EINVAL, "project: unsupported widths (sum src_ch=%d <= 64, hidden=%d <= 128, out_ch=%d <= 4 required)",
                    ctot, d->hidden, d->out_ch);
    return 0;
}

}  // namespace

// =====================================================================================================
extern "C" {

const char* uno_last_error(void) { return g_err.c_str(); }
int uno_version(void) { return 100; }
const char* uno_backend_name(void) { return be_name(); }

void uno_clear_plans(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_spectral.clear();
    g_resample.clear();
}

int uno_spectral_conv_check(const uno_conv_desc* d) { return check_conv_desc(d, true); }

size_t uno_spectral_conv_workspace_bytes(const uno_conv_desc* d) {
    if (check_conv_desc(d, true)) return 0;
    return spectral_ws_floats(d) * sizeof(float);
}

size_t uno_spectral_conv_xhat_elems(const uno_conv_desc* d) {
    if (check_conv_desc(d, true)) return 0;
    size_t q = d->modes[d->ndim - 1];
    for (int a = 0; a < d->ndim - 1; ++a) q *= 2 * d->modes[a];
    return (size_t)d->batch * d->in_ch * q;
}

int uno_spectral_conv_fwd(const uno_conv_desc* d, const float* x, const float* const* w, float* y,
                          float* xhat, void* ws, size_t ws_bytes, void* stream) {
    ApiRange api_range("uno_spectral_conv_fwd");
    UNO_TRY(check_conv_desc(d, true));
    if (!x || !w || !y) return fail(UNO_EINVAL, "null tensor pointer");
    SpectralPlan* p;
    UNO_TRY(get_spectral_plan(d, &p));
    Arena ar(ws, ws_bytes);
    return spectral_fwd_impl(d, p, x, w, y, EPI_STORE, nullptr, xhat, ar, stream);
}

int uno_spectral_conv_bwd(const uno_conv_desc* d, const float* gy, const float* xhat,
                          const float* const* w, float* gx, float* const* gw, int accumulate_gx,
                          void* ws, size_t ws_bytes, void* stream) {
    ApiRange api_range("uno_spectral_conv_bwd");
    UNO_TRY(check_conv_desc(d, true));
    if (!gy || !xhat || !w) return fail(UNO_EINVAL, "null tensor pointer");
    SpectralPlan* p;
    UNO_TRY(get_spectral_plan(d, &p));
    Arena ar(ws, ws_bytes);
    return spectral_bwd_impl(d, p, gy, xhat, w, gx, gw, accumulate_gx, ar, stream);
}

size_t uno_pointwise_workspace_bytes(const uno_conv_desc* d) {
    if (check_conv_desc(d, false) || d->ndim < 2) return 0;
    return pw_ws_floats(d) * sizeof(float);
}

size_t uno_pointwise_saved_elems(const uno_conv_desc* d) {
    if (check_conv_desc(d, false) || d->ndim < 2) return 0;
    PwGeom g = pw_geom(d);
    if (g.identity || !resample_first(d)) return 0;
    return (size_t)d->batch * d->in_ch * g.n_out;
}

int uno_pointwise_fwd(const uno_conv_desc* d, const float* x, const float* conv_w,
                      const float* conv_b, float* z, float* saved, void* ws, size_t ws_bytes,
                      void* stream) {
    ApiRange api_range("uno_pointwise_fwd");
    UNO_TRY(check_conv_desc(d, false));
    if (!x || !conv_w || !conv_b || !z) return fail(UNO_EINVAL, "null tensor pointer");
    ResamplePlan* rp;
    UNO_TRY(get_resample_plan(d, &rp));
    Arena ar(ws, ws_bytes);
    return pointwise_fwd_impl(d, rp, x, conv_w, conv_b, z, saved, ar, stream);
}

int uno_pointwise_bwd(const uno_conv_desc* d, const float* gz, const float* x, const float* saved,
                      const float* conv_w, float* gx, float* gconv_w, float* gconv_b, void* ws,
                      size_t ws_bytes, void* stream) {
    ApiRange api_range("uno_pointwise_bwd");
    UNO_TRY(check_conv_desc(d, false));
    if (!gz || !x || !conv_w) return fail(UNO_EINVAL, "null tensor pointer");
    ResamplePlan* rp;
    UNO_TRY(get_resample_plan(d, &rp));
    Arena ar(ws, ws_bytes);
    return pointwise_bwd_impl(d, rp, gz, x, saved, conv_w, gx, gconv_w, gconv_b, ar, stream);
}

size_t uno_operator_block_workspace_bytes(const uno_block_desc* bd) {
    if (!bd || check_conv_desc(&bd->conv, true) || bd->conv.ndim < 2) return 0;
    return block_ws_floats(bd) * sizeof(float);
}

int uno_operator_block_fwd(const uno_block_desc* bd, const float* x, const float* const* w,
                           const float* conv_w, const float* conv_b, const float* gamma,
                           const float* beta, float* y, float* xhat, float* pw_saved, float* pre,
                           float* stats, void* ws, size_t ws_bytes, void* stream) {
    ApiRange api_range("uno_operator_block_fwd");
    if (!bd) return fail(UNO_EINVAL, "null descriptor");
    const uno_conv_desc* d = &bd->conv;
    UNO_TRY(check_conv_desc(d, true));
    if (!x || !w || !conv_w || !conv_b || !y) return fail(UNO_EINVAL, "null tensor pointer");
    if (bd->normalize && (!gamma || !beta)) return fail(UNO_EINVAL, "normalize requires gamma and beta");
    if (!bd->normalize && !bd->non_lin && pre) return fail(UNO_EINVAL, "pre must be NULL when neither normalize nor non_lin is set");
    SpectralPlan* sp;
    ResamplePlan* rp;
    UNO_TRY(get_spectral_plan(d, &sp));
    UNO_TRY(get_resample_plan(d, &rp));
    PwGeom g = pw_geom(d);
    const long planes = (long)d->batch * d->out_ch;
    Arena ar(ws, ws_bytes);
    // where the sum conv(x)+w(x) is assembled
    float* acc = pre;
    if (!acc) acc = bd->normalize ? ar.take((size_t)planes * g.n_out) : y;
    if (bd->normalize && !stats) stats = ar.take((size_t)2 * planes);
    if (!ar.ok) return fail(UNO_EWORKSPACE, "workspace too small for operator block forward");
    // The two branches are independent until the last synthesis stage adds the spectral part onto w(x): the pointwise
    // branch runs on the library's side stream, the spectral analysis / contraction / leading-axis synthesis on the
    // caller's; they use disjoint halves of the workspace and meet again before the fused epilogue.
    const size_t pw_floats = pw_ws_floats(d);
    if (ar.off + pw_floats > ar.cap) return fail(UNO_EWORKSPACE, "workspace too small for operator block forward");
    stream_t side = be_side_stream();
    if (side) BE_TRY(be_fork(stream, side));
    {
        Arena sub(ar.base + ar.off, pw_floats * sizeof(float));
        UNO_TRY(pointwise_fwd_impl(d, rp, x, conv_w, conv_b, acc, pw_saved, sub, side ? side : stream));
    }
    int epi = EPI_ACCUM;
    float* y2 = nullptr;
    if (!bd->normalize && bd->non_lin) {
        if (acc == y) epi = EPI_ACCUM_GELU_INPLACE;
        else { epi = EPI_ACCUM_GELU; y2 = y; }
    }
    {
        Arena sub(ar.base + ar.off + pw_floats, (ar.cap - ar.off - pw_floats) * sizeof(float));
        UNO_TRY(spectral_fwd_impl(d, sp, x, w, acc, epi, y2, xhat, sub, stream, side));
    }
    if (bd->normalize) {
        BE_TRY(be_norm_fused_fwd(acc, stats, gamma, beta, y, planes, d->out_ch, g.n_out, bd->eps, bd->non_lin, stream));
    }
    return 0;
}

int uno_operator_block_bwd(const uno_block_desc* bd, const float* gy, const float* x,
                           const float* xhat, const float* pw_saved, const float* pre,
                           const float* stats, const float* const* w, const float* conv_w,
                           const float* gamma, const float* beta, float* gx, float* const* gw,
                           float* gconv_w, float* gconv_b, float* ggamma, float* gbeta, void* ws,
                           size_t ws_bytes, void* stream) {
    return uno_operator_block_bwd2(bd, gy, 0, nullptr, 0, x, xhat, pw_saved, pre, stats, w, conv_w, gamma, beta, gx, gw, gconv_w,
                                   gconv_b, ggamma, gbeta, ws, ws_bytes, stream);
}

int uno_operator_block_bwd2(const uno_block_desc* bd, const float* gy, long gy_batch_stride, const float* gy2,
                            long gy2_batch_stride, const float* x, const float* xhat, const float* pw_saved,
                            const float* pre, const float* stats, const float* const* w, const float* conv_w,
                            const float* gamma, const float* beta, float* gx, float* const* gw, float* gconv_w,
                            float* gconv_b, float* ggamma, float* gbeta, void* ws, size_t ws_bytes, void* stream) {
    ApiRange api_range("uno_operator_block_bwd2");
    if (!bd) return fail(UNO_EINVAL, "null descriptor");
    const uno_conv_desc* d = &bd->conv;
    UNO_TRY(check_conv_desc(d, true));
    if (!gy || !x || !xhat || !w || !conv_w) return fail(UNO_EINVAL, "null tensor pointer");
    long n_out_chk = 1;
    for (int a = 0; a < d->ndim; ++a) n_out_chk *= d->out_dim[a];
    const long contiguous_bs = (long)d->out_ch * n_out_chk;
    if (gy_batch_stride == 0) gy_batch_stride = contiguous_bs;
    if (gy2 && gy2_batch_stride == 0) gy2_batch_stride = contiguous_bs;
    if (gy_batch_stride < contiguous_bs || (gy2 && gy2_batch_stride < contiguous_bs))
        return fail(UNO_EINVAL, "upstream-gradient batch stride smaller than one sample (%ld < %ld)", gy_batch_stride, contiguous_bs);
    if ((bd->normalize || bd->non_lin) && !pre) return fail(UNO_EINVAL, "backward needs the saved pre tensor");
    if (bd->normalize && (!stats || !gamma || !beta || !ggamma || !gbeta))
        return fail(UNO_EINVAL, "normalize backward needs stats, gamma, beta, ggamma, gbeta");
    SpectralPlan* sp;
    ResamplePlan* rp;
    UNO_TRY(get_spectral_plan(d, &sp));
    UNO_TRY(get_resample_plan(d, &rp));
    PwGeom g = pw_geom(d);
    const long planes = (long)d->batch * d->out_ch;
    const size_t nact = (size_t)planes * g.n_out;
    Arena ar(ws, ws_bytes);
    // Gradient w.r.t. conv(x)+w(x).  The upstream gradient may be the sum of two tensors, each a channel slice of a wider one
    // (backend.h UpGrad): the first kernel of the backward -- the activation / normalisation backward -- reads them in place.
    UpGrad up;
    up.p0 = gy; up.bs0 = gy_batch_stride; up.p1 = gy2; up.bs1 = gy2_batch_stride;
    const bool plain = !gy2 && gy_batch_stride == contiguous_bs;
    const float* gs = gy;
    // the conv-bias gradient (gain * per-channel sum of gs) is folded into the kernel that produces gs
    const float bias_gain = g.identity ? 1.0f : (float)rp->gain;
    bool bias_done = false;
    if (bd->normalize) {
        float* buf = ar.take(nact);
        if (!ar.ok) return fail(UNO_EWORKSPACE, "workspace too small for operator block backward");
        BE_TRY(be_memset(ggamma, 0, d->out_ch * sizeof(float), stream));
        BE_TRY(be_memset(gbeta, 0, d->out_ch * sizeof(float), stream));
        // The conv bias feeds an InstanceNorm, which projects the per-plane mean out: its gradient is identically
        // zero in exact arithmetic (every plane of `buf` sums to zero).  The reference's autograd returns the
        // floating-point residue of that sum (~1e-7 of the gradient scale); we return the exact value.
        if (gconv_b) BE_TRY(be_memset(gconv_b, 0, d->out_ch * sizeof(float), stream));
        BE_TRY(be_norm_act_bwd(up, pre, stats, gamma, beta, buf, ggamma, gbeta, planes, d->out_ch, g.n_out, bd->non_lin, stream));
        bias_done = true;
        gs = buf;
    } else if (bd->non_lin || !plain) {
        // GELU backward; for a block without activation whose gradient arrives strided or in two parts, the same kernel in
        // its identity mode gathers / sums the sources into one contiguous tensor
        float* buf = ar.take(nact);
        if (!ar.ok) return fail(UNO_EWORKSPACE, "workspace too small for operator block backward");
        if (gconv_b) BE_TRY(be_memset(gconv_b, 0, d->out_ch * sizeof(float), stream));
        BE_TRY(be_gelu_bwd_bias(up, bd->non_lin ? pre : nullptr, buf, planes, d->out_ch, g.n_out, gconv_b, bias_gain, stream));
        bias_done = gconv_b != nullptr;
        gs = buf;
    }
    // as in forward: the pointwise backward (which overwrites gx) on the side stream, the spectral backward on the
    // caller's stream up to its last stage, which accumulates onto gx after the join
    const size_t pw_floats = pw_ws_floats(d);
    if (ar.off + pw_floats > ar.cap) return fail(UNO_EWORKSPACE, "workspace too small for operator block backward");
    // A third stream takes the two weight-gradient reductions (the 1x1 conv's and the spectral weights'): they only read
    // tensors the other two chains read too and write nothing those chains touch.
    stream_t side = be_side_stream(0);
    stream_t side2 = side ? be_side_stream(1) : nullptr;
    if (side) BE_TRY(be_fork(stream, side));
    {
        Arena sub(ar.base + ar.off, pw_floats * sizeof(float));
        UNO_TRY(pointwise_bwd_impl(d, rp, gs, x, pw_saved, conv_w, gx, gconv_w, bias_done ? nullptr : gconv_b, sub, side ? side : stream, side2));
    }
    {
        Arena sub(ar.base + ar.off + pw_floats, (ar.cap - ar.off - pw_floats) * sizeof(float));
        UNO_TRY(spectral_bwd_impl(d, sp, gs, xhat, w, gx, gw, /*accumulate_gx=*/1, sub, stream, side, side2));
    }
    if (side2) BE_TRY(be_join(stream, side2));
    return 0;
}

// ---- model glue ------------------------------------------------------------------------------------
int uno_lift_check(const uno_lift_desc* d) { LiftArgs a; return make_lift_args(d, &a); }

int uno_lift_fwd(const uno_lift_desc* d, const float* a_, const float* grid, const float* w_a,
                 const float* b_a, const float* w_b, const float* b_b, float* h, void* stream) {
    ApiRange api_range("uno_lift_fwd");
    LiftArgs a;
    UNO_TRY(make_lift_args(d, &a));
    if (!a_ || (!grid && d->grid_ch > 0) || !w_a || !b_a || !w_b || !b_b || !h) return fail(UNO_EINVAL, "null tensor pointer");
    a.a = a_; a.grid = grid; a.w_a = w_a; a.b_a = b_a; a.w_b = w_b; a.b_b = b_b; a.h = h;
    BE_TRY(be_lift_fwd(a, stream));
    return 0;
}

int uno_lift_bwd(const uno_lift_desc* d, const float* gh, const float* a_, const float* grid,
                 const float* w_a, const float* b_a, const float* w_b, const float* b_b, float* ga,
                 float* gw_a, float* gb_a, float* gw_b, float* gb_b, void* stream) {
    return uno_lift_bwd2(d, gh, nullptr, a_, grid, w_a, b_a, w_b, b_b, ga, gw_a, gb_a, gw_b, gb_b, stream);
}

int uno_lift_bwd2(const uno_lift_desc* d, const float* gh, const float* gh2, const float* a_, const float* grid,
                  const float* w_a, const float* b_a, const float* w_b, const float* b_b, float* ga,
                  float* gw_a, float* gb_a, float* gw_b, float* gb_b, void* stream) {
    ApiRange api_range("uno_lift_bwd2");
    LiftArgs a;
    UNO_TRY(make_lift_args(d, &a));
    if (!gh || !a_ || (!grid && d->grid_ch > 0) || !w_a || !b_a || !w_b || !b_b || !gw_a || !gb_a || !gw_b || !gb_b)
        return fail(UNO_EINVAL, "null tensor pointer");
    a.a = a_; a.grid = grid; a.w_a = w_a; a.b_a = b_a; a.w_b = w_b; a.b_b = b_b;
    a.gh = gh; a.gh2 = gh2; a.ga = ga; a.gw_a = gw_a; a.gb_a = gb_a; a.gw_b = gw_b; a.gb_b = gb_b;
    const int cin = d->raw_ch + d->grid_ch;
    BE_TRY(be_memset(gw_a, 0, sizeof(float) * d->hidden * cin, stream));
    BE_TRY(be_memset(gb_a, 0, sizeof(float) * d->hidden, stream));
    BE_TRY(be_memset(gw_b, 0, sizeof(float) * d->out_ch * d->hidden, stream));
    BE_TRY(be_memset(gb_b, 0, sizeof(float) * d->out_ch, stream));
    BE_TRY(be_lift_bwd(a, stream));
    return 0;
}

int uno_project_check(const uno_project_desc* d) { ProjArgs a; return make_proj_args(d, &a); }

int uno_project_fwd(const uno_project_desc* d, const float* const* src, const float* w1,
                    const float* b1, const float* w2, const float* b2, float* out, float* hidden_pre,
                    void* stream) {
    ApiRange api_range("uno_project_fwd");
    ProjArgs a;
    UNO_TRY(make_proj_args(d, &a));
    if (!src || !w1 || !b1 || !w2 || !b2 || !out) return fail(UNO_EINVAL, "null tensor pointer");
    for (int s = 0; s < d->nsrc; ++s) {
        if (!src[s]) return fail(UNO_EINVAL, "null source pointer %d", s);
        a.src[s] = src[s];
    }
    a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2; a.out = out; a.pre_out = hidden_pre;
    BE_TRY(be_proj_fwd(a, stream));
    return 0;
}

int uno_project_bwd(const uno_project_desc* d, const float* gout, const float* const* src,
                    const float* hidden_pre, const float* w1, const float* b1, const float* w2,
                    float* const* gsrc, float* gw1, float* gb1, float* gw2, float* gb2, void* stream) {
    ApiRange api_range("uno_project_bwd");
    ProjArgs a;
    UNO_TRY(make_proj_args(d, &a));
    if (!gout || !src || !w1 || !b1 || !w2 || !gw1 || !gb1 || !gw2 || !gb2) return fail(UNO_EINVAL, "null tensor pointer");
    int ctot = 0;
    for (int s = 0; s < d->nsrc; ++s) {
        if (!src[s]) return fail(UNO_EINVAL, "null source pointer %d", s);
        a.src[s] = src[s];
        a.gsrc[s] = gsrc ? gsrc[s] : nullptr;
        ctot += d->src_ch[s];
    }
    a.w1 = w1; a.b1 = b1; a.w2 = w2; a.gout = gout; a.gw1 = gw1; a.gb1 = gb1; a.gw2 = gw2; a.gb2 = gb2;
    a.pre_in = hidden_pre;
    BE_TRY(be_memset(gw1, 0, sizeof(float) * d->hidden * ctot, stream));
    BE_TRY(be_memset(gb1, 0, sizeof(float) * d->hidden, stream));
    BE_TRY(be_memset(gw2, 0, sizeof(float) * d->out_ch * d->hidden, stream));
    BE_TRY(be_memset(gb2, 0, sizeof(float) * d->out_ch, stream));
    BE_TRY(be_proj_bwd(a, stream));
    return 0;
}

// ---- training-step ops ---------------------------------------------------------------------------------
int uno_adam_step(const uno_adam_tensor* tensors, int n, const uno_adam_hyper* h, void* stream) {
    if (!h || n < 0 || (n > 0 && !tensors)) return fail(UNO_EINVAL, "null argument");
    if (!(h->lr >= 0.0)) return fail(UNO_EINVAL, "Invalid learning rate: %g", h->lr);
    if (!(h->eps >= 0.0)) return fail(UNO_EINVAL, "Invalid epsilon value: %g", h->eps);
    if (!(h->beta1 >= 0.0 && h->beta1 < 1.0)) return fail(UNO_EINVAL, "Invalid beta parameter at index 0: %g", h->beta1);
    if (!(h->beta2 >= 0.0 && h->beta2 < 1.0)) return fail(UNO_EINVAL, "Invalid beta parameter at index 1: %g", h->beta2);
    if (!(h->weight_decay >= 0.0)) return fail(UNO_EINVAL, "Invalid weight_decay value: %g", h->weight_decay);
    if (h->step < 1) return fail(UNO_EINVAL, "step must be >= 1 (got %d)", h->step);
    std::vector<AdamTensor> v((size_t)n);
    for (int i = 0; i < n; ++i) {
        const uno_adam_tensor& a = tensors[i];
        if (!a.param || !a.grad || !a.exp_avg || !a.exp_avg_sq || a.numel < 0) return fail(UNO_EINVAL, "tensor %d: null pointer", i);
        if (a.is_complex && (a.numel & 1)) return fail(UNO_EINVAL, "tensor %d: complex tensor with an odd float count", i);
        if (h->amsgrad && a.is_complex) return fail(UNO_EINVAL, "tensor %d: amsgrad is not defined for complex parameters (torch.maximum)", i);
        if (h->amsgrad && !a.max_exp_avg_sq) return fail(UNO_EINVAL, "tensor %d: amsgrad needs max_exp_avg_sq", i);
        v[i].param = a.param; v[i].grad = a.grad; v[i].exp_avg = a.exp_avg; v[i].exp_avg_sq = a.exp_avg_sq;
        v[i].max_exp_avg_sq = a.max_exp_avg_sq; v[i].numel = a.numel; v[i].is_complex = a.is_complex;
    }
    AdamHyper hh;
    hh.lr = h->lr; hh.beta1 = h->beta1; hh.beta2 = h->beta2; hh.eps = h->eps; hh.weight_decay = h->weight_decay;
    hh.amsgrad = h->amsgrad; hh.step = h->step;
    BE_TRY(be_adam_step(v.data(), n, hh, stream));
    return 0;
}

int uno_lp_loss_fwd(const float* x, const float* y, int batch, long n, int reduction, float* loss, float* norms,
                    void* ws, size_t ws_bytes, void* stream) {
    if (!x || !y || !loss || !norms || !ws) return fail(UNO_EINVAL, "null tensor pointer");
    if (batch < 1 || n < 1 || reduction < 0 || reduction > 2) return fail(UNO_EINVAL, "bad shape / reduction");
    if (ws_bytes < sizeof(double) * 2 * (size_t)batch) return fail(UNO_EWORKSPACE, "workspace too small for the loss");
    BE_TRY(be_lp_loss_fwd(x, y, batch, n, reduction, loss, norms, (double*)ws, stream));
    return 0;
}

int uno_lp_loss_bwd(const float* x, const float* y, const float* norms, const float* gloss, int batch, long n,
                    int reduction, float* gx, void* stream) {
    if (!x || !y || !norms || !gloss || !gx) return fail(UNO_EINVAL, "null tensor pointer");
    if (batch < 1 || n < 1 || reduction < 0 || reduction > 2) return fail(UNO_EINVAL, "bad shape / reduction");
    BE_TRY(be_lp_loss_bwd(x, y, norms, gloss, batch, n, reduction, gx, stream));
    return 0;
}

int uno_config_set(const char* name, int value) {
    if (cfg_set(name, value)) return fail(UNO_EINVAL, "unknown switch '%s'", name ? name : "(null)");
    return 0;
}
int uno_config_get(const char* name, int* value) {
    if (cfg_get(name, value)) return fail(UNO_EINVAL, "unknown switch '%s'", name ? name : "(null)");
    return 0;
}
const char* uno_config_name(int index) { return cfg_name(index); }

void uno_profile_enable(int on) { be_profile_enable(on); }
size_t uno_profile_report(char* buf, size_t cap) { return be_profile_report(buf, cap); }
size_t uno_profile_report_levels(char* buf, size_t cap) { return be_profile_report_scopes(buf, cap); }
long uno_launch_count(void) { return be_launch_count(); }

// ---- host-only planning helpers -------------------------------------------------------------------
int uno_plan_dft_last_analysis(int n, int m, double scale, float* out) {
    if (n < 1 || m < 1 || !out) return fail(UNO_EINVAL, "bad arguments");
    auto v = dft_last_analysis(n, m, scale);
    memcpy(out, v.data(), v.size() * sizeof(float));
    return 0;
}
int uno_plan_dft_last_synthesis(int n, int m, double scale, int hermitian, float* out) {
    if (n < 1 || m < 1 || !out) return fail(UNO_EINVAL, "bad arguments");
    auto v = dft_last_synthesis(n, m, scale, hermitian != 0);
    memcpy(out, v.data(), v.size() * sizeof(float));
    return 0;
}
int uno_plan_dft_mid_analysis(int n, int m, float* out) {
    if (n < 1 || m < 1 || !out) return fail(UNO_EINVAL, "bad arguments");
    auto v = dft_mid_analysis(n, m);
    memcpy(out, v.data(), v.size() * sizeof(float));
    return 0;
}
int uno_plan_dft_mid_synthesis(int n, int m, float* out) {
    if (n < 1 || m < 1 || !out) return fail(UNO_EINVAL, "bad arguments");
    auto v = dft_mid_synthesis(n, m);
    memcpy(out, v.data(), v.size() * sizeof(float));
    return 0;
}
int uno_plan_sr_mid(int n_in, int n_out, float* out) {
    if (n_in < 1 || n_out < 1 || !out) return fail(UNO_EINVAL, "bad arguments");
    auto v = sr_mid(n_in, n_out);
    memcpy(out, v.data(), v.size() * sizeof(float));
    return 0;
}
int uno_plan_sr_last_modes(int n_in, int n_out) { return sr_last_modes(n_in, n_out); }
int uno_plan_sr_mid_fixed(int n_in, int n_out, float* out) {
    if (n_in < 1 || n_out < 1 || !out) return fail(UNO_EINVAL, "bad arguments");
    auto v = sr_mid_fixed(n_in, n_out);
    memcpy(out, v.data(), v.size() * sizeof(float));
    return 0;
}
int uno_plan_sr_last_modes_fixed(int n_in, int n_out) { return sr_last_modes_fixed(n_in, n_out); }
int uno_plan_bicubic_aa(int n_in, int n_out, int transpose, float* out) {
    if (n_in < 1 || n_out < 1 || !out) return fail(UNO_EINVAL, "bad arguments");
    Banded b = bicubic_aa(n_in, n_out);
    if (transpose) b = banded_transpose(b);
    auto v = banded_dense(b);
    memcpy(out, v.data(), v.size() * sizeof(float));
    return 0;
}

int uno_plan_band_groups(int n_in, int n_out, int transpose, float* out, int* gw) {
    if (n_in < 1 || n_out < 1 || !out || !gw) return fail(UNO_EINVAL, "bad arguments");
    Banded b = bicubic_aa(n_in, n_out);
    if (transpose) b = banded_transpose(b);
    BandGroups g = band_groups(b);
    gw[0] = g.ok ? g.G : 0;
    gw[1] = g.ok ? g.W : 0;
    if (!g.ok) return 0;
    auto v = band_groups_dense(g);
    memcpy(out, v.data(), v.size() * sizeof(float));
    return 0;
}

}  // extern "C"
