// TEST INFRASTRUCTURE ONLY -- compiles the device fit pipeline (luci_b200/csrc/lucifit_core.cuh) for the CPU
// with LANES = 1 so that the GPU-less test tier can exercise the very same prep / LM / uncertainty / output code
// against the oracle.  Never loaded by the luci_b200 package.
#include <stdlib.h>
#include <vector>

#include "../../luci_b200/csrc/lucifit_host.hpp"
#include "../../luci_b200/csrc/lucifit_instances.h"

template <int MODEL, int L, int NG>
static void run_batch(const LfPlan& pl, int64_t n, const float* cube, const int64_t* index, const float* theta,
                      const float* vel0, const float* broad0, float* out_rows, float* out_sol, float* out_unc,
                      int32_t* status)
{
    LfWarp<1> w;
    const LfLayout& ly = pl.ly;
    size_t bytes = ly.prep_bytes;
    if ((size_t)ly.lm_bytes > bytes) bytes = ly.lm_bytes;
    if ((size_t)ly.un_bytes > bytes) bytes = ly.un_bytes;
    if ((size_t)ly.out_bytes > bytes) bytes = ly.out_bytes;
    std::vector<unsigned char> buf(bytes + 64);
    unsigned char* base = buf.data();
    base += (16 - ((uintptr_t)base & 15)) & 15;
    const int K = 7 * L + 5, PF = 3 * L + 1;
    std::vector<double> state_d(pl.dev.state_floats / 2 + 1);
    float* state = reinterpret_cast<float*>(state_d.data());
    std::vector<double> unc(PF + 1);
    for (int64_t s = 0; s < n; ++s) {
        int64_t row = index ? index[s] : s;
        const float* gspec = cube + row * (int64_t)pl.dev.nz;
        const int pnv = pl.dev.prior_groups ? pl.nv : 1, pns = pl.dev.prior_groups ? pl.ns : 1;
        lf_stage_prep<MODEL, L, NG, NG, 1>(pl.dev, w, gspec, theta[s], vel0 ? vel0 + s * pnv : 0, broad0 ? broad0 + s * pns : 0,
                                           lf_work_at(LF_STAGE_PREP, ly, base), state);
        lf_stage_lm<MODEL, L, NG, NG, 1, 1>(pl.dev, w, pl.dev.xo, gspec, lf_work_at(LF_STAGE_LM, ly, base), state);
        lf_stage_lm<MODEL, L, NG, NG, 1, 2>(pl.dev, w, pl.dev.xo, gspec, lf_work_at(LF_STAGE_LM, ly, base), state);
        if (pl.dev.uncertainty)
            lf_stage_unc<MODEL, L, NG, NG, 1>(pl.dev, w, pl.dev.xo, gspec, lf_work_at(LF_STAGE_UNC, ly, base), state, unc.data());
        lf_stage_out<MODEL, L, NG, NG, 1>(pl.dev, w, pl.dev.xo, gspec, lf_work_at(LF_STAGE_OUT, ly, base), state,
                                          pl.dev.uncertainty ? unc.data() : 0, out_rows + s * K, status + s,
                                          out_sol ? out_sol + s * PF : 0, out_unc ? out_unc + s * PF : 0);
    }
}

extern "C" int hostemu_fit_batch(const lucifit_config* cfg, int64_t n, const float* cube, const int64_t* index,
                                 const float* theta, const float* vel0, const float* broad0, const float* bkg,
                                 float* out_rows, float* out_sol, float* out_unc, int32_t* status)
{
    LfPlan pl;
    std::string err = lf_make_plan(cfg, pl, 1);
    if (!err.empty()) { fprintf(stderr, "hostemu: %s\n", err.c_str()); return -1; }
    // the uncertainty scratch is always reserved in the harness
    pl.dev.axis = pl.axis.data();
    pl.dev.xo = pl.xo.data();
    pl.dev.trans = pl.trans.empty() ? 0 : pl.trans.data();
    pl.dev.bkg = bkg;
#define X(L_, G_)                                                                                            \
    if (pl.L == L_ && pl.nv_t == G_) {                                                                       \
        if (cfg->model == 0) run_batch<0, L_, G_>(pl, n, cube, index, theta, vel0, broad0, out_rows, out_sol, out_unc, status); \
        else if (cfg->model == 1) run_batch<1, L_, G_>(pl, n, cube, index, theta, vel0, broad0, out_rows, out_sol, out_unc, status); \
        else run_batch<2, L_, G_>(pl, n, cube, index, theta, vel0, broad0, out_rows, out_sol, out_unc, status); \
        return 0;                                                                                            \
    }
    LF_INSTANCES(X)
#undef X
    fprintf(stderr, "hostemu: tie pattern not instantiated (L=%d groups=%d)\n", pl.L, pl.nv_t);
    return -1;
}

// unit-profile value and partials (for math unit tests)
extern "C" void hostemu_profile(int model, float pp, float sig, float sinc_w, int n, const float* x, float* g,
                                float* gp, float* gs)
{
    LfLine ln;
    lf_line_setup(model, pp, 0.0f, sig, sinc_w, ln);
    float piw = LF_PI_F / sinc_w, iw = 1.0f / sinc_w;
    for (int i = 0; i < n; ++i) {
        float dx = lf_dx(ln, x[i]);
        if (model == 0) lf_profile<0>(ln, dx, piw, iw, false, g[i], gp[i], gs[i]);
        else if (model == 1) lf_profile<1>(ln, dx, piw, iw, false, g[i], gp[i], gs[i]);
        else lf_profile<2>(ln, dx, piw, iw, lf_is_far(ln, dx), g[i], gp[i], gs[i]);
    }
}

extern "C" void hostemu_faddeeva(int n, const float* x, const float* y, float* wr, float* wi)
{
    for (int i = 0; i < n; ++i) lf_faddeeva(x[i], y[i], wr[i], wi[i]);
}

// Debug/test entry: sum r^2, J^T r and J^T J of the LM objective at an arbitrary reduced parameter vector `th`
// (layout [A | v | beta | continuum] of the (L, bucket) instance) for spectrum 0 of `cube`.
template <int MODEL, int L, int NG>
static void run_eval(const LfPlan& pl, const float* gspec, float theta, const float* th_in, double* chi2, float* hg)
{
    typedef LfDims<MODEL, L, NG, NG> D;
    LfWarp<1> w;
    const LfLayout& ly = pl.ly;
    size_t bytes = ly.prep_bytes > ly.lm_bytes ? ly.prep_bytes : ly.lm_bytes;
    std::vector<unsigned char> buf(bytes + 64);
    unsigned char* base = buf.data();
    base += (16 - ((uintptr_t)base & 15)) & 15;
    std::vector<double> state_d(pl.dev.state_floats / 2 + 1);
    float* state = reinterpret_cast<float*>(state_d.data());
    float v0 = 0.f, b0 = 30.f;
    lf_stage_prep<MODEL, L, NG, NG, 1>(pl.dev, w, gspec, theta, &v0, &b0, lf_work_at(LF_STAGE_PREP, ly, base), state);
    LfWork wk = lf_work_at(LF_STAGE_LM, ly, base);
    const double sinc_w_d = reinterpret_cast<const double*>(state)[LF_SO_SINCW / 2];
    lf_stage_window<MODEL, 1>(pl.dev, w, pl.dev.xo, gspec, state[LF_SO_INVW], false, sinc_w_d, wk, 0);
    float* th = wk.thb;
    for (int j = 0; j < D::P; ++j) th[j] = th_in[j];
    lf_setup_lines<MODEL, L, NG, NG, 1>(pl.dev, w, th, (float)sinc_w_d, 1.0 / sinc_w_d, wk.lines, wk.tie);
    lf_eval<MODEL, L, NG, NG, 1>(pl.dev, w, pl.dev.xo, wk, pl.dev.fit_hi - pl.dev.fit_lo, (float)sinc_w_d, th[D::IC], *chi2, wk.hb);
    for (int e = 0; e < D::NA; ++e) hg[e] = wk.hb[e];
}

extern "C" int hostemu_eval(const lucifit_config* cfg, const float* gspec, float theta, const float* th, double* chi2, float* hg)
{
    LfPlan pl;
    std::string err = lf_make_plan(cfg, pl, 1);
    if (!err.empty()) return -1;
    pl.dev.axis = pl.axis.data(); pl.dev.xo = pl.xo.data(); pl.dev.trans = 0; pl.dev.bkg = 0;
    if (pl.L == 5 && pl.nv_t == 1 && cfg->model == 2) { run_eval<2, 5, 1>(pl, gspec, theta, th, chi2, hg); return 0; }
    if (pl.L == 5 && pl.nv_t == 1 && cfg->model == 0) { run_eval<0, 5, 1>(pl, gspec, theta, th, chi2, hg); return 0; }
    return -2;
}

// The prior estimator (lucifit_prior.cuh) with LANES = 1.
extern "C" int hostemu_estimate_prior(const lucifit_config* cfg, int64_t n, const float* cube, const float* theta,
                                      const float* bkg, double v_min, double v_max, double v_step, float* vel, float* broad)
{
    LfPriorCfg pc;
    std::string err = lf_make_prior_cfg(cfg, v_min, v_max, v_step, pc);
    if (!err.empty()) { fprintf(stderr, "hostemu: %s\n", err.c_str()); return -1; }
    pc.bkg = bkg;
    std::vector<float> z(pc.nfit), sm(pc.nfit), score(pc.nv);
    LfWarp<1> w;
    for (int64_t s = 0; s < n; ++s)
        lf_estimate_prior<1>(pc, w, cube + s * (int64_t)pc.nz, theta[s], z.data(), sm.data(), score.data(), vel + s, broad + s);
    return 0;
}
