// lucifit_host.hpp -- host-side translation of the public lucifit_config into the device constants (LfCfg),
// the device-table layout inside the caller's workspace and the per-warp scratch layout.  Shared by the C ABI
// (lucifit_capi.cu) and by the CPU test harness (tests/hostemu/hostemu.cpp).
#pragma once
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/lucifit.h"
#include "lucifit_core.cuh"
#include "lucifit_prior.cuh"

struct LfPlan {
    LfCfg dev;                 // pointers filled by the caller once the tables are placed
    int L, nv, ns;             // actual group counts
    int nv_t, ns_t;            // template bucket used (>= actual)
    int nfit, ncont;
    std::vector<double> axis;  // [nz]
    std::vector<float> xo;     // [nfit]
    std::vector<float> trans;  // [nz] or empty
    // workspace table offsets (bytes); the per-spaxel state records follow the tables
    size_t off_axis, off_xo, off_trans, table_bytes;
    size_t state_bytes;        // per spaxel: state record (+ float64 uncertainties when requested)
    LfLayout ly;               // per-warp scratch layout of the four stage kernels (for `lanes` lanes)
};

// Spaxels per pass of the four stage kernels: bounds the state area of the workspace.
#define LF_SLICE_SPAXELS (1 << 20)

static inline size_t lf_align(size_t v, size_t a) { return (v + a - 1) / a * a; }

static inline int lf_dense_groups(const int* labels, int L, int* dense)
{
    int n = 0;
    int seen[LF_MAX_LINES];
    for (int k = 0; k < L; ++k) {
        int g = -1;
        for (int j = 0; j < n; ++j) if (seen[j] == labels[k]) g = j;
        if (g < 0) { seen[n] = labels[k]; g = n++; }
        dense[k] = g;
    }
    return n;
}

// Smallest pre-instantiated group-count bucket that can hold (nv, ns): (1,1), (2,2) or (L,L).
static inline int lf_bucket(int L, int nv, int ns)
{
    int m = nv > ns ? nv : ns;
    if (m <= 1) return 1;
    if (m <= 2 && L >= 2) return 2;
    return L;
}

// Per-warp scratch layout for `lanes` lanes per spaxel (32 on the device, 1 in the CPU harness).
static inline void lf_make_layout(int L, int P, int nz, int nfit, int ncont, int lanes, LfLayout& ly)
{
    const size_t NA = (size_t)P * (P + 1) / 2 + P, PF = 3 * (size_t)L + 1;
    int cap = ncont > LF_MAX_CONT_WIN ? LF_MAX_CONT_WIN : ncont;
    int npad = 16;                             // >= 2 LF_REFINE_HALF + 1: the prep stage also keeps its refinement scores here
    while (npad < cap) npad <<= 1;
    ly.lanes = lanes;
    ly.npad = npad;
    const size_t win = lf_align(sizeof(float) * nfit, 16);
    size_t o;
    // prep: spec | sorted window
    o = lf_align(sizeof(float) * nz, 16);
    ly.prep_sortv = (int)o; o += sizeof(float) * npad;
    ly.prep_bytes = (int)lf_align(o, 16);
    // lm: yw | sw | cw | lines[L] | tie | gt[L][lanes] | hb[2][NA] + S[P][P] + 3 [P] | thb[4][P]
    o = win;
    ly.lm_sw = (int)o; o += win;
    ly.lm_cw = (int)o; o += win;
    ly.lm_lines = (int)o; o += sizeof(LfLineX) * L;
    ly.lm_tie = (int)o; o = lf_align(o + sizeof(LfTieX), 16);
    ly.lm_gt = (int)o; o = lf_align(o + sizeof(float) * 2 * L * lanes, 16);                   // unit profiles + their (v, beta) partials as bf16 pairs
    ly.lm_hb = (int)o; o = lf_align(o + sizeof(float) * (2 * (NA + 2 * (size_t)L) + (size_t)P * P + 3 * P), 16);   // [2][NA + 2 L] + secant term S[P][P] and three [P] vectors
    ly.lm_thb = (int)o; o = lf_align(o + sizeof(float) * 4 * P, 16);
    ly.lm_bytes = (int)o;
    // unc: variants[13 L] (later: hess) | apd[3 L] | gt[(13 + 5 L)][lanes] | tile[lanes][PF|1]   (no staged window:
    // the stage forms data and phase chunk by chunk)
    o = 0;
    ly.un_sw = 0; ly.un_cw = 0;
    {   // the float64 Hessian is assembled after the sample loop, when the variant records are dead: same bytes
        size_t rec = sizeof(LfLineV) * 13 * L, hess = sizeof(double) * (PF * PF + 2 * PF);
        ly.un_lines = (int)o; ly.un_hess = (int)o; o = lf_align(o + (rec > hess ? rec : hess), 16);
    }
    ly.un_apd = (int)o; o = lf_align(o + sizeof(double) * 3 * L, 16);
    ly.un_gt = (int)o; o = lf_align(o + sizeof(float) * (13 + 5 * L) * lanes, 16);
    ly.un_tile = (int)o; o = lf_align(o + sizeof(float) * PF * (lanes + 4), 16);     // D tile [PF][lanes + 4]
    ly.un_bytes = (int)o;
    // out: yw | sw | cw | lines[L] | apd[3 L]
    o = win;
    ly.out_sw = (int)o; o += win;
    ly.out_cw = (int)o; o += win;
    ly.out_lines = (int)o; o += sizeof(LfLineX) * L;
    ly.out_apd = (int)o; o = lf_align(o + sizeof(double) * 3 * L, 16);
    ly.out_bytes = (int)o;
}

// Returns "" on success, else an error message.
static inline std::string lf_make_plan(const lucifit_config* c, LfPlan& pl, int lanes = 32)
{
    if (!c) return "config is NULL";
    if (c->model < 0 || c->model > 2) return "model must be 0 (gaussian), 1 (sinc) or 2 (sincgauss)";
    const int L = c->n_lines;
    if (L < 1 || L > LF_MAX_LINES) return "n_lines must be in [1, 8]";
    if (c->nz < 8 || !c->axis) return "axis missing or nz < 8";
    if (!(c->fit_lo >= 0 && c->fit_hi <= c->nz && c->fit_hi - c->fit_lo > 3 * L + 1))
        return "fit window must hold more than 3L+1 samples";
    if (!(c->noise_lo >= 0 && c->noise_hi <= c->nz && c->noise_lo <= c->noise_hi)) return "bad noise window";
    if (!(c->cont_lo >= 0 && c->cont_hi <= c->nz && c->cont_lo <= c->cont_hi)) return "bad continuum window";
    if (c->cont_hi - c->cont_lo > LF_MAX_CONT_WIN)
        return "continuum window longer than 1024 samples (LuciFit.py:450-482 clips the whole window; it is not truncated here)";
    if (!(c->step_nm > 0) || !(c->n_steps - c->zpd_index > 0)) return "STEP and (STEPNB - ZPDINDEX) must be positive";
    memset(&pl.dev, 0, sizeof(pl.dev));
    LfCfg& d = pl.dev;
    d.model = c->model; d.L = L; d.nz = c->nz;
    d.fit_lo = c->fit_lo; d.fit_hi = c->fit_hi; d.noise_lo = c->noise_lo; d.noise_hi = c->noise_hi;
    d.cont_lo = c->cont_lo; d.cont_hi = c->cont_hi;
    pl.L = L;
    pl.nv = lf_dense_groups(c->vel_group, L, d.vg);
    pl.ns = lf_dense_groups(c->sigma_group, L, d.sg);
    d.nv = pl.nv; d.ns = pl.ns;
    pl.nv_t = pl.ns_t = lf_bucket(L, pl.nv, pl.ns);
    d.i48 = d.i83 = -1;
    if (c->nii6548_idx >= 0 || c->nii6583_idx >= 0) {
        if (c->nii6548_idx < 0 || c->nii6548_idx >= L || c->nii6583_idx < 0 || c->nii6583_idx >= L ||
            c->nii6548_idx == c->nii6583_idx)
            return "nii6548_idx / nii6583_idx must both be valid and distinct, or both -1";
        d.i48 = c->nii6548_idx; d.i83 = c->nii6583_idx;
    }
    d.sigbox_group = d.sg[L - 1];
    d.uncertainty = c->uncertainty ? 1 : 0;
    d.ml_scale = c->scale_mode ? 1 : 0;
    d.prior_groups = c->prior_groups ? 1 : 0;
    d.freeze = c->freeze ? 1 : 0;
    if (d.freeze) {
        if (c->model != LUCIFIT_SINCGAUSS) return "freeze is defined for the sincgauss model only (the reference's gaussian / sinc freeze paths fail)";
        if (c->uncertainty) return "freeze with uncertainty_bool is broken in the reference (LuciFit.py:713-722 on an L+1 vector) and not offered";
        d.i48 = d.i83 = -1;                            // no constraints in the freeze branch (LuciFit.py:699-703)
        d.ml_scale = 0;                                // scale = max(spectrum) (LuciFit.py:776-782: interpolate_spectrum is skipped)
    }
    d.refine_prior = c->no_prior_refine ? 0 : 1;
    d.max_iter = c->max_iter > 0 ? c->max_iter : 60;
#ifndef LF_DEFAULT_FTOL
#define LF_DEFAULT_FTOL 1e-8f
#endif
    d.ftol = c->ftol > 0 ? (float)c->ftol : LF_DEFAULT_FTOL;
    d.n_out = 7 * L + 5;
    pl.axis.assign(c->axis, c->axis + c->nz);
    for (int i = 1; i < c->nz; ++i)
        if (!(pl.axis[i] > pl.axis[i - 1])) return "axis must be strictly increasing (cm^-1)";
    d.x_ref = pl.axis[c->fit_lo];
    for (int k = 0; k < L; ++k) {
        if (!(c->line_nm[k] > 0)) return "line_nm must be positive";
        d.line_nm[k] = c->line_nm[k];
        double p0 = 1e7 / c->line_nm[k];
        d.p0[k] = (float)p0;
        d.p0_off[k] = (float)(p0 - d.x_ref);
        d.p0_lo[k] = (float)((p0 - d.x_ref) - (double)d.p0_off[k]);
    }
    d.axis_k = 1e7 / (2.0 * c->step_nm * (c->n_steps - c->zpd_index));
    pl.nfit = c->fit_hi - c->fit_lo;
    d.ichan = (float)((double)(pl.nfit - 1) / (pl.axis[c->fit_hi - 1] - d.x_ref));
    pl.ncont = c->cont_hi - c->cont_lo;
    pl.xo.resize(pl.nfit);
    for (int n = 0; n < pl.nfit; ++n) pl.xo[n] = (float)(pl.axis[c->fit_lo + n] - d.x_ref);
    pl.trans.clear();
    if (c->transmission) {
        pl.trans.resize(c->nz);
        for (int i = 0; i < c->nz; ++i) pl.trans[i] = (float)c->transmission[i];
    }
    pl.off_axis = 0;
    pl.off_xo = lf_align(pl.off_axis + sizeof(double) * c->nz, 256);
    pl.off_trans = lf_align(pl.off_xo + sizeof(float) * pl.nfit, 256);
    pl.table_bytes = lf_align(pl.off_trans + (c->transmission ? sizeof(float) * c->nz : 0), 256);
    const int P = L + 2 * pl.nv_t + 1;
    d.state_floats = (int)lf_align((size_t)LF_SO_TH + P, 2);
    pl.state_bytes = sizeof(float) * d.state_floats + (d.uncertainty ? sizeof(double) * (3 * (size_t)L + 1) : 0);
    lf_make_layout(L, P, c->nz, pl.nfit, pl.ncont, lanes, pl.ly);
    return "";
}

// Constants of lucifit_estimate_prior from the public config (velocity grid in km/s).  Returns "" on success.
static inline std::string lf_make_prior_cfg(const lucifit_config* c, double v_min, double v_max, double v_step, LfPriorCfg& pc)
{
    if (!c) return "config is NULL";
    if (c->model < 0 || c->model > 2) return "model must be 0 (gaussian), 1 (sinc) or 2 (sincgauss)";
    const int L = c->n_lines;
    if (L < 1 || L > LF_MAX_LINES) return "n_lines must be in [1, 8]";
    if (c->nz < 8 || !c->axis) return "axis missing or nz < 8";
    if (!(c->fit_lo >= 0 && c->fit_hi <= c->nz && c->fit_hi - c->fit_lo >= 8)) return "fit window must hold at least 8 samples";
    if (!(c->step_nm > 0) || !(c->n_steps - c->zpd_index > 0)) return "STEP and (STEPNB - ZPDINDEX) must be positive";
    if (!(v_step > 0) || !(v_max >= v_min)) return "velocity grid needs v_step > 0 and v_max >= v_min";
    const double nv = floor((v_max - v_min) / v_step + 0.5) + 1.0;
    if (nv > LF_PRIOR_MAX_GRID) return "velocity grid has more than 512 points";
    memset(&pc, 0, sizeof(pc));
    pc.model = c->model; pc.L = L; pc.nz = c->nz; pc.fit_lo = c->fit_lo; pc.nfit = c->fit_hi - c->fit_lo;
    if (pc.nfit > 4096) return "fit window too long for the prior estimator's shared-memory layout";
    pc.nv = (int)nv; pc.v_min = (float)v_min; pc.v_step = (float)v_step;
    pc.inv_vmax = (float)(1.0 / fmax(fmax(fabs(v_min), fabs(v_max)), 1.0));
    const double x_ref = c->axis[c->fit_lo];
    for (int i = c->fit_lo + 1; i < c->fit_hi; ++i)
        if (!(c->axis[i] > c->axis[i - 1])) return "axis must be strictly increasing (cm^-1)";
    for (int k = 0; k < L; ++k) {
        if (!(c->line_nm[k] > 0)) return "line_nm must be positive";
        pc.p0[k] = (float)(1e7 / c->line_nm[k]);
        pc.p0_off[k] = (float)(1e7 / c->line_nm[k] - x_ref);
    }
    pc.chan = (float)((c->axis[c->fit_hi - 1] - x_ref) / (double)(pc.nfit - 1));
    pc.axis_k = 1e7 / (2.0 * c->step_nm * (c->n_steps - c->zpd_index));
    pc.bkg = nullptr;
    return "";
}
