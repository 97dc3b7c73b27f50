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

struct LfPlan {
    LfCfg dev;                 // pointers filled by the caller once the tables are placed
    int L, nv, ns;             // actual group counts
    int nv_t, ns_t;            // template bucket used (>= actual)
    int nfit, ncont;
    std::vector<double> axis;  // [nz]
    std::vector<float> xo;     // [nfit]
    std::vector<float> trans;  // [nz] or empty
    // workspace table offsets (bytes)
    size_t off_axis, off_xo, off_trans, table_bytes;
    // per-warp shared scratch (bytes)
    size_t off_u, off_v, per_warp_bytes, cont_cap;
};

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

static inline size_t lf_hess_bytes(int L)
{
    size_t pf = 3 * (size_t)L + 1;
    size_t a = 13 * (size_t)L * sizeof(LfLine);
    size_t b = (pf * pf + 2 * pf) * sizeof(double);
    return a > b ? a : b;
}

// Returns "" on success, else an error message.
static inline std::string lf_make_plan(const lucifit_config* c, LfPlan& pl)
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
    d.max_iter = c->max_iter > 0 ? c->max_iter : 60;
    d.ftol = c->ftol > 0 ? (float)c->ftol : 1e-8f;
    d.n_out = 7 * L + 5;
    pl.axis.assign(c->axis, c->axis + c->nz);
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
    // per-warp scratch: spectrum | U = max(continuum clip scratch, Hessian scratch) | V = max(LM stash, tile)
    size_t cap = (size_t)(pl.ncont > LF_MAX_CONT_WIN ? LF_MAX_CONT_WIN : pl.ncont);
    cap = lf_align(cap > 0 ? cap : 1, 8);
    pl.cont_cap = cap;
    size_t u_bytes = 2 * cap * sizeof(double) + cap;
    const int P = L + 2 * pl.nv_t + 1;
    size_t v_bytes = (size_t)(P * (P + 1) / 2 + P) * sizeof(float);
    if (d.uncertainty) {
        size_t hb = lf_hess_bytes(L);
        if (hb > u_bytes) u_bytes = hb;
        size_t tb = 32 * (size_t)((3 * L + 1) | 1) * sizeof(float);
        if (tb > v_bytes) v_bytes = tb;
    }
    pl.off_u = lf_align(sizeof(float) * c->nz, 16);
    pl.off_v = lf_align(pl.off_u + u_bytes, 16);
    pl.per_warp_bytes = lf_align(pl.off_v + v_bytes, 16);
    return "";
}

static inline LfScratch lf_scratch_at(const LfPlan& pl, unsigned char* warp_base)
{
    LfScratch sc;
    unsigned char* u = warp_base + pl.off_u;
    sc.dvals = reinterpret_cast<double*>(u);
    sc.ddev = sc.dvals + pl.cont_cap;
    sc.alive = reinterpret_cast<unsigned char*>(sc.ddev + pl.cont_cap);
    sc.hess = reinterpret_cast<double*>(u);
    sc.st = reinterpret_cast<float*>(warp_base + pl.off_v);
    sc.tile = sc.st;
    return sc;
}
