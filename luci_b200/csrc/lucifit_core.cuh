// lucifit_core.cuh -- per-spaxel pipeline executed by ONE WARP: prep -> Levenberg-Marquardt -> derived outputs.
//
// Replaces, for one spectrum, LUCI.LuciFit.Fit.__init__/fit (LUCI/LuciFit.py:49-173, 761-835):
//   normalise + window      LuciFit.py:110-112, 225-280        -> lf_prep
//   noise                   LuciFit.py:282-331                 -> lf_prep
//   continuum estimate      LuciFit.py:432-486                 -> lf_cont_estimate (MAD sigma-clip + median)
//   line estimates          LuciFit.py:382-430                 -> lf_prep
//   SLSQP on the nll        LuciFit.py:635-687                 -> lf_lm  (tie-eliminated, box-projected LM)
//   chi2 / flux / velocity  LuciFit.py:799-834, 1050-1069,
//                           LUCI/LuciFitParameters.py:14-185   -> lf_outputs
//
// The tie constraints of the reference (LuciFit.py:523-598) are eliminated analytically: the unknowns are
//   theta = [A_0..A_{L-1} | v_0..v_{NV-1} | beta_0..beta_{NS-1} | continuum]
// with line centre p_k = 1e7/(lambda_k (1 + v_g/c)) and sigma_k = p_k beta_h / c, so equal velocity /
// equal broadening inside a group hold identically, and (if enabled) A_6548 = A_6583 G(sigma_6583)/(3 G(sigma_6548)).
// Line centres are carried as offsets from x_ref = axis[fit_lo] so FP32 never subtracts two 1.5e4-sized numbers.
//
// Lane model: LANES = 32 on the device (one warp per spaxel, sample n owned by lane n % 32), LANES = 1 in the
// CPU test harness (tests/hostemu) which compiles this very file with g++.
#pragma once
#include <stdint.h>
#include "lucifit_math.cuh"

#if defined(__CUDACC__)
#define LF_CORE __device__ __forceinline__
#define LF_CORE_NOINLINE __device__ __noinline__
#else
#define LF_CORE inline
#define LF_CORE_NOINLINE
#endif

#define LF_MAX_LINES 8
#define LF_C_KMS 299792.0f        /* LuciFit.py:26 (integer km/s in the reference) */
#define LF_MAX_CONT_WIN 256       /* samples of the continuum window handled by the clip */

// status word: bits 0-7 accepted LM iterations, bits 8-15 model evaluations, bits 16.. flags
enum {
    LF_ST_CONVERGED = 1 << 16, LF_ST_MAXITER = 1 << 17, LF_ST_BOUND = 1 << 18, LF_ST_SINGULAR = 1 << 19,
    LF_ST_NAN_INPUT = 1 << 20, LF_ST_MASKED = 1 << 21, LF_ST_BAD_NORM = 1 << 22, LF_ST_NAN_RESULT = 1 << 23
};

struct LfCfg {                     // device-side constants of one lucifit_fit_batch call
    int model, L, nv, ns;
    int nz, fit_lo, fit_hi, noise_lo, noise_hi, cont_lo, cont_hi;
    int vg[LF_MAX_LINES], sg[LF_MAX_LINES];   // dense 0-based group index of each line
    int i48, i83;                  // [NII] tie: indices of NII6548 / NII6583, -1 when off
    int sigbox_group;              // sigma group of the LAST line: 0 <= sigma <= 10 cm^-1 (LuciFit.py:539-541)
    int uncertainty, ml_scale, max_iter, n_out;
    float ftol;
    float p0[LF_MAX_LINES];        // 1e7/lambda_k                     [cm^-1]
    float p0_off[LF_MAX_LINES];    // 1e7/lambda_k - x_ref             [cm^-1]
    float p0_lo[LF_MAX_LINES];     // FP32 remainder of the above
    double line_nm[LF_MAX_LINES];
    double x_ref;
    double axis_k;                 // 1e7/(2 STEP (n_steps - zpd)); axis_step = sinc_width = corr * axis_k
    const double* axis;            // [nz]  float32-rounded axis values, widened
    const float* xo;               // [n_fit] axis[fit_lo + n] - x_ref (kernels stage it into shared memory)
    const float* trans;            // [nz] or null
    const float* bkg;              // [nz] or null (already multiplied by binning^2)
};

// ------------------------------------------------------------------------------------------------ lanes
template <int LANES> struct LfWarp;

template <> struct LfWarp<1> {
    int lane;
    LfWarp() : lane(0) {}
    float sum(float v) const { return v; }
    double sum(double v) const { return v; }
    float max(float v) const { return v; }
    int sum(int v) const { return v; }
    bool all(bool p) const { return p; }
    bool any(bool p) const { return p; }
    void argmin(double& v, int& i) const {}
    void sync() const {}
};

#if defined(__CUDACC__)
template <> struct LfWarp<32> {
    int lane;
    __device__ LfWarp() : lane(threadIdx.x & 31) {}
    __device__ float sum(float v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ double sum(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ int sum(int v) const { return __reduce_add_sync(0xffffffffu, v); }
    __device__ float max(float v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    __device__ bool all(bool p) const { return __all_sync(0xffffffffu, p); }
    __device__ bool any(bool p) const { return __any_sync(0xffffffffu, p); }
    __device__ void argmin(double& v, int& i) const {   // smallest v, ties -> smallest index (np.argmin)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, v, o);
            int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
    }
    __device__ void sync() const { __syncwarp(); }
};
#endif

// ------------------------------------------------------------------------------------------- small helpers
template <int P> struct LfSym {            // packed upper triangle, row-major
    static LF_HD constexpr int idx(int i, int j) { return i * P - (i * (i - 1)) / 2 + (j - i); }
    enum { N = P * (P + 1) / 2 };
};

// index of the axis sample nearest to `pos` (first one on ties, like np.argmin(|axis - pos|)), all lanes.
template <int LANES>
LF_CORE int lf_argmin_axis(const LfWarp<LANES>& w, const double* axis, int nz, double pos)
{
    double best = 1e300;
    int bi = 0x7fffffff;
    for (int i = w.lane; i < nz; i += LANES) {
        double d = fabs(axis[i] - pos);
        if (d < best) { best = d; bi = i; }
    }
    w.argmin(best, bi);
    return bi;
}

// k-th smallest (0-based) of the alive entries of v[0..n), by rank counting; all lanes get the result.
template <int LANES>
LF_CORE double lf_kth(const LfWarp<LANES>& w, const double* v, const unsigned char* alive, int n, int k)
{
    double found = 0.0;
    int hit = 0;
    for (int i = w.lane; i < n; i += LANES) {
        if (!alive[i]) continue;
        double vi = v[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            if (!alive[j]) continue;
            double vj = v[j];
            rank += (vj < vi) || (vj == vi && j < i);
        }
        if (rank == k) { found = vi; hit = 1; }
    }
    // exactly one lane hits
    double r = w.sum(hit ? found : 0.0);
    return r;
}

template <int LANES>
LF_CORE double lf_median(const LfWarp<LANES>& w, const double* v, const unsigned char* alive, int n, int m)
{
    if (m <= 0) return 0.0;
    if (m & 1) return lf_kth(w, v, alive, n, m / 2);
    return 0.5 * (lf_kth(w, v, alive, n, m / 2 - 1) + lf_kth(w, v, alive, n, m / 2));
}

// Continuum guess: median of the window after <=3 rounds of 1-sigma clipping about the median with
// mad_std as the spread (LuciFit.py:474-486; astropy sigma_clip semantics restated in oracle/astro_stats.py).
// `vals`/`dev`/`alive` are per-warp scratch of LF_MAX_CONT_WIN entries.
template <int LANES>
LF_CORE double lf_cont_estimate(const LfWarp<LANES>& w, const float* spec, int lo, int hi, double sigma_level,
                                double* vals, double* dev, unsigned char* alive)
{
    int n = hi - lo;
    if (n > LF_MAX_CONT_WIN) n = LF_MAX_CONT_WIN;
    if (n <= 0) return 0.0;
    for (int i = w.lane; i < n; i += LANES) { vals[i] = (double)spec[lo + i]; alive[i] = 1; }
    w.sync();
    int m = n;
    for (int it = 0; it < 3 && m > 0; ++it) {
        double med = lf_median(w, vals, alive, n, m);
        for (int i = w.lane; i < n; i += LANES) dev[i] = fabs(vals[i] - med);
        w.sync();
        double mad = 1.482602218505602 * lf_median(w, dev, alive, n, m);
        double lo_b = med - sigma_level * mad, hi_b = med + sigma_level * mad;
        int removed = 0;
        for (int i = w.lane; i < n; i += LANES)
            if (alive[i] && !(vals[i] >= lo_b && vals[i] <= hi_b)) { alive[i] = 0; ++removed; }
        removed = w.sum(removed);
        w.sync();
        m -= removed;
        if (removed == 0) break;
    }
    if (m < 1) {                                   // LuciFit.py:481-482: fall back to the unclipped window
        for (int i = w.lane; i < n; i += LANES) alive[i] = 1;
        w.sync();
        m = n;
    }
    return lf_median(w, vals, alive, n, m);
}


// Area ratio of the [NII] doublet tie, A_6548 = R A_6583 (LuciFit.py:584-595): gaussian & sinc
// R = sigma_83/(3 sigma_48) = p_83 beta_83/(3 p_48 beta_48) (beta cancels inside one sigma group);
// sincgauss R = G(sigma_83)/(3 G(sigma_48)), G = sigma/erf(sigma/(sqrt2 w)).
LF_HD float lf_tie_ratio(int model, float p83, float b83, float s83, float p48, float b48, float s48,
                         bool same_group, float sinc_w)
{
    if (model == LF_SINCGAUSS) return lf_G(model, s83, sinc_w) / (3.0f * lf_G(model, s48, sinc_w));
    if (same_group) return p83 / (3.0f * p48);
    return (p83 * b83) / (3.0f * p48 * fmaxf(b48, 1e-6f));
}

// --------------------------------------------------------------------------------------------- per spaxel
struct LfPrep {
    float smax;        // max of the (background-subtracted) spectrum
    float wmax;        // max inside the fit window
    float inv_wmax;
    double noise;      // std of the normalised spectrum in the noise window
    double cont0;      // continuum guess (normalised by smax)
    double scale;      // spectrum_scale (LuciFit.py:374-375 / :782)
    double corr;       // 1/cos(theta)
    double sinc_w;     // == axis_step
    int bad;           // status flags raised during prep
};

template <int MODEL, int L, int NV, int NS> struct LfDims {
    enum { P = L + NV + NS + 1, IV = L, IS = L + NV, IC = L + NV + NS, NH = P * (P + 1) / 2 };
};

// Reads the spaxel's spectrum into `spec` (per-warp, nz floats) and computes everything Fit.__init__ does.
template <int LANES>
LF_CORE void lf_prep(const LfCfg& cfg, const LfWarp<LANES>& w, const float* gspec, float theta_deg,
                     float* spec, double* sc_vals, double* sc_dev, unsigned char* sc_alive, LfPrep& pr)
{
    const int nz = cfg.nz;
    float smax = -INFINITY, wmax = -INFINITY, tmax = -INFINITY;
    int nan_ct = 0;
    for (int i = w.lane; i < nz; i += LANES) {
        float s = gspec[i];
        if (cfg.bkg) s -= cfg.bkg[i];
        spec[i] = s;
        nan_ct += (s != s);
        smax = fmaxf(smax, s);
        if (i >= cfg.fit_lo && i < cfg.fit_hi) wmax = fmaxf(wmax, s);
        float st = s;
        if (cfg.trans) { float t = cfg.trans[i]; if (t > 0.5f) st = s / t; }
        tmax = fmaxf(tmax, st);
    }
    smax = w.max(smax); wmax = w.max(wmax); tmax = w.max(tmax);
    nan_ct = w.sum(nan_ct);
    w.sync();
    pr.bad = 0;
    if (nan_ct > 0) pr.bad |= LF_ST_NAN_INPUT;
    if (!(smax > 0.0f) || !(wmax > 0.0f) || !isfinite(smax)) pr.bad |= LF_ST_BAD_NORM;
    pr.smax = smax; pr.wmax = wmax; pr.inv_wmax = 1.0f / wmax;
    // noise: population std of spectrum_normalized over the noise window (LuciFit.py:327-331)
    {
        int n = cfg.noise_hi - cfg.noise_lo;
        double s1 = 0.0;
        for (int i = cfg.noise_lo + w.lane; i < cfg.noise_hi; i += LANES) s1 += (double)spec[i];
        s1 = w.sum(s1);
        double mean = n > 0 ? s1 / n : 0.0;
        double s2 = 0.0;
        for (int i = cfg.noise_lo + w.lane; i < cfg.noise_hi; i += LANES) { double d = (double)spec[i] - mean; s2 += d * d; }
        s2 = w.sum(s2);
        pr.noise = n > 0 ? sqrt(s2 / n) / (double)smax : NAN;
    }
    pr.cont0 = lf_cont_estimate(w, spec, cfg.cont_lo, cfg.cont_hi, 1.0, sc_vals, sc_dev, sc_alive) / (double)smax;
    // scale: ML-prior path max(spectrum_restricted) * max(spectrum_T) ; plain path max(spectrum_T)
    pr.scale = cfg.ml_scale ? ((double)wmax / (double)smax) * (double)tmax : (double)tmax;
    double ct = fabs(cos((double)theta_deg * 0.017453292519943295));
    pr.corr = 1.0 / ct;
    pr.sinc_w = pr.corr * cfg.axis_k;
}

// One pass over the fit window: chi2 = sum r^2, H = J^T J (packed upper), gr = J^T r at parameters th.
template <int MODEL, int L, int NV, int NS, int LANES>
LF_CORE void lf_eval(const LfCfg& cfg, const LfWarp<LANES>& w, const float* xo, const float* spec, float inv_wmax, float sinc_w,
                     const float (&th)[LfDims<MODEL, L, NV, NS>::P], double& chi2_out,
                     float (&H)[LfDims<MODEL, L, NV, NS>::NH], float (&gr)[LfDims<MODEL, L, NV, NS>::P])
{
    typedef LfDims<MODEL, L, NV, NS> D;
    constexpr int P = D::P;
    LfLine ln[L];
    float aeff[L], dpdv[L], dsdv[L], dsdb[L], pk[L], bk[L];
    bool tie = cfg.i48 >= 0;
    float R = 0.0f, cRv[NV], cRb[NS];
#pragma unroll
    for (int g = 0; g < NV; ++g) cRv[g] = 0.0f;
#pragma unroll
    for (int h = 0; h < NS; ++h) cRb[h] = 0.0f;
#pragma unroll
    for (int k = 0; k < L; ++k) {
        int g = NV == 1 ? 0 : cfg.vg[k], h = NS == 1 ? 0 : cfg.sg[k];
        float v = 0.0f, beta = 0.0f;
#pragma unroll
        for (int gg = 0; gg < NV; ++gg) if (gg == g) v = th[D::IV + gg];
#pragma unroll
        for (int hh = 0; hh < NS; ++hh) if (hh == h) beta = th[D::IS + hh];
        float vt = v * (1.0f / LF_C_KMS);
        float i1v = 1.0f / (1.0f + vt);
        float p = cfg.p0[k] * i1v;
        float sig = p * beta * (1.0f / LF_C_KMS);
        lf_line_setup(MODEL, cfg.p0_off[k], cfg.p0_lo[k] - cfg.p0[k] * vt * i1v, sig, sinc_w, ln[k]);
        pk[k] = p; bk[k] = beta;
        dpdv[k] = -p * i1v * (1.0f / LF_C_KMS);
        dsdv[k] = beta * (1.0f / LF_C_KMS) * dpdv[k];
        dsdb[k] = p * (1.0f / LF_C_KMS);
        aeff[k] = th[k];
    }
    if (tie && MODEL == LF_SINC) {
        // sinc: the constraint A_83 sigma_83/3 = A_48 sigma_48 (LuciFit.py:587-589) is vacuous when sigma = 0
        // (no prior broadening, the reference's ML_bool=False path) -> both amplitudes are free
#pragma unroll
        for (int k = 0; k < L; ++k) if ((k == cfg.i83 || k == cfg.i48) && bk[k] == 0.0f) tie = false;
    }
    if (tie) {
        // A_6548 = R A_6583, R = G(sigma_6583)/(3 G(sigma_6548)); ln G = ln p + ln beta - ln c [- ln erf(u)]
        float p83 = 1.f, p48 = 1.f, b83 = 1.f, b48 = 1.f, s83 = 1.f, s48 = 1.f, a83 = 0.f;
        float dv83 = 0.f, dv48 = 0.f, db83 = 0.f, db48 = 0.f, dp83 = 0.f, dp48 = 0.f;
#pragma unroll
        for (int k = 0; k < L; ++k) {
            if (k == cfg.i83) { p83 = pk[k]; b83 = bk[k]; s83 = ln[k].sig; a83 = th[k]; dv83 = dsdv[k]; db83 = dsdb[k]; dp83 = dpdv[k]; }
            if (k == cfg.i48) { p48 = pk[k]; b48 = bk[k]; s48 = ln[k].sig; dv48 = dsdv[k]; db48 = dsdb[k]; dp48 = dpdv[k]; }
        }
        int g83 = NV == 1 ? 0 : cfg.vg[cfg.i83], g48 = NV == 1 ? 0 : cfg.vg[cfg.i48];
        int h83 = NS == 1 ? 0 : cfg.sg[cfg.i83], h48 = NS == 1 ? 0 : cfg.sg[cfg.i48];
        R = lf_tie_ratio(MODEL, p83, b83, s83, p48, b48, s48, h83 == h48, sinc_w);
        float E83 = 0.f, E48 = 0.f;
        if (MODEL == LF_SINCGAUSS) { E83 = 1.0f / s83 - lf_dlnG(MODEL, s83, sinc_w); E48 = 1.0f / s48 - lf_dlnG(MODEL, s48, sinc_w); }
#pragma unroll
        for (int g = 0; g < NV; ++g)
            cRv[g] = a83 * R * ((g == g83 ? dp83 / p83 - E83 * dv83 : 0.f) - (g == g48 ? dp48 / p48 - E48 * dv48 : 0.f));
#pragma unroll
        for (int h = 0; h < NS; ++h) {
            float t = 0.f;
            if (h83 != h48) t = (h == h83 ? 1.0f / fmaxf(b83, 1e-6f) : 0.f) - (h == h48 ? 1.0f / fmaxf(b48, 1e-6f) : 0.f);
            cRb[h] = a83 * R * (t - (h == h83 ? E83 * db83 : 0.f) + (h == h48 ? E48 * db48 : 0.f));
        }
#pragma unroll
        for (int k = 0; k < L; ++k) if (k == cfg.i48) aeff[k] = R * a83;
    }
    const float piw = LF_PI_F / sinc_w, iw = 1.0f / sinc_w;
    const float cont = th[D::IC];
    double chi2 = 0.0;
#pragma unroll
    for (int e = 0; e < D::NH; ++e) H[e] = 0.0f;
#pragma unroll
    for (int j = 0; j < P; ++j) gr[j] = 0.0f;
    const int nfit = cfg.fit_hi - cfg.fit_lo;
    const float* y = spec + cfg.fit_lo;
    for (int n0 = 0; n0 < nfit; n0 += LANES) {
        int n = n0 + w.lane;
        bool live = n < nfit;
        float x = live ? xo[n] : 0.0f;
        float yn = live ? y[n] * inv_wmax : 0.0f;
        float J[P];
#pragma unroll
        for (int j = 0; j < P; ++j) J[j] = 0.0f;
        J[D::IC] = 1.0f;
        float m = cont, g48v = 0.0f;
#pragma unroll
        for (int k = 0; k < L; ++k) {
            float dx = lf_dx(ln[k], x);
            bool all_far = false;
            if (MODEL == LF_SINCGAUSS) all_far = w.all(lf_is_far(ln[k], dx));
            float g, gp, gs;
            lf_profile<MODEL>(ln[k], dx, piw, iw, all_far, g, gp, gs);
            float A = aeff[k];
            m += A * g;
            float tp = A * gp, ts = A * gs;
            float tv = tp * dpdv[k] + ts * dsdv[k];
            float tb = ts * dsdb[k];
            if (tie && k == cfg.i48) { g48v = g; } else { J[k] += g; }
#pragma unroll
            for (int gg = 0; gg < NV; ++gg) if (NV == 1 || cfg.vg[k] == gg) J[D::IV + gg] += tv;
            if (MODEL != LF_SINC) {
#pragma unroll
                for (int hh = 0; hh < NS; ++hh) if (NS == 1 || cfg.sg[k] == hh) J[D::IS + hh] += tb;
            }
        }
        if (tie) {
#pragma unroll
            for (int k = 0; k < L; ++k) if (k == cfg.i83) J[k] += R * g48v;
#pragma unroll
            for (int gg = 0; gg < NV; ++gg) J[D::IV + gg] += g48v * cRv[gg];
            if (MODEL != LF_SINC) {
#pragma unroll
                for (int hh = 0; hh < NS; ++hh) J[D::IS + hh] += g48v * cRb[hh];
            }
        }
        if (live) {
            float r = yn - m;
            chi2 += (double)r * (double)r;
#pragma unroll
            for (int i = 0; i < P; ++i) {
                gr[i] += J[i] * r;
#pragma unroll
                for (int j = i; j < P; ++j) H[LfSym<P>::idx(i, j)] += J[i] * J[j];
            }
        }
    }
    chi2_out = w.sum(chi2);
#pragma unroll
    for (int e = 0; e < D::NH; ++e) H[e] = w.sum(H[e]);
#pragma unroll
    for (int j = 0; j < P; ++j) gr[j] = w.sum(gr[j]);
}

// Solve (D^-1/2 H D^-1/2 + lam I) s = D^-1/2 g with fixed columns pinned, delta = D^-1/2 s.  Returns false if the
// Cholesky factorisation meets a non-positive pivot.
template <int P>
LF_CORE bool lf_solve(const float (&H)[P * (P + 1) / 2], const float (&gr)[P], unsigned fixedmask, float lam,
                      float (&delta)[P], float& pred)
{
    float sc[P], M[P * (P + 1) / 2], b[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        float d = H[LfSym<P>::idx(j, j)];
        bool fx = ((fixedmask >> j) & 1u) || !(d > 1e-30f);
        sc[j] = fx ? 0.0f : LF_RSQRT(d);
        b[j] = gr[j] * sc[j];
    }
#pragma unroll
    for (int i = 0; i < P; ++i)
#pragma unroll
        for (int j = i; j < P; ++j) {
            float v = H[LfSym<P>::idx(i, j)] * sc[i] * sc[j];
            if (i == j) v = (sc[i] == 0.0f) ? 1.0f : v + lam;
            M[LfSym<P>::idx(i, j)] = v;
        }
    // in-place Cholesky M = U^T U (upper), then forward/back substitution
    bool ok = true;
#pragma unroll
    for (int i = 0; i < P; ++i) {
        float d = M[LfSym<P>::idx(i, i)];
#pragma unroll
        for (int k = 0; k < i; ++k) d -= M[LfSym<P>::idx(k, i)] * M[LfSym<P>::idx(k, i)];
        if (!(d > 1e-12f)) { ok = false; d = 1.0f; }
        float inv = LF_RSQRT(d);
        M[LfSym<P>::idx(i, i)] = inv;                      // store 1/U_ii
#pragma unroll
        for (int j = i + 1; j < P; ++j) {
            float v = M[LfSym<P>::idx(i, j)];
#pragma unroll
            for (int k = 0; k < i; ++k) v -= M[LfSym<P>::idx(k, i)] * M[LfSym<P>::idx(k, j)];
            M[LfSym<P>::idx(i, j)] = v * inv;
        }
    }
    float z[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {                          // U^T z = b
        float v = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v -= M[LfSym<P>::idx(k, i)] * z[k];
        z[i] = v * M[LfSym<P>::idx(i, i)];
    }
    float s[P];
#pragma unroll
    for (int i = P - 1; i >= 0; --i) {                     // U s = z
        float v = z[i];
#pragma unroll
        for (int k = i + 1; k < P; ++k) v -= M[LfSym<P>::idx(i, k)] * s[k];
        s[i] = v * M[LfSym<P>::idx(i, i)];
    }
    pred = 0.0f;                                           // predicted decrease of sum r^2: s.(lam s + b)
#pragma unroll
    for (int j = 0; j < P; ++j) { pred += s[j] * (lam * s[j] + b[j]); delta[j] = s[j] * sc[j]; }
    return ok;
}

struct LfFitResult {
    double chi2;       // sum of squared residuals of the normalised window at the solution
    int iters, evals, flags;
};

// Box of the reduced parameters (LuciFit.py:614-633 amplitudes/continuum; :539-541 sigma of the last line).
template <int MODEL, int L, int NV, int NS>
LF_CORE void lf_bounds(const LfCfg& cfg, const float (&th)[LfDims<MODEL, L, NV, NS>::P],
                       float (&lo)[LfDims<MODEL, L, NV, NS>::P], float (&hi)[LfDims<MODEL, L, NV, NS>::P])
{
    typedef LfDims<MODEL, L, NV, NS> D;
#pragma unroll
    for (int k = 0; k < L; ++k) { lo[k] = -1e-8f; hi[k] = 1.1f; }
#pragma unroll
    for (int g = 0; g < NV; ++g) { lo[D::IV + g] = -INFINITY; hi[D::IV + g] = INFINITY; }
    // sigma box applies to the LAST line only (late-binding lambda in the reference); other groups: beta >= 0
    float vlast = 0.0f;
#pragma unroll
    for (int g = 0; g < NV; ++g) if (NV == 1 || cfg.vg[L - 1] == g) vlast = th[D::IV + g];
    float plast = cfg.p0[L - 1] / (1.0f + vlast * (1.0f / LF_C_KMS));
#pragma unroll
    for (int h = 0; h < NS; ++h) {
        lo[D::IS + h] = 0.0f;
        hi[D::IS + h] = (h == cfg.sigbox_group) ? 10.0f * LF_C_KMS / plast : INFINITY;
    }
    lo[D::IC] = -1e-8f; hi[D::IC] = 0.99f;
}

// Levenberg-Marquardt on sum r^2 (the noise is a common factor of the reference's nll, LuciFit.py:516, and does
// not move the optimum).  `st` = per-warp scratch holding the current (H, g) so a rejected step can be re-solved.
template <int MODEL, int L, int NV, int NS, int LANES>
LF_CORE void lf_lm(const LfCfg& cfg, const LfWarp<LANES>& w, const float* xo, const float* spec, float inv_wmax, float sinc_w,
                   float (&th)[LfDims<MODEL, L, NV, NS>::P], float* st, LfFitResult& res)
{
    typedef LfDims<MODEL, L, NV, NS> D;
    constexpr int P = D::P, NH = D::NH;
    float H[NH], gr[P];
    double chi2;
    float lo[P], hi[P];
    lf_bounds<MODEL, L, NV, NS>(cfg, th, lo, hi);
#pragma unroll
    for (int j = 0; j < P; ++j) th[j] = fminf(fmaxf(th[j], lo[j]), hi[j]);
    lf_eval<MODEL, L, NV, NS, LANES>(cfg, w, xo, spec, inv_wmax, sinc_w, th, chi2, H, gr);
    int evals = 1, iters = 0, flags = 0;
    float lam = 1e-3f, nu = 2.0f;
    const double ftol = (double)cfg.ftol;
    bool done = false;
    while (!done && iters < cfg.max_iter && evals < 250) {
        // stash current normal equations (warp-uniform) so a rejection can re-solve without re-evaluating
#pragma unroll
        for (int e = 0; e < NH; ++e) if ((e % LANES) == w.lane || LANES == 1) st[e] = H[e];
#pragma unroll
        for (int j = 0; j < P; ++j) if (((NH + j) % LANES) == w.lane || LANES == 1) st[NH + j] = gr[j];
        w.sync();
        unsigned fixedmask = 0;
#pragma unroll
        for (int j = 0; j < P; ++j) {
            bool at_lo = th[j] <= lo[j] && gr[j] <= 0.0f;
            bool at_hi = th[j] >= hi[j] && gr[j] >= 0.0f;
            if (at_lo || at_hi) fixedmask |= 1u << j;
        }
        bool accepted = false;
        for (int tries = 0; tries < 12 && !accepted && !done; ++tries) {
            float delta[P], pred;
            bool ok = lf_solve<P>(H, gr, fixedmask, lam, delta, pred);
            if (!ok) { lam *= 10.0f; flags |= LF_ST_SINGULAR; continue; }
            float tht[P];
            float rel = 0.0f;
#pragma unroll
            for (int j = 0; j < P; ++j) {
                tht[j] = fminf(fmaxf(th[j] + delta[j], lo[j]), hi[j]);
                float sc = (j >= D::IV && j < D::IS) ? 1.0f : fmaxf(fabsf(th[j]), 1e-3f);   // velocities: km/s absolute
                rel = fmaxf(rel, fabsf(tht[j] - th[j]) / sc);
            }
            // The predicted decrease is below what an FP32 evaluation of sum r^2 can resolve: take the (Newton-like)
            // step and stop -- evaluating it would only return rounding noise.
            if ((double)pred <= ftol * chi2 && rel < 1e-2f) {
#pragma unroll
                for (int j = 0; j < P; ++j) th[j] = tht[j];
                done = true; flags |= LF_ST_CONVERGED;
                ++iters;
                break;
            }
            double chi2t;
            float Ht[NH], grt[P];
            lf_eval<MODEL, L, NV, NS, LANES>(cfg, w, xo, spec, inv_wmax, sinc_w, tht, chi2t, Ht, grt);
            ++evals;
            bool finite = chi2t == chi2t;
#if defined(LF_TRACE)
            printf("  it %d try %d lam %.2e chi2 %.10e chi2t %.10e pred %.3e rel %.2e\n", iters, tries, lam, chi2, chi2t, pred, rel);
#endif
            // An FP32 evaluation resolves sum r^2 to a few 1e-6 relative; once the predicted decrease is that small
            // the gradient (still accurate) is trusted and the step is taken even if chi2 appears to rise by noise.
            const double slack = ((double)pred <= 1e-4 * chi2) ? 4e-6 * chi2 : 0.0;
            if (finite && chi2t <= chi2 + slack) {
                double act = chi2 - chi2t;
                float rho = pred > 0.0f ? (float)(act / (double)pred) : 1.0f;
                float f = 1.0f - (2.0f * rho - 1.0f) * (2.0f * rho - 1.0f) * (2.0f * rho - 1.0f);
                lam *= fmaxf(1.0f / 3.0f, f);
                lam = fmaxf(lam, 1e-7f);
                nu = 2.0f;
#pragma unroll
                for (int j = 0; j < P; ++j) { th[j] = tht[j]; gr[j] = grt[j]; }
#pragma unroll
                for (int e = 0; e < NH; ++e) H[e] = Ht[e];
                accepted = true;
                ++iters;
                chi2 = chi2t;
            } else {
                // rejected: restore the stashed system and increase the damping; inside the FP32 noise floor of
                // sum r^2 (relative ~1e-6) a rejection only means "converged"
#pragma unroll
                for (int e = 0; e < NH; ++e) H[e] = st[e];
#pragma unroll
                for (int j = 0; j < P; ++j) gr[j] = st[NH + j];
                lam *= nu; nu *= 2.0f;
                if (finite && (double)pred <= 1e-5 * chi2) { done = true; flags |= LF_ST_CONVERGED; }
            }
        }
        if (!accepted && !done) { done = true; flags |= LF_ST_MAXITER; }
        if (accepted) lf_bounds<MODEL, L, NV, NS>(cfg, th, lo, hi);
    }
    if (!done) flags |= LF_ST_MAXITER;
#pragma unroll
    for (int j = 0; j < P; ++j) if (th[j] <= lo[j] || th[j] >= hi[j]) flags |= LF_ST_BOUND;
    res.chi2 = chi2; res.iters = iters; res.evals = evals; res.flags = flags;
}

// Expand the reduced solution to the reference's fit_sol layout [A, p, sigma]*L + [continuum] in normalised units.
template <int MODEL, int L, int NV, int NS>
LF_CORE void lf_expand(const LfCfg& cfg, float sinc_w, const float (&th)[LfDims<MODEL, L, NV, NS>::P],
                       double (&A)[L], double (&p)[L], double (&sg)[L], double (&vel)[L], double (&beta)[L])
{
    typedef LfDims<MODEL, L, NV, NS> D;
#pragma unroll
    for (int k = 0; k < L; ++k) {
        float v = 0.f, b = 0.f;
#pragma unroll
        for (int g = 0; g < NV; ++g) if (NV == 1 || cfg.vg[k] == g) v = th[D::IV + g];
#pragma unroll
        for (int h = 0; h < NS; ++h) if (NS == 1 || cfg.sg[k] == h) b = th[D::IS + h];
        vel[k] = (double)v; beta[k] = (double)b;
        p[k] = 1e7 / (cfg.line_nm[k] * (1.0 + (double)v / 299792.0));
        sg[k] = p[k] * (double)b / 299792.0;
        A[k] = (double)th[k];
    }
    bool tie = cfg.i48 >= 0;
    if (tie && MODEL == LF_SINC) {
#pragma unroll
        for (int k = 0; k < L; ++k) if ((k == cfg.i83 || k == cfg.i48) && beta[k] == 0.0) tie = false;
    }
    if (tie) {
        float p83 = 1.f, p48 = 1.f, b83 = 1.f, b48 = 1.f, s83 = 1.f, s48 = 1.f;
        double a83 = 0;
        LfLine t;
#pragma unroll
        for (int k = 0; k < L; ++k) {
            if (k == cfg.i83) { p83 = (float)p[k]; b83 = (float)beta[k]; lf_line_setup(MODEL, 0.f, 0.f, (float)sg[k], sinc_w, t); s83 = t.sig; a83 = A[k]; }
            if (k == cfg.i48) { p48 = (float)p[k]; b48 = (float)beta[k]; lf_line_setup(MODEL, 0.f, 0.f, (float)sg[k], sinc_w, t); s48 = t.sig; }
        }
        bool same = (NS == 1) || cfg.sg[cfg.i83] == cfg.sg[cfg.i48];
        float R = lf_tie_ratio(MODEL, p83, b83, s83, p48, b48, s48, same, sinc_w);
#pragma unroll
        for (int k = 0; k < L; ++k) if (k == cfg.i48) A[k] = a83 * (double)R;
    }
}

// ---------------------------------------------------------------------------------------------- outputs
// Row layout of the per-spaxel result (K = 7 L + 5 floats), the 12 quantities Luci.fit_calc collects
// (LuciBase.py:381-392, 406):
//   [0,L) amplitudes | [L,2L) fluxes | [2L,3L) flux errors | [3L,4L) velocities | [4L,5L) velocity errors |
//   [5L,6L) broadenings | [6L,7L) broadening errors | chi2 | corr | axis_step | continuum | continuum error
template <int MODEL, int L, int NV, int NS, int LANES>
LF_CORE void lf_outputs(const LfCfg& cfg, const LfWarp<LANES>& w, const float* xo, const float* spec, const LfPrep& pr,
                        const float (&th)[LfDims<MODEL, L, NV, NS>::P], const double* unc /* [3L+1] normalised or null */,
                        float* out_row, float* out_sol, float* out_unc)
{
    typedef LfDims<MODEL, L, NV, NS> D;
    const float sinc_w = (float)pr.sinc_w;
    double A[L], p[L], sg[L], vel[L], beta[L];
    lf_expand<MODEL, L, NV, NS>(cfg, sinc_w, th, A, p, sg, vel, beta);
    const double cont = (double)th[D::IC];
    // chi2 with every line centre snapped to the nearest axis sample (Model.plot, LuciFunctions.py:113-115 etc.;
    // LuciFit.py:799-801, 1065-1068).  The scale cancels: chi2 = sum (s_T/scale - m)^2 / (noise (N - P_full)).
    LfLine ln[L];
#pragma unroll
    for (int k = 0; k < L; ++k) {
        int idx = lf_argmin_axis(w, cfg.axis, cfg.nz, p[k]);
        float pps, ppl;
        lf_split(cfg.axis[idx] - cfg.x_ref, pps, ppl);
        lf_line_setup(MODEL, pps, ppl, (float)sg[k], sinc_w, ln[k]);
    }
    const float piw = LF_PI_F / sinc_w, iw = 1.0f / sinc_w;
    const int nfit = cfg.fit_hi - cfg.fit_lo;
    const float inv_scale = (float)(1.0 / pr.scale);
    double S = 0.0;
    for (int n0 = 0; n0 < nfit; n0 += LANES) {
        int n = n0 + w.lane;
        bool live = n < nfit;
        float x = live ? xo[n] : 0.0f;
        float m = (float)cont;
#pragma unroll
        for (int k = 0; k < L; ++k) {
            float dx = lf_dx(ln[k], x);
            bool all_far = false;
            if (MODEL == LF_SINCGAUSS) all_far = w.all(lf_is_far(ln[k], dx));
            float g, gp, gs;
            lf_profile<MODEL>(ln[k], dx, piw, iw, all_far, g, gp, gs);
            m += (float)A[k] * g;
        }
        if (live) {
            float s = spec[cfg.fit_lo + n];
            if (cfg.trans) { float t = cfg.trans[cfg.fit_lo + n]; if (t > 0.5f) s = s / t; }
            float r = s * inv_scale - m;
            S += (double)r * (double)r;
        }
    }
    S = w.sum(S);
    const double chi2 = S / (pr.noise * (double)(nfit - (3 * L + 1)));
    const double FWHM_COEFF = 2.3548200450309493;
    const double SQ2PI = 2.5066282746310002, PI = 3.141592653589793;
    float row[7 * L + 5];
#pragma unroll
    for (int k = 0; k < L; ++k) {
        double Ak = A[k] * pr.scale, s = sg[k], pk = p[k];
        double uA = unc ? unc[3 * k] * pr.scale : 0.0, up = unc ? unc[3 * k + 1] : 0.0, us = unc ? unc[3 * k + 2] : 0.0;
        double flux, ferr;
        if (MODEL == LF_GAUSSIAN) {
            flux = (1.20671 / FWHM_COEFF) * SQ2PI * Ak * s;
            ferr = flux * sqrt((uA / Ak) * (uA / Ak) + (us / s) * (us / s));
        } else if (MODEL == LF_SINC) {
            flux = PI * sqrt(PI) * Ak * pr.sinc_w;
            ferr = flux * sqrt((uA / Ak) * (uA / Ak) + (us / s) * (us / s));
        } else {
            double c0 = 1.4142135623730951 * pr.sinc_w;
            double ef = erf(s / c0);
            flux = (1.20671 / (PI * FWHM_COEFF)) * Ak * (SQ2PI * s / ef);
            double t = ef - (2.0 * PI) / (sqrt(PI) * c0) * exp(-s * s / (c0 * c0));
            ferr = flux * sqrt((uA / Ak) * (uA / Ak) + (us / s) * (us / s) * t * t);
        }
        double lam = cfg.line_nm[k];
        double v = 299792.0 * ((1e7 / pk - lam) / lam);                         // LuciFitParameters.py:34-36
        double verr = 299792.0 * up * (1e7 / (lam * pk * pk));                  // :57
        double broad = pk > 0 ? fabs(299792.0 * s / pk) : 0.0;                  // :82-88
        double berr = (pk > 0 && s > 0) ? (299792.0 * s / pk) * sqrt((us / s) * (us / s) + (up / pk) * (up / pk)) : 0.0;
        row[k] = (float)Ak; row[L + k] = (float)flux; row[2 * L + k] = (float)ferr;
        row[3 * L + k] = (float)v; row[4 * L + k] = (float)verr;
        row[5 * L + k] = (float)broad; row[6 * L + k] = (float)berr;
    }
    row[7 * L] = (float)chi2;
    row[7 * L + 1] = (float)pr.corr;
    row[7 * L + 2] = (float)pr.sinc_w;             // axis_step == sinc_width (LuciFit.py:214-215, 222-223)
    row[7 * L + 3] = (float)(cont * pr.scale);
    row[7 * L + 4] = (float)(unc ? unc[3 * L] * pr.scale : 0.0);
#pragma unroll
    for (int j = 0; j < 7 * L + 5; ++j) if (LANES == 1 || (j % LANES) == w.lane) out_row[j] = row[j];
    if (out_sol) {
#pragma unroll
        for (int k = 0; k < L; ++k) {
            if (w.lane == 0) {
                out_sol[3 * k] = (float)(A[k] * pr.scale);
                out_sol[3 * k + 1] = (float)p[k];
                out_sol[3 * k + 2] = (float)sg[k];
            }
        }
        if (w.lane == 0) out_sol[3 * L] = (float)(cont * pr.scale);
    }
    if (out_unc) {
        for (int j = w.lane; j < 3 * L + 1; j += LANES) {
            double u = unc ? unc[j] : 0.0;
            if (j == 3 * L || (j % 3) == 0) u *= pr.scale;
            out_unc[j] = (float)u;
        }
    }
}

// Per-warp scratch layout (bytes), after the nz-float spectrum buffer.
struct LfScratch {
    double* dvals;            // LF_MAX_CONT_WIN
    double* ddev;             // LF_MAX_CONT_WIN
    unsigned char* alive;     // LF_MAX_CONT_WIN
    float* st;                // NH + P   (stash of the normal equations)
    double* hess;             // (3L+1)^2 + 2(3L+1)  (uncertainty stage)
    float* tile;              // LANES x (3L+2)       (uncertainty stage)
};

template <int MODEL, int L, int NV, int NS, int LANES>
LF_CORE void lf_uncertainty(const LfCfg& cfg, const LfWarp<LANES>& w, const float* xo, const float* spec, const LfPrep& pr,
                            const float (&th)[LfDims<MODEL, L, NV, NS>::P], const LfScratch& sc, double* unc);

// The whole pipeline for one spaxel.
template <int MODEL, int L, int NV, int NS, int LANES>
LF_CORE void lf_fit_spaxel(const LfCfg& cfg, const LfWarp<LANES>& w, const float* xo, const float* gspec, float theta_deg,
                           float vel0, float broad0, float* spec, const LfScratch& sc,
                           float* out_row, int32_t* out_status, float* out_sol, float* out_unc)
{
    typedef LfDims<MODEL, L, NV, NS> D;
    LfPrep pr;
    lf_prep(cfg, w, gspec, theta_deg, spec, sc.dvals, sc.ddev, sc.alive, pr);
    if (pr.bad) {                                      // reference returns zeros for unusable spectra
        for (int j = w.lane; j < 7 * L + 5; j += LANES) out_row[j] = 0.0f;
        if (out_sol) for (int j = w.lane; j < 3 * L + 1; j += LANES) out_sol[j] = 0.0f;
        if (out_unc) for (int j = w.lane; j < 3 * L + 1; j += LANES) out_unc[j] = 0.0f;
        if (w.lane == 0) *out_status = pr.bad;
        return;
    }
    // initial vector (LuciFit.py:663-672, 410-423): amplitude = max over +-5 channels around the expected
    // centre minus the continuum guess; centres from the prior velocity; sigma from the prior broadening.
    float th[D::P];
    if (vel0 != vel0) vel0 = 0.0f;                     // LuciFit.py:406-409
    if (broad0 != broad0) broad0 = 1.0f;
    broad0 = fabsf(broad0);
#pragma unroll
    for (int k = 0; k < L; ++k) {
        double lam = cfg.line_nm[k];
        double pos = 1e7 / (((double)vel0 / 299792.0) * lam + lam);
        int ind = lf_argmin_axis(w, cfg.axis, cfg.nz, pos);
        float amp;
        if (ind + 5 > cfg.nz - 1) {
            amp = spec[ind];                           // IndexError branch (LuciFit.py:420-421)
        } else {
            amp = -INFINITY;
            for (int d = -5; d <= 5; ++d) { int i = ind + d; if (i < 0) i += cfg.nz; amp = fmaxf(amp, spec[i]); }
        }
        th[k] = (float)((double)amp / (double)pr.smax - pr.cont0);
    }
#pragma unroll
    for (int g = 0; g < NV; ++g) th[D::IV + g] = vel0;
#pragma unroll
    for (int h = 0; h < NS; ++h) th[D::IS + h] = broad0;
    th[D::IC] = (float)pr.cont0;
    LfFitResult res;
    lf_lm<MODEL, L, NV, NS, LANES>(cfg, w, xo, spec, pr.inv_wmax, (float)pr.sinc_w, th, sc.st, res);
    const double* unc = 0;
    if (cfg.uncertainty) {
        lf_uncertainty<MODEL, L, NV, NS, LANES>(cfg, w, xo, spec, pr, th, sc, sc.hess + (3 * L + 1) * (3 * L + 1));
        unc = sc.hess + (3 * L + 1) * (3 * L + 1);
    }
    lf_outputs<MODEL, L, NV, NS, LANES>(cfg, w, xo, spec, pr, th, unc, out_row, out_sol, out_unc);
    int st = (res.iters & 0xff) | ((res.evals & 0xff) << 8) | res.flags;
    if (w.lane == 0) *out_status = st;
}

// ------------------------------------------------------------------------------------------ uncertainties
// 1-sigma errors exactly as the reference defines them (LuciFit.py:713-722, LuciUtility.py:409-434): finite
// difference Hessian of the nll with delta = 0.1 on ALL 3L+1 unconstrained parameters [A, p, sigma]*L + [c]
// (normalised units), unc = sqrt(|diag(-(H/2)^-1)|).  The 4 (3L+1)^2 objective calls of hessianComp are
// restructured using the additivity of the model over lines: for i, j in different lines
// H_ij = sum D_i D_j / noise^2 with D_i the central difference of the line profile; inside a line the exact
// second-difference expressions are accumulated.  13 unit-profile evaluations per line and sample.
struct LfLineVar { LfLine v[13]; };
// variant order: 0 (p,s) 1 (p+d,s) 2 (p-d,s) 3 (p,s+d) 4 (p,s-d) 5 (p+2d,s) 6 (p-2d,s) 7 (p,s+2d) 8 (p,s-2d)
//                9 (p+d,s+d) 10 (p+d,s-d) 11 (p-d,s+d) 12 (p-d,s-d)
template <int MODEL, int L, int NV, int NS, int LANES>
LF_CORE void lf_uncertainty(const LfCfg& cfg, const LfWarp<LANES>& w, const float* xo, const float* spec, const LfPrep& pr,
                            const float (&th)[LfDims<MODEL, L, NV, NS>::P], const LfScratch& sc, double* unc)
{
    typedef LfDims<MODEL, L, NV, NS> D;
    constexpr int PF = 3 * L + 1;
    constexpr int STRIDE = PF | 1;
    constexpr int NE = PF * (PF + 1) / 2;
    constexpr int SLOTS = (NE + LANES - 1) / LANES;
    const float sinc_w = (float)pr.sinc_w;
    const float dl = 0.1f;
    double A[L], p[L], sg[L], vel[L], beta[L];
    lf_expand<MODEL, L, NV, NS>(cfg, sinc_w, th, A, p, sg, vel, beta);
    // line set-ups of all variants: built lane-parallel into per-warp scratch (tile area is big enough: see host)
    LfLine* lines = reinterpret_cast<LfLine*>(sc.hess);    // 13 L structs (hess is reused afterwards -> copy out first)
    const float dp_[13] = {0, dl, -dl, 0, 0, 2 * dl, -2 * dl, 0, 0, dl, dl, -dl, -dl};
    const float ds_[13] = {0, 0, 0, dl, -dl, 0, 0, 2 * dl, -2 * dl, dl, -dl, dl, -dl};
    for (int e = w.lane; e < 13 * L; e += LANES) {
        int k = e / 13, v = e % 13;
        float pp, pl;
        lf_split(p[k] - cfg.x_ref, pp, pl);
        float s = (float)sg[k] + ds_[v];
        LfLine t;
        lf_line_setup(MODEL, pp, pl - dp_[v], s, sinc_w, t);
        lines[e] = t;
    }
    w.sync();
    const float piw = LF_PI_F / sinc_w, iw = 1.0f / sinc_w;
    const int nfit = cfg.fit_hi - cfg.fit_lo;
    const float* y = spec + cfg.fit_lo;
    const float cont = th[D::IC];
    float gram[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) gram[s] = 0.0f;
    float sAp[L], sAs[L], spp[L], sss[L], sps[L];
#pragma unroll
    for (int k = 0; k < L; ++k) sAp[k] = sAs[k] = spp[k] = sss[k] = sps[k] = 0.0f;
    float* tile = sc.tile;
    for (int n0 = 0; n0 < nfit; n0 += LANES) {
        int n = n0 + w.lane;
        bool live = n < nfit;
        float x = live ? xo[n] : 0.0f;
        float yn = live ? y[n] * pr.inv_wmax : 0.0f;
        float g0[L];
        float m = cont;
#pragma unroll
        for (int k = 0; k < L; ++k) {
            const LfLine& l0 = lines[13 * k];
            float dx = lf_dx(l0, x);
            bool far = false;
            if (MODEL == LF_SINCGAUSS) far = w.all(lf_is_far(l0, dx));
            float gp, gs;
            lf_profile<MODEL>(l0, dx, piw, iw, far, g0[k], gp, gs);
            m += (float)A[k] * g0[k];
        }
        float r0 = yn - m;
        float* trow = tile + w.lane * STRIDE;
#pragma unroll
        for (int k = 0; k < L; ++k) {
            float Ak = (float)A[k];
            float gv[13];
            gv[0] = g0[k];
#pragma unroll
            for (int v = 1; v < 13; ++v) {
                if (MODEL == LF_SINC && ds_[v] != 0.0f) { gv[v] = (dp_[v] == 0.0f) ? g0[k] : 0.0f; continue; }
                const LfLine& lv = lines[13 * k + v];
                float dx = lf_dx(lv, x);
                bool far = false;
                if (MODEL == LF_SINCGAUSS) far = w.all(lf_is_far(lv, dx));
                float gp, gs;
                lf_profile<MODEL>(lv, dx, piw, iw, far, gv[v], gp, gs);
            }
            if (MODEL == LF_SINC) { gv[9] = gv[10] = gv[1]; gv[11] = gv[12] = gv[2]; }
            float Dp = Ak * (gv[1] - gv[2]) * (0.5f / dl);
            float Ds = Ak * (gv[3] - gv[4]) * (0.5f / dl);
            trow[3 * k] = live ? g0[k] : 0.0f;
            trow[3 * k + 1] = live ? Dp : 0.0f;
            trow[3 * k + 2] = live ? Ds : 0.0f;
            if (live) {
                float base = Ak * g0[k];
                // (A, p): D^{s,t} = (A + s d) g(p + t d) - A g0
                {
                    float a = (Ak + dl) * gv[1] - base, b = (Ak + dl) * gv[2] - base;
                    float c = (Ak - dl) * gv[1] - base, d = (Ak - dl) * gv[2] - base;
                    sAp[k] += a * a - b * b - c * c + d * d - 2.0f * r0 * (a - b - c + d);
                }
                {
                    float a = (Ak + dl) * gv[3] - base, b = (Ak + dl) * gv[4] - base;
                    float c = (Ak - dl) * gv[3] - base, d = (Ak - dl) * gv[4] - base;
                    sAs[k] += a * a - b * b - c * c + d * d - 2.0f * r0 * (a - b - c + d);
                }
                {
                    float a = Ak * (gv[5] - g0[k]), b = Ak * (gv[6] - g0[k]);
                    spp[k] += a * a + b * b - 2.0f * r0 * (a + b);
                }
                {
                    float a = Ak * (gv[7] - g0[k]), b = Ak * (gv[8] - g0[k]);
                    sss[k] += a * a + b * b - 2.0f * r0 * (a + b);
                }
                {
                    float a = Ak * (gv[9] - g0[k]), b = Ak * (gv[10] - g0[k]);
                    float c = Ak * (gv[11] - g0[k]), d = Ak * (gv[12] - g0[k]);
                    sps[k] += a * a - b * b - c * c + d * d - 2.0f * r0 * (a - b - c + d);
                }
            }
        }
        trow[3 * L] = live ? 1.0f : 0.0f;
        w.sync();
        // Gram matrix of the D vectors: entry e = (i, j), i <= j, owned by lane e % LANES
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            int e = s * LANES + w.lane;
            if (e < NE) {
                int i = 0, rem = e;
                while (rem >= PF - i) { rem -= PF - i; ++i; }
                int j = i + rem;
                float acc = 0.0f;
                for (int t = 0; t < LANES; ++t) acc += tile[t * STRIDE + i] * tile[t * STRIDE + j];
                gram[s] += acc;
            }
        }
        w.sync();
    }
    // assemble H (double) in scratch
    double noise2 = pr.noise * pr.noise;
    if (!(noise2 == noise2)) noise2 = 1e-2;                       // LuciFit.py:514-515
    double* Hm = sc.hess;
    w.sync();
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        int e = s * LANES + w.lane;
        if (e < NE) {
            int i = 0, rem = e;
            while (rem >= PF - i) { rem -= PF - i; ++i; }
            int j = i + rem;
            double v = (double)gram[s] / noise2;
            Hm[i * PF + j] = v; Hm[j * PF + i] = v;
        }
    }
    w.sync();
    const double f8 = 1.0 / (8.0 * (double)dl * (double)dl * noise2);
#pragma unroll
    for (int k = 0; k < L; ++k) {
        double vAp = (double)w.sum(sAp[k]) * f8, vAs = (double)w.sum(sAs[k]) * f8;
        double vpp = (double)w.sum(spp[k]) * f8, vss = (double)w.sum(sss[k]) * f8, vps = (double)w.sum(sps[k]) * f8;
        if (w.lane == 0) {
            int a = 3 * k, b = 3 * k + 1, c = 3 * k + 2;
            Hm[a * PF + b] = Hm[b * PF + a] = vAp;
            Hm[a * PF + c] = Hm[c * PF + a] = vAs;
            Hm[b * PF + b] = vpp;
            Hm[c * PF + c] = vss;
            Hm[b * PF + c] = Hm[c * PF + b] = vps;
            if (MODEL == LF_SINC) {                                // sigma rows/cols are identically zero -> pinv
                for (int j = 0; j < PF; ++j) { Hm[c * PF + j] = 0.0; Hm[j * PF + c] = 0.0; }
                Hm[c * PF + c] = 1.0;
            }
        }
    }
    w.sync();
    // in-place Gauss-Jordan inverse with row pivoting
    int* piv = reinterpret_cast<int*>(unc + PF);                  // PF ints after the unc vector
    bool singular = false;
    for (int c = 0; c < PF; ++c) {
        double best = 1e300; int bi = 0x7fffffff;
        for (int r = c + w.lane; r < PF; r += LANES) { double a = -fabs(Hm[r * PF + c]); if (a < best) { best = a; bi = r; } }
        w.argmin(best, bi);
        if (!(best < -1e-300)) { singular = true; break; }
        if (w.lane == 0) piv[c] = bi;
        if (bi != c)
            for (int j = w.lane; j < PF; j += LANES) { double t = Hm[c * PF + j]; Hm[c * PF + j] = Hm[bi * PF + j]; Hm[bi * PF + j] = t; }
        w.sync();
        double pv = Hm[c * PF + c];
        w.sync();
        double ip = 1.0 / pv;
        for (int j = w.lane; j < PF; j += LANES) Hm[c * PF + j] = (j == c ? 1.0 : Hm[c * PF + j]) * ip;
        w.sync();
        for (int r = 0; r < PF; ++r) {
            if (r == c) continue;
            double f = Hm[r * PF + c];
            w.sync();
            for (int j = w.lane; j < PF; j += LANES) Hm[r * PF + j] = (j == c ? 0.0 : Hm[r * PF + j]) - f * Hm[c * PF + j];
            w.sync();
        }
    }
    if (!singular) {
        for (int c = PF - 1; c >= 0; --c) {
            int bi = piv[c];
            if (bi != c)
                for (int r = w.lane; r < PF; r += LANES) { double t = Hm[r * PF + c]; Hm[r * PF + c] = Hm[r * PF + bi]; Hm[r * PF + bi] = t; }
            w.sync();
        }
    }
    for (int j = w.lane; j < PF; j += LANES) {
        double u = singular ? 0.0 : sqrt(fabs(2.0 * Hm[j * PF + j]));
        if (MODEL == LF_SINC && j < 3 * L && (j % 3) == 2) u = 0.0;
        unc[j] = u;
    }
    w.sync();
}
