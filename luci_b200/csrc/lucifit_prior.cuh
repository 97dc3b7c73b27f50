// lucifit_prior.cuh -- initial velocity / broadening of a spaxel straight from its spectrum, one warp per spaxel.
//
// Stands in for the reference's CNN prior (Fit.estimate_priors_ML, LUCI/LuciFit.py:335-362: a Keras model that maps the
// interpolated spectrum to (velocity, broadening) in km/s and whose output seeds line_vals_estimate, :382-430).  The
// weights of that network are TensorFlow SavedModels (partly absent from the reference checkout) and TensorFlow is out
// of scope here, so the same two numbers are measured directly:
//   velocity   : matched filter over a velocity grid.  The continuum (clipped mean of the fit window) is removed, the window
//                is smoothed with (1/4, 1/2, 1/4), and score(v) = sum_k max(0, z(p_k(v)))^2 over the expected centres
//                p_k(v) = 1e7 / (lambda_k (1 + v/c)) of all fitted lines (linear interpolation between samples), times
//                a mild weight 1 - 0.1 (v / v_max)^2; the maximum is refined with a parabola through its neighbours;
//   broadening : full width at half maximum of the strongest line at that velocity (walk from the peak sample to the
//                half-maximum crossings, linearly interpolated), with the instrumental sinc width removed in quadrature
//                (FWHM_sinc = 1.20671 w, LuciFunctions.py:12-14) for sinc / sincgauss, sigma = FWHM / (2 sqrt(2 ln 2)),
//                broadening = c sigma / p.
// The fit only needs a start inside the basin of its optimum (about half a channel in velocity); the optimum itself
// does not depend on the prior.  The same source is compiled for the CPU test harness with LANES = 1.
#pragma once
#include "lucifit_core.cuh"

#define LF_PRIOR_MAX_GRID 512

struct LfPriorCfg {
    int model, L, nz, fit_lo, nfit;
    int nv;                                    // velocity grid points (<= LF_PRIOR_MAX_GRID)
    float v_min, v_step;                       // km/s
    float inv_vmax;                            // 1 / max(|v_min|, |v_max|)
    float p0[LF_MAX_LINES];                    // 1e7 / lambda_k                [cm^-1]
    float p0_off[LF_MAX_LINES];                // 1e7 / lambda_k - axis[fit_lo] [cm^-1]
    float chan;                                // mean channel width of the fit window [cm^-1]
    double axis_k;                             // sinc width = axis_k / cos(theta)
    const float* bkg;                          // [nz] or null
};

// z: [nfit] continuum-subtracted window, sm: [nfit] the smoothed window,
// score: [nv] scratch.  Writes (velocity, broadening) in km/s; (0, 0) for spectra the fit will flag anyway.
template <int LANES>
LF_CORE void lf_estimate_prior(const LfPriorCfg& c, const LfWarp<LANES>& w, const float* gspec, float theta_deg,
                               float* z, float* sm, float* score, float* vel_out, float* broad_out)
{
    const int nfit = c.nfit;
    int bad = 0, nvalid = 0;
    float s1 = 0.0f;
    w.sync();
    for (int n = w.lane; n < nfit; n += LANES) {
        float s = gspec[c.fit_lo + n];
        if (c.bkg) s -= c.bkg[c.fit_lo + n];
        if (isinf(s)) { bad = 1; s = 0.0f; }
        z[n] = s;                                      // a NaN sample stays NaN until the window mean is known
        if (s == s) { s1 += s; ++nvalid; }
    }
    bad = w.any(bad != 0) ? 1 : 0;
    nvalid = w.sum(nvalid);
    w.sync();
    if (bad || nvalid < 8) {
        if (w.lane == 0) { *vel_out = 0.0f; *broad_out = 0.0f; }
        return;
    }
    // continuum: mean of the window after three rounds of clipping about the running mean (the lines are a small
    // fraction of the window; 1.5 sigma first -- the spread still contains them --, then 2 sigma twice)
    float mu = w.sum(s1) / (float)nvalid, sd;
    if (nvalid < nfit) {                               // NaN samples (dropped by the fit, LuciBase.py:358-360) sit at the mean
        for (int n = w.lane; n < nfit; n += LANES) if (z[n] != z[n]) z[n] = mu;
        w.sync();
    }
    {
        float s2 = 0.0f;
        for (int n = w.lane; n < nfit; n += LANES) { const float d = z[n] - mu; s2 += d * d; }
        sd = sqrtf(w.sum(s2) / (float)nfit);
    }
    for (int it = 0; it < 3; ++it) {
        const float lim = (it == 0 ? 1.5f : 2.0f) * sd;
        float a1 = 0.0f, a2 = 0.0f;
        int cnt = 0;
        for (int n = w.lane; n < nfit; n += LANES) {
            const float d = z[n] - mu;
            if (fabsf(d) <= lim) { a1 += d; a2 += d * d; ++cnt; }
        }
        a1 = w.sum(a1); a2 = w.sum(a2); cnt = w.sum(cnt);
        if (cnt < 8) break;
        const float m1 = a1 / (float)cnt;
        sd = sqrtf(fmaxf(a2 / (float)cnt - m1 * m1, 0.0f));
        mu += m1;
    }
    w.sync();
    for (int n = w.lane; n < nfit; n += LANES) z[n] -= mu;
    w.sync();
    for (int n = w.lane; n < nfit; n += LANES) {
        const float l = z[n > 0 ? n - 1 : n], r = z[n + 1 < nfit ? n + 1 : n];
        sm[n] = 0.25f * l + 0.5f * z[n] + 0.25f * r;
    }
    w.sync();
    // ---- velocity: matched filter on the grid
    const float ichan = 1.0f / c.chan;
    for (int iv = w.lane; iv < c.nv; iv += LANES) {
        const float vt = (c.v_min + iv * c.v_step) * (1.0f / LF_C_KMS);
        const float shift = vt / (1.0f + vt);                 // p(v) - p0 = -p0 v/c / (1 + v/c)
        float sc = 0.0f;
        for (int k = 0; k < c.L; ++k) {
            const float t = (c.p0_off[k] - c.p0[k] * shift) * ichan;
            const int i = (int)floorf(t);
            if (i < 0 || i + 1 >= nfit) continue;
            const float f = t - (float)i;
            const float val = fmaxf(sm[i] + f * (sm[i + 1] - sm[i]), 0.0f);
            sc += val * val;
        }
        // single-line spectra cannot tell a line from its neighbour in the template (H alpha at v or [NII]6548 at
        // v + 690 km/s score the same up to noise): a mild weight breaks that tie towards the cube's systemic velocity
        const float u = (c.v_min + iv * c.v_step) * c.inv_vmax;
        score[iv] = sc * (1.0f - 0.1f * u * u);
    }
    w.sync();
    double best = -1.0;
    int ibest = 0;
    for (int iv = w.lane; iv < c.nv; iv += LANES) {
        const double s = (double)score[iv];
        if (s > best) { best = s; ibest = iv; }
    }
    { double nb = -best; w.argmin(nb, ibest); best = -nb; }   // largest score, lowest index on ties
    float v = c.v_min + ibest * c.v_step;
    if (ibest > 0 && ibest + 1 < c.nv) {
        const float a = score[ibest - 1], b = score[ibest], d = score[ibest + 1];
        const float den = a - 2.0f * b + d;
        if (den < 0.0f) v += c.v_step * fminf(fmaxf(0.5f * (a - d) / den, -0.5f), 0.5f);
    }
    // ---- broadening: FWHM of the strongest line at that velocity.  Peak: the largest of the +-2 samples around
    // every expected centre (one candidate per lane); half-maximum crossings: the first sample at or below half the
    // peak on either side (one sample per lane and round, lowest lane wins), linearly interpolated.
    const float vt = v * (1.0f / LF_C_KMS);
    const float shift = vt / (1.0f + vt);
    double amp_d = 0.0;                                            // argmin over (-value, candidate index)
    int cand = -1;
    for (int q = w.lane; q < 5 * c.L; q += LANES) {
        const int k = q / 5;
        const float t = (c.p0_off[k] - c.p0[k] * shift) * ichan;
        const int i = (int)floorf(t + 0.5f) + (q - 5 * k) - 2;
        if (i >= 0 && i < nfit && z[i] > 0.0f && (double)z[i] > amp_d) { amp_d = (double)z[i]; cand = q; }
    }
    { double neg = cand >= 0 ? -amp_d : 1.0; int idx = cand >= 0 ? cand : 0x7fffffff; w.argmin(neg, idx); amp_d = -neg; cand = neg < 0.0 ? idx : -1; }
    float broad = 0.0f;
    if (cand >= 0) {                                               // warp-uniform from here on
        const int k = cand / 5;
        const float t = (c.p0_off[k] - c.p0[k] * shift) * ichan;
        const int ipk = (int)floorf(t + 0.5f) + (cand - 5 * k) - 2;
        const float amp = (float)amp_d, half = 0.5f * amp;
        const float pk = c.p0[k] * (1.0f - shift);
        float xl = (float)ipk - 0.5f, xr = (float)ipk + 0.5f;
        for (int j0 = 1; j0 <= 48; j0 += LANES) {                  // left side: samples ipk - j0 - lane
            const int jj = ipk - j0 - w.lane;
            const int hit = w.first_true(jj >= 0 && jj >= ipk - 48 && z[jj] <= half);
            if (hit >= 0) { const int jl = ipk - j0 - hit; xl = (float)jl + (half - z[jl]) / (z[jl + 1] - z[jl]); break; }
            if (ipk - j0 - LANES < 0) break;
        }
        for (int j0 = 1; j0 <= 48; j0 += LANES) {                  // right side: samples ipk + j0 + lane
            const int jj = ipk + j0 + w.lane;
            const int hit = w.first_true(jj < nfit && jj <= ipk + 48 && z[jj] <= half);
            if (hit >= 0) { const int jr = ipk + j0 + hit; xr = (float)jr - (half - z[jr]) / (z[jr - 1] - z[jr]); break; }
            if (ipk + j0 + LANES >= nfit) break;
        }
        const float fwhm = (xr - xl) * c.chan;
        const float ct = fabsf(cosf(theta_deg * 0.017453292519943295f));
        const float fs = (c.model == LF_GAUSSIAN) ? 0.0f : 1.20671f * (float)c.axis_k / fmaxf(ct, 1e-3f);
        const float fg2 = fmaxf(fwhm * fwhm - fs * fs, 0.09f * fwhm * fwhm);
        const float sig = sqrtf(fg2) * (1.0f / 2.3548200450309493f);
        broad = fminf(fmaxf(LF_C_KMS * sig / pk, 1.0f), 1000.0f);
    }
    if (w.lane == 0) { *vel_out = v; *broad_out = broad; }
    w.sync();
}
