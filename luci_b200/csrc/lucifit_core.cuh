// lucifit_core.cuh -- per-spaxel pipeline of the fit path, executed by ONE WARP per spaxel, in four stages that the
// device runs as four small kernels (each hot loop has to fit the 32 KB instruction cache of an SM):
//
//   stage        replaces (LUCI/LuciFit.py unless noted)                                         function
//   prep         Fit.__init__ :49-173 (normalise :110-112, window :225-280, noise :282-331),       lf_stage_prep
//                cont_estimate :432-486, line_vals_estimate :382-430, initial vector :663-672
//   lm           calculate_params :635-687 (SLSQP on the nll with the constraints :523-633)        lf_stage_lm
//   unc          uncertainties :713-722 + LuciUtility.hessianComp :409-434                          lf_stage_unc
//   out          fit() tail :799-834, calc_chisquare :1050-1069, LuciFitParameters.py:14-185        lf_stage_out
//
// Between stages a spaxel is a small state record in the caller's workspace (LF_SO_* offsets below).
//
// The tie constraints of the reference (:523-598) are eliminated analytically: the unknowns are
//   theta = [A_0..A_{L-1} | v_0..v_{NV-1} | beta_0..beta_{NS-1} | continuum]
// with line centre p_k = 1e7/(lambda_k (1 + v_g/c)) and sigma_k = p_k beta_h / c, so equal velocity / equal
// broadening inside a group hold identically, and (if enabled) A_6548 = A_6583 G(sigma_6583)/(3 G(sigma_6548)).
// Line centres are carried as offsets from x_ref = axis[fit_lo] so FP32 never subtracts two 1.5e4-sized numbers.
//
// Lane model: LANES = 32 on the device (sample n of the fit window is owned by lane n % 32), LANES = 1 in the CPU
// test harness (tests/hostemu), which compiles this very file with g++.
//
// Code-size rules of this file: loops over lines are ROLLED (`#pragma unroll 1`) with the per-line constants in
// per-warp shared memory, so there is one copy of the profile code per stage; only the small fixed-size
// normal-equation arrays are unrolled into registers.
#pragma once
#include <stdint.h>
#include "lucifit_math.cuh"

#if defined(__CUDACC__)
#include <cuda_bf16.h>
#endif

// Two floats as one 32-bit word of bfloat16s (round to nearest even): the per-sample partials that only feed the Hessian's
// cross term (lf_eval) travel through shared memory at half the traffic; 3 significant digits are plenty there.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned lf_pack_bf16(float lo, float hi)
{
    const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const unsigned*>(&p);
}
__device__ __forceinline__ void lf_unpack_bf16(unsigned u, float& lo, float& hi)
{
    lo = __uint_as_float(u << 16); hi = __uint_as_float(u & 0xffff0000u);
}
#else
inline unsigned lf_bf16_bits(float f)
{
    unsigned u; memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0u;             // NaN
    return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
inline unsigned lf_pack_bf16(float lo, float hi) { return lf_bf16_bits(lo) | (lf_bf16_bits(hi) << 16); }
inline void lf_unpack_bf16(unsigned u, float& lo, float& hi)
{
    unsigned a = u << 16, b = u & 0xffff0000u;
    memcpy(&lo, &a, 4); memcpy(&hi, &b, 4);
}
#endif

struct alignas(8) LfF2 { float x, y; };
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float lf_bits_float(unsigned u) { return __uint_as_float(u); }
__device__ __forceinline__ unsigned lf_float_bits(float f) { return __float_as_uint(f); }
#else
inline float lf_bits_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned lf_float_bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
#endif

#if defined(__CUDACC__)
#define LF_CORE __device__ __forceinline__
#define LF_ALIGN16 __align__(16)
#else
#define LF_CORE inline
#define LF_ALIGN16 alignas(16)
#endif

#if defined(__CUDACC__)
#define LF_COLD __device__ __noinline__     /* rarely executed: kept out of the hot loops' instruction stream */
#else
#define LF_COLD inline
#endif

#define LF_MAX_LINES 8
#define LF_C_KMS 299792.0f        /* LuciFit.py:26 (integer km/s in the reference) */
#define LF_MAX_CONT_WIN 1024      /* longest continuum window (samples) the clip sorts in per-warp shared memory; longer ones are refused */
#define LF_PREP_MLP 9             /* loads in flight per lane while the prep stage reads a spectrum */
#define LF_REFINE_HALF 6          /* prior-velocity refinement: grid of 2*6+1 points ... */
#define LF_REFINE_STEP 6.0f        /* ... 6 km/s apart (+-36 km/s around the prior) */
#define LF_REFINE_ROUNDS 3         /* the grid follows a maximum on its edge: up to 3 grids */
static_assert(2 * LF_REFINE_HALF + 1 <= 16, "the refinement scores live in the sort scratch (npad >= 16, lf_make_layout)");
#ifndef LF_LM_LAM0
#define LF_LM_LAM0 1e-3f
#define LF_LM_DEC (1.0f / 3.0f)
#define LF_LM_LAMREJ 0.0f
#endif
/* LM acceptance / stopping constants (tuned on tests/golden/parity_sincgauss_cfg2.npz: 1000 reference fits) */
#define LF_TUNE_SLACK_PRED 1e-4   /* below this predicted relative decrease an FP32-noise-sized rise of sum r^2 is accepted */
#ifndef LF_TUNE_NOISE_PRED
#define LF_TUNE_NOISE_PRED 1e-8
#endif
/* LF_TUNE_NOISE_PRED: below this predicted relative decrease (the convergence tolerance itself) the measured gain ratio is the
   FP32 rounding of the residuals (3e-8 of sum r^2 observed) and no longer steers the damping; above it it must -- Gauss-Newton on
   a line group lost in the noise oscillates without the damping the gain ratio asks for (fit_sincgauss_groups[9]) */
#define LF_TUNE_SLACK 4e-6        /* ... of at most this relative size */
#define LF_TUNE_REJ_CONV 1e-5     /* a rejected step predicted to gain less than this means converged */
#ifndef LF_CONV_CONTRACTION
#define LF_CONV_CONTRACTION 0.05f
#endif /* fast contraction: the predicted gain of a step is at most this fraction of the previous step's */
#define LF_CONV_PRED_CAP 1e-5     /* ... and itself below this fraction of sum r^2: only then is the next gain extrapolated */
#ifndef LF_CONV_REMAIN
#define LF_CONV_REMAIN 5e-4f      /* estimated parameter error left after the final (unevaluated) step: half the 1e-3 parity bar */
#endif
#ifndef LF_CONV_FLAT
#define LF_CONV_FLAT 1e-2         /* predicted gain below this fraction of the tolerance: the parameter test is waived */
#endif
#ifndef LF_CONV_AMP_FLOOR
#define LF_CONV_AMP_FLOOR 1e-3f    /* amplitudes (units of the window maximum) below this are compared absolutely against it in the step-size test */
#endif
#ifndef LF_SECANT
#define LF_SECANT 1                /* structured secant term of the Hessian in the LM stage (lf_secant_update) */
#endif
#ifndef LF_CROSS_TERM
#define LF_CROSS_TERM 1
#endif
#ifndef LF_CROSS_GATE
#define LF_CROSS_GATE 1.0          /* the cross term is used once the accepted step predicted less than this fraction of sum r^2 (1: from the second point on) */
#endif
#ifndef LF_FAST_ITERS
#define LF_FAST_ITERS 8            /* accepted iterations after which the first LM kernel hands a spaxel to the second one (see lf_stage_lm) */
#endif
#ifndef LF_SEC_LEARN
#define LF_SEC_LEARN 1e-3          /* ... learnt only from steps predicted to gain less than this fraction of sum r^2 (near the optimum, where the residuals -- and with them S -- have settled) */
#endif
#define LF_CONV_RHO0 0.5f         /* contraction rate assumed when the previous step was damped / rejected (no measurement) */

// status word: bits 0-7 accepted LM iterations, bits 8-15 model evaluations, bits 16.. flags
enum {
    LF_ST_CONVERGED = 1 << 16, LF_ST_MAXITER = 1 << 17, LF_ST_BOUND = 1 << 18, LF_ST_SINGULAR = 1 << 19,
    LF_ST_NAN_INPUT = 1 << 20, LF_ST_MASKED = 1 << 21, LF_ST_BAD_NORM = 1 << 22, LF_ST_NAN_RESULT = 1 << 23,
    LF_ST_NAN_SAMPLES = 1 << 24   /* some samples were NaN and were left out of the fit (LuciBase.py:358-360) */
};
// bits 25-30: number of trailing samples of the fit window the reference's window loses when the sample at its upper
// edge is NaN and the nearest valid sample lies inside (internal to the stage kernels, cleared in the output word)
#define LF_ST_CUT_SHIFT 25
#define LF_ST_CUT_MASK (63 << LF_ST_CUT_SHIFT)
#define LF_ST_BADMASK (LF_ST_NAN_INPUT | LF_ST_MASKED | LF_ST_BAD_NORM)

// per-spaxel state record (float words; the record starts 8-byte aligned and its stride is even)
enum { LF_SO_NOISE = 0, LF_SO_SCALE = 2, LF_SO_CORR = 4, LF_SO_SINCW = 6, LF_SO_INVW = 8, LF_SO_STATUS = 9, LF_SO_LAM = 10, LF_SO_TH = 12 };

struct LfCfg {                     // device-side constants of one lucifit_fit_batch call
    int model, L, nv, ns;
    int nz, fit_lo, fit_hi, noise_lo, noise_hi, cont_lo, cont_hi;
    int vg[LF_MAX_LINES], sg[LF_MAX_LINES];   // dense 0-based group index of each line
    int i48, i83;                  // [NII] tie: indices of NII6548 / NII6583, -1 when off
    int sigbox_group;              // sigma group of the LAST line: 0 <= sigma <= 10 cm^-1 (LuciFit.py:539-541)
    int uncertainty, ml_scale, max_iter, n_out;
    int freeze;                    // 1: velocity / broadening columns pinned at the priors, no ties, no boxes (LuciFit.py:688-709)
    int prior_groups;              // 1: priors are per group ([nv] velocities, [ns] broadenings per spaxel), 0: one pair
    int state_floats;              // stride of the state record
    int refine_prior;              // 1: matched-filter refinement of the prior velocity in the prep stage
    float ichan;                   // 1 / mean channel width of the fit window [1/cm^-1]
    float ftol;
    float p0[LF_MAX_LINES];        // 1e7/lambda_k                     [cm^-1]
    float p0_off[LF_MAX_LINES];    // 1e7/lambda_k - x_ref             [cm^-1]
    float p0_lo[LF_MAX_LINES];     // FP32 remainder of the above
    double line_nm[LF_MAX_LINES];
    double x_ref;
    double axis_k;                 // 1e7/(2 STEP (n_steps - zpd)); axis_step = sinc_width = corr * axis_k
    const double* axis;            // [nz]  float32-rounded axis values, widened; strictly increasing
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
    int first_true(bool p) const { return p ? 0 : -1; }      // lowest lane whose predicate holds, -1 if none
    float xchg(float v, int) const { return v; }
    unsigned or_all(unsigned v) const { return v; }
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
    __device__ int first_true(bool p) const { return __ffs((int)__ballot_sync(0xffffffffu, p)) - 1; }
    __device__ float xchg(float v, int o) const { return __shfl_xor_sync(0xffffffffu, v, o); }
    __device__ unsigned or_all(unsigned v) const { return __reduce_or_sync(0xffffffffu, v); }
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

// Which branch of w(z) the lanes of a chunk take for one sincgauss line (see lf_profile_sc).
template <int LANES>
LF_CORE int lf_region_hint(const LfWarp<LANES>& w, const LfLine& ln, float dx)
{
    const float b = dx * ln.is2s;
    const float bb = b * b, r2 = bb + ln.a * ln.a;
    if (w.all(bb >= 40.0f)) return LF_HINT_FAR;
    if (ln.a >= LF_NEAR_A_MIN && w.any(r2 < LF_W_R2_CF) && w.all(r2 < LF_NEAR_R2_MAX)) return LF_HINT_NEAR;
    return LF_HINT_LANE;
}

// ------------------------------------------------------------------------------------------- small helpers
template <int P> struct LfSym {            // packed upper triangle, row-major
    static LF_HD constexpr int idx(int i, int j) { return i * P - (i * (i - 1)) / 2 + (j - i); }
    enum { N = P * (P + 1) / 2 };
};

template <int MODEL, int L, int NV, int NS> struct LfDims {
    enum { P = L + NV + NS + 1, IV = L, IS = L + NV, IC = L + NV + NS, NH = P * (P + 1) / 2, NA = NH + P,
           // exact second-derivative term of the (amplitude, velocity / broadening) block of the Hessian (see lf_eval): kept
           // for the single-group fits; with several groups (double components) the exact Hessian is indefinite away from
           // the optimum and plain Gauss-Newton is the safer model
           CROSS = (LF_CROSS_TERM && NV == 1 && NS == 1) ? 1 : 0, NAX = NA + 2 * L * CROSS };
};

// Index of the axis sample nearest to `pos` (first one on ties, like np.argmin(|axis - pos|)) for a strictly
// increasing axis: binary search for the first sample >= pos, then the closer of it and its left neighbour.
LF_CORE int lf_nearest_axis(const double* axis, int nz, double pos)
{
    int lo = 0, hi = nz;                                // first index with axis[i] >= pos, in [0, nz]
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (axis[mid] < pos) lo = mid + 1; else hi = mid;
    }
    if (lo == 0) return 0;
    if (lo == nz) return nz - 1;
    return (fabs(axis[lo - 1] - pos) <= fabs(axis[lo] - pos)) ? lo - 1 : lo;
}

// The same index found from a guess: the axis is close to uniform, so `guess` = fit_lo + (pos - x_ref) / (mean channel
// width) is within a sample or two and a short walk replaces the ten dependent loads of the bisection.
LF_CORE int lf_nearest_axis_from(const double* axis, int nz, double pos, int guess)
{
    int lo = guess < 0 ? 0 : (guess > nz ? nz : guess);        // -> first index with axis[i] >= pos, in [0, nz]
    while (lo < nz && axis[lo] < pos) ++lo;
    while (lo > 0 && axis[lo - 1] >= pos) --lo;
    if (lo == 0) return 0;
    if (lo == nz) return nz - 1;
    return (fabs(axis[lo - 1] - pos) <= fabs(axis[lo] - pos)) ? lo - 1 : lo;
}

LF_CORE int lf_axis_guess(const LfCfg& cfg, double pos)
{
    const double g = (double)cfg.fit_lo + (pos - cfg.x_ref) * (double)cfg.ichan;
    return (int)fmin(fmax(g, 0.0), (double)cfg.nz);
}

// Where the reference's np.argmin(|axis - target|) lands once the NaN samples have been removed from spectrum and axis
// (LuciBase.py:358-360): the valid sample nearest to `target`, given the nearest sample `idx` of the full axis.
template <typename IsNan>
LF_CORE int lf_valid_near(const double* axis, int nz, int idx, double target, IsNan is_nan)
{
    if (!is_nan(idx)) return idx;
    int l = idx - 1, r = idx + 1;
    while (l >= 0 && is_nan(l)) --l;
    while (r < nz && is_nan(r)) ++r;
    if (l < 0) return r < nz ? r : idx;
    if (r >= nz) return l;
    return (fabs(axis[l] - target) <= fabs(axis[r] - target)) ? l : r;
}

// In-place ascending bitonic sort of v[0..npad) (npad a power of two; pad with +inf), all lanes cooperate.
template <int LANES>
LF_CORE void lf_sort(const LfWarp<LANES>& w, float* v, int npad)
{
#if defined(__CUDA_ARCH__)
    if (LANES == 32 && npad <= 64) {
        // register form for windows of up to 64 samples: element e lives in lane e & 31, register e >> 5; partners
        // closer than 32 are reached with one shuffle, the distance-32 step is the lane's own pair
        float r0 = w.lane < npad ? v[w.lane] : INFINITY, r1 = w.lane + 32 < npad ? v[w.lane + 32] : INFINITY;
#pragma unroll
        for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                if (j == 32) {
                    const float lo = fminf(r0, r1), hi = fmaxf(r0, r1);
                    r0 = lo; r1 = hi;
                } else {
                    const bool asc = k == 64 || (w.lane & k) == 0;
                    const bool keep_min = ((w.lane & j) == 0) == asc;
                    const float p0 = __shfl_xor_sync(0xffffffffu, r0, j), p1 = __shfl_xor_sync(0xffffffffu, r1, j);
                    r0 = keep_min ? fminf(r0, p0) : fmaxf(r0, p0);
                    r1 = keep_min ? fminf(r1, p1) : fmaxf(r1, p1);
                }
            }
        }
        w.sync();
        if (w.lane < npad) v[w.lane] = r0;
        if (w.lane + 32 < npad) v[w.lane + 32] = r1;
        w.sync();
        return;
    }
#endif
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = w.lane; t < (npad >> 1); t += LANES) {
                int i = 2 * (t & ~(j - 1)) + (t & (j - 1));       // j is a power of two: no division
                int l = i + j;
                bool up = (i & k) == 0;
                float a = v[i], b = v[l];
                if ((a > b) == up) { v[i] = b; v[l] = a; }
            }
            w.sync();
        }
    }
}

LF_CORE double lf_sorted_median(const float* v, int a, int m)   // median of the sorted range v[a .. a+m), float64
{
    if (m <= 0) return 0.0;
    if (m & 1) return (double)v[a + m / 2];
    return 0.5 * ((double)v[a + m / 2 - 1] + (double)v[a + m / 2]);
}

// k-th smallest (0-based) of |v[a+i] - med|, i in [0, m), for a SORTED range v: the deviations left of the median
// (descending index) and right of it are two ascending lists; every element finds its rank in the union with one
// binary search in the other list, the element of rank k is the answer.  No second sort.
template <int LANES>
LF_CORE double lf_dev_kth(const LfWarp<LANES>& w, const float* v, int a, int m, int i0, double med, int k)
{
    // left list:  dL(p) = med - v[a + i0 - 1 - p], p in [0, i0)      right list: dR(q) = v[a + i0 + q] - med, q in [0, m - i0)
    const int nL = i0, nR = m - i0;
    double found = 0.0;
    int hit = 0;
    for (int i = w.lane; i < m; i += LANES) {
        const bool left = i < i0;
        const double d = left ? med - (double)v[a + i] : (double)v[a + i] - med;
        int lo = 0, hi = left ? nR : nL;                            // count of the other list's elements before d
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const double o = left ? (double)v[a + i0 + mid] - med : med - (double)v[a + i0 - 1 - mid];
            const bool before = left ? (o < d) : (o <= d);          // ties: left elements first (a strict total order)
            if (before) lo = mid + 1; else hi = mid;
        }
        const int rank = (left ? i0 - 1 - i : i - i0) + lo;
        if (rank == k) { found = d; hit = 1; }
    }
    return w.sum(hit ? found : 0.0);                                // exactly one lane hits
}

// Continuum guess: median of the window after <=3 rounds of 1-sigma clipping about the median with mad_std as the
// spread (LuciFit.py:474-486; astropy sigma_clip semantics restated in oracle/astro_stats.py).  The window is
// sorted once (float keys are exact: the data are float32); a clip keeps the values inside an interval, so the
// survivors are always a contiguous range of the sorted window, and the MAD comes from rank counting in it.
// `sv` is per-warp scratch of `npad` floats.
template <int LANES>
LF_CORE double lf_cont_estimate(const LfWarp<LANES>& w, const float* spec, int lo, int hi, double sigma_level,
                                float* sv, int npad)
{
    int n = hi - lo;
    if (n > npad) n = npad;                            // lf_make_plan sizes npad for the whole window (<= LF_MAX_CONT_WIN)
    if (n <= 0) return 0.0;
    w.sync();
    int n_nan = 0;
    for (int i = w.lane; i < npad; i += LANES) {
        float v = i < n ? spec[lo + i] : INFINITY;
        if (v != v) { v = INFINITY; ++n_nan; }         // NaN samples are not part of the spectrum (LuciBase.py:358-360)
        sv[i] = v;
    }
    n -= w.sum(n_nan);
    w.sync();
    if (n <= 0) return 0.0;
    lf_sort(w, sv, npad);
    int a = 0, m = n;
#pragma unroll 1
    for (int it = 0; it < 3 && m > 0; ++it) {
        const double med = lf_sorted_median(sv, a, m);
        int i0 = 0;
        for (int i = w.lane; i < m; i += LANES) i0 += ((double)sv[a + i] < med);
        i0 = w.sum(i0);
        double dmed = lf_dev_kth(w, sv, a, m, i0, med, m / 2);
        if (!(m & 1)) dmed = 0.5 * (lf_dev_kth(w, sv, a, m, i0, med, m / 2 - 1) + dmed);
        const double mad = 1.482602218505602 * dmed;
        const double lo_b = med - sigma_level * mad, hi_b = med + sigma_level * mad;
        int below = 0, above = 0;
        for (int i = w.lane; i < m; i += LANES) { double x = (double)sv[a + i]; below += !(x >= lo_b); above += !(x <= hi_b); }
        below = w.sum(below); above = w.sum(above);
        if (below + above == 0) break;
        a += below; m -= below + above;
    }
    if (m < 1) { a = 0; m = n; }                       // LuciFit.py:481-482: fall back to the unclipped window
    return lf_sorted_median(sv, a, m);
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

// ------------------------------------------------------------------------------------- per-warp work areas
// One model evaluation's per-line working set, kept in per-warp shared memory (128 bytes).
struct LF_ALIGN16 LfLineX {
    LfLine ln;
    float aeff;        // amplitude in effect (tied [NII]6548: R * A_6583)
    float dpdv;        // d centre / d v_g
    float dsdv;        // d sigma  / d v_g
    float dsdb;        // d sigma  / d beta_h
    float ck, sk;      // cos, sin of the line's phase pi (centre - x_ref)/w
    int vg, sg;        // group indices
    float p, beta;
    // sincgauss wing path (all lanes of a chunk beyond |z|^2 = 300, exp(-b^2) = 0): every per-line factor folded
    float a2;          // a^2
    float neae;        // -exp(-a^2)/erf(a):                      g    = neae T
    float c_m;         // A neae:                                 model += c_m T
    float c_q;         // 2a + q,  U = c_q T - c_ia2 r^2 T2       (dg/dalpha = -neae U)
    float c_ia2;       // 2/a
    float c_v1;        // -A neae dpdv / (sqrt2 sigma):           dm/dv    = c_v1 Tb + c_v2 U
    float c_v2;        // -A neae a dsdv / sigma
    float c_b2;        // -A neae a dsdb / sigma:                 dm/dbeta = c_b2 U
    // area factor of the [NII] tie (LuciFit.py:584-595), per line so that all lines form it at once (lane k = line k)
    float Gk;          // sincgauss: G = sigma / erf(u), u = sigma / (sqrt2 w); gaussian, sinc: sigma
    float Ek;          // 1/sigma - d ln G / d sigma = (2/sqrt pi) exp(-u^2) / (erf(u) sqrt2 w)   (0 for gaussian, sinc)
};

// Folds the amplitude in effect into the wing-path factors of a line record (after aeff is final).
LF_CORE void lf_fold_wing(LfLineX& x)
{
    const LfLine& l = x.ln;
    const float eae = l.ea * l.ierf;
    x.a2 = l.a * l.a;
    x.neae = -eae;
    x.c_m = -x.aeff * eae;
    x.c_q = 2.0f * l.a + l.q;
    x.c_ia2 = 2.0f * l.ia;
    x.c_v1 = x.aeff * eae * l.is2s * x.dpdv;
    x.c_v2 = x.aeff * eae * l.a * l.isig * x.dsdv;
    x.c_b2 = x.aeff * eae * l.a * l.isig * x.dsdb;
}
struct LF_ALIGN16 LfLineV {        // compact record of one shifted variant of a line (uncertainty stage), 64 bytes
    LfLine ln;
    float ck, sk;      // cos, sin of the variant's phase
    float a2, neae;    // a^2, -exp(-a^2)/erf(a)  (wing form)
    float aeff, pad;
};
struct LfTieX {        // warp-uniform constants of the [NII] tie for one evaluation
    float R;
    float cRv[LF_MAX_LINES];
    float cRb[LF_MAX_LINES];
    int active;
};

struct LfWork {        // pointers into the per-warp scratch (which of them are valid depends on the stage)
    float* spec;       // prep: [nz] background-subtracted spectrum
    float* sortv;      // prep: [npad] sorted continuum window
    int npad;
    float* yw;         // lm/unc/out: [nfit] data of the fit window (normalised as the stage needs)
    float* sw;         // [nfit] sin(pi xo/w)
    float* cw;         // [nfit] cos(pi xo/w)
    LfLineX* lines;    // [L]
    LfLineV* vlines;   // unc: [13 L] shifted variants
    LfTieX* tie;
    float* gt;         // [(L+1)][LANES] unit profiles of the current chunk (unc: [13][LANES])
    float* hb;         // lm: [2][NA] normal equations (accepted / trial)
    float* thb;        // lm: [4][P]  theta, trial theta, lower, upper
    double* apd;       // unc/out: [3][L] amplitude, centre, sigma of every line (float64)
    float* tile;       // unc: [LANES][(3L+1)|1]
    double* hess;      // unc: [PF*PF + 2 PF]
};

// Fills sw/cw (phase table) for the window: the angle of sample n is pi xo[n]/w.
template <int LANES>
LF_CORE void lf_phase_table(const LfWarp<LANES>& w, const float* xo, int nfit, double inv_w, float* sw, float* cw)
{
    for (int n = w.lane; n < nfit; n += LANES) {
        float s, c;
        lf_phase((double)xo[n] * inv_w, s, c);
        sw[n] = s; cw[n] = c;
    }
}

// Per-line records of one evaluation at parameters th (per-warp memory), lane k builds line k.
template <int MODEL, int L, int NV, int NS, int LANES>
LF_CORE void lf_setup_lines(const LfCfg& cfg, const LfWarp<LANES>& w, const float* th, float sinc_w, double inv_w,
                            LfLineX* sl, LfTieX* tx)
{
    typedef LfDims<MODEL, L, NV, NS> D;
    w.sync();
    for (int k = w.lane; k < L; k += LANES) {
        int g = NV == 1 ? 0 : cfg.vg[k], h = NS == 1 ? 0 : cfg.sg[k];
        float v = th[D::IV + g], beta = th[D::IS + h];
        float vt = v * (1.0f / LF_C_KMS);
        float i1v = LF_SETUP_RCP(1.0f + vt);
        float p = cfg.p0[k] * i1v;
        float sig = p * beta * (1.0f / LF_C_KMS);
        LfLineX x;
        lf_line_setup(MODEL, cfg.p0_off[k], cfg.p0_lo[k] - cfg.p0[k] * vt * i1v, sig, sinc_w, x.ln);
        x.p = p; x.beta = beta;
        x.dpdv = -p * i1v * (1.0f / LF_C_KMS);
        x.dsdv = beta * (1.0f / LF_C_KMS) * x.dpdv;
        x.dsdb = p * (1.0f / LF_C_KMS);
        x.aeff = th[k];
        x.vg = g; x.sg = h;
        x.ck = 1.0f; x.sk = 0.0f;
        if (MODEL != LF_GAUSSIAN) lf_phase(((double)x.ln.pp + (double)x.ln.dp) * inv_w, x.sk, x.ck);
        lf_fold_wing(x);
        x.Gk = x.ln.sig; x.Ek = 0.0f;
        if (MODEL == LF_SINCGAUSS && cfg.i48 >= 0) {
            const float iw2 = LF_SETUP_RCP(LF_SQRT2_F * sinc_w);
            const float u = x.ln.sig * iw2;
            const float ierf_u = LF_SETUP_RCP(erff(u));
            x.Gk = x.ln.sig * ierf_u;
            x.Ek = LF_2ISQRTPI_F * LF_SETUP_EXP(-u * u) * ierf_u * iw2;
        }
        sl[k] = x;
    }
    w.sync();
    if (w.lane == 0) {
        bool tie = cfg.i48 >= 0;
        if (tie && MODEL == LF_SINC) {
            // sinc: A_83 sigma_83/3 = A_48 sigma_48 (LuciFit.py:587-589) is vacuous when sigma = 0 (no prior
            // broadening, the reference's ML_bool=False path) -> both amplitudes are free
            if (sl[cfg.i83].beta == 0.0f || sl[cfg.i48].beta == 0.0f) tie = false;
        }
        tx->active = tie ? 1 : 0;
        tx->R = 0.0f;
#pragma unroll
        for (int g = 0; g < NV; ++g) tx->cRv[g] = 0.0f;
#pragma unroll
        for (int h = 0; h < NS; ++h) tx->cRb[h] = 0.0f;
        if (tie) {
            // A_6548 = R A_6583, R = G(sigma_6583)/(3 G(sigma_6548)); ln G = ln p + ln beta - ln c [- ln erf(u)]
            const LfLineX& x83 = sl[cfg.i83];
            const LfLineX& x48 = sl[cfg.i48];
            float p83 = x83.p, p48 = x48.p, b83 = x83.beta, b48 = x48.beta, s83 = x83.ln.sig, s48 = x48.ln.sig;
            float a83 = x83.aeff;
            int g83 = x83.vg, g48 = x48.vg, h83 = x83.sg, h48 = x48.sg;
            float R, E83 = 0.f, E48 = 0.f;
            if (MODEL == LF_SINCGAUSS) { R = x83.Gk * LF_SETUP_RCP(3.0f * x48.Gk); E83 = x83.Ek; E48 = x48.Ek; }
            else R = lf_tie_ratio(MODEL, p83, b83, s83, p48, b48, s48, h83 == h48, sinc_w);
#pragma unroll
            for (int g = 0; g < NV; ++g)
                tx->cRv[g] = a83 * R * ((g == g83 ? x83.dpdv * LF_SETUP_RCP(p83) - E83 * x83.dsdv : 0.f) - (g == g48 ? x48.dpdv * LF_SETUP_RCP(p48) - E48 * x48.dsdv : 0.f));
#pragma unroll
            for (int h = 0; h < NS; ++h) {
                float t = 0.f;
                if (h83 != h48) t = (h == h83 ? 1.0f / fmaxf(b83, 1e-6f) : 0.f) - (h == h48 ? 1.0f / fmaxf(b48, 1e-6f) : 0.f);
                tx->cRb[h] = a83 * R * (t - (h == h83 ? E83 * x83.dsdb : 0.f) + (h == h48 ? E48 * x48.dsdb : 0.f));
            }
            tx->R = R;
            sl[cfg.i48].aeff = R * a83;
            lf_fold_wing(sl[cfg.i48]);
        }
    }
    w.sync();
}

// Sum over lanes of acc[0..N): recursive halving, afterwards every element lives (fully summed) in exactly one
// lane, which stores it to dst[element].  ~4 instructions per element and step instead of a 5-step butterfly
// on all N values.
template <int N, int OFF, int LANES> struct LfReduceScatter {
    static LF_CORE void run(const LfWarp<LANES>& w, float* v, int base, int size, float* dst)
    {
        constexpr int H = (N + 1) / 2;
        const bool up = (w.lane & OFF) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            float lo = v[i], hi = (H + i < N) ? v[H + i] : 0.0f;
            float keep = up ? hi : lo, send = up ? lo : hi;
            v[i] = keep + w.xchg(send, OFF);
        }
        LfReduceScatter<H, OFF / 2, LANES>::run(w, v, base + (up ? H : 0), up ? size - H : (size < H ? size : H), dst);
    }
};
template <int N, int LANES> struct LfReduceScatter<N, 0, LANES> {
    static LF_CORE void run(const LfWarp<LANES>&, float* v, int base, int size, float* dst)
    {
#pragma unroll
        for (int i = 0; i < N; ++i) if (i < size) dst[base + i] = v[i];
    }
};

// Wing of the sincgauss profile (|z|^2 >= LF_W_R2_ASY, exp(-b^2) = 0): T = Re[e^{-i phi} w(z)], z = -b + i a, from the
// three-term asymptotic series w = (i/sqrt pi) (1/z) (1 + 1/(2 z^2) + 3/(4 z^4)), split as w = (i/sqrt pi)/z + w2
// (T2 is the w2 part); g = -exp(-a^2)/erf(a) T.  With 1/z = conj(z)/|z|^2 every power of 1/z is a power of conj(z)
// times a power of inv = 1/|z|^2, so all the complex products are formed from b, a, s, c BEFORE the reciprocal is
// needed (they overlap the MUFU latency) and only a short chain depends on it:
//   A1 = c a - s b,  A2 = s a + c b                (e^{-i phi} conj(z): T1 = inv A1 / sqrt pi)
//   P = conj(z)^2 = (b^2 - a^2, 2ab),  Q = P^2,    B1 = (P.A)/2,  B2 = 3 (Q.A)/4   with X.A = Xr A1 + Xi A2
//   T2 = inv^3 (B1 + inv^2 B2) / sqrt pi,  T = inv A1 / sqrt pi + T2,  E2 = inv A2 / sqrt pi
// E2 serves the b-derivative: (1/sqrt pi)(c Re(1/z) + s Im(1/z)) = -E2.
LF_CORE void lf_wing_terms(float b, float a, float bb, float a2, float r2, float s, float c, float& T, float& T2, float& E2)
{
    const float inv = LF_RCP(r2);
    const float A1 = c * a - s * b, A2 = s * a + c * b;
    const float Pr = bb - a2, Pi = 2.0f * a * b;
    const float Qr = Pr * Pr - Pi * Pi, Qi = 2.0f * Pr * Pi;
    const float B1 = 0.5f * (Pr * A1 + Pi * A2), B2 = 0.75f * (Qr * A1 + Qi * A2);
    const float v = inv * inv, iv = LF_ISQRTPI_F * inv;
    T2 = (iv * v) * (B1 + v * B2);
    T = iv * A1 + T2;
    E2 = iv * A2;
}

// Complex w(z), z = -b + i a, from the same three-term asymptotic series (|z|^2 >= LF_W_R2_ASY), and its first two
// Taylor coefficients from the exact identities w' = -2 z w + 2i/sqrt pi, w'' = -2 w - 2 z w' (the cancellation in them
// costs ~1e-7 absolute, which a step |dz| <= 0.2 scales down further).  Used by the uncertainty stage: the variants of a
// line that share a sigma differ, in the wing, by a real step dz = -dp / (sqrt2 sigma) of at most 0.5 % of |z|, so they
// are w + dz (w' + dz w''/2) -- third-order error 2e-7 relative -- with each variant's own, exact, phase rotation.
LF_CORE void lf_wing_w(float b, float a, float r2, float& wr, float& wi)
{
    const float inv = LF_RCP(r2);
    const float ur = -b * inv, ui = -a * inv;                       // 1/z
    const float sr = ur * ur - ui * ui, si = 2.0f * ur * ui;        // 1/z^2
    const float tr = 0.5f + 0.75f * sr, ti = 0.75f * si;
    const float pr = 1.0f + (sr * tr - si * ti), pi = sr * ti + si * tr;
    const float qr = ur * pr - ui * pi, qi = ur * pi + ui * pr;
    wr = -LF_ISQRTPI_F * qi; wi = LF_ISQRTPI_F * qr;
}
LF_CORE void lf_wing_w_taylor(float b, float a, float wr, float wi, float& d1r, float& d1i, float& d2r, float& d2i)
{
    const float x = -b, y = a;
    d1r = -2.0f * (x * wr - y * wi);
    d1i = LF_2ISQRTPI_F - 2.0f * (x * wi + y * wr);
    d2r = -wr - (x * d1r - y * d1i);                                // w''/2
    d2i = -wi - (x * d1i + y * d1r);
}

// Gauss-Newton drops -sum r grad^2 m from the Hessian of 1/2 sum r^2.  Its (amplitude, velocity) and (amplitude,
// broadening) entries are known exactly at no extra profile work -- d2m / dA_k dv = d g_k / dv, the unit-profile partial the
// Jacobian is built from -- and they are what couples the weak lines' amplitudes to the common velocity / width on noisy
// spectra: with them in the model the iteration needs 10 % fewer evaluations on the benchmark cube (E 4.45 -> 3.98; 7.0 ->
// 3.1 on the sinc strip of config 1) and ends closer to the optimum (worst flux deviation of the held-out sample 7e-4 ->
// 2.4e-4).  With the velocity entries alone the evaluation is cheaper, but then the broadening stays the one slowly
// contracting direction and the stopping rule, which measures ONE contraction rate, stops a weak line 2e-3 short of its
// flux.  hg[NA + k] = sum r dg_k/dv, hg[NA + L + k] = sum r dg_k/dbeta (lf_eval, which has already moved the tied
// [NII]6548 line's sums onto the 6583 column with weight R).  sign = -1 puts the term into the packed matrix, +1 takes it
// out again (the solver does that when the corrected matrix is not positive definite, and proposes a plain Gauss-Newton
// step instead).
template <int MODEL, int L, int NV, int NS, int LANES>
LF_CORE void lf_cross_apply(const LfWarp<LANES>& w, float* hg, float sign)
{
    typedef LfDims<MODEL, L, NV, NS> D;
    constexpr int P = D::P, NA = D::NA;
    w.sync();
    for (int k = w.lane; k < L; k += LANES) {
        hg[LfSym<P>::idx(k, D::IV)] += sign * hg[NA + k];
        if (MODEL != LF_SINC) hg[LfSym<P>::idx(k, D::IS)] += sign * hg[NA + L + k];
    }
    w.sync();
}

// One pass over the fit window: chi2 = sum r^2, and (into dst[NA], per-warp memory) H = J^T J packed upper
// followed by g = J^T r, at the parameters the line records `sl` were built for.
template <int MODEL, int L, int NV, int NS, int LANES>
LF_CORE void lf_eval(const LfCfg& cfg, const LfWarp<LANES>& w, const float* xo, const LfWork& wk, int nfit, float sinc_w,
                     float cont, double& chi2_out, float* dst)
{
    typedef LfDims<MODEL, L, NV, NS> D;
    constexpr int P = D::P, NH = D::NH, NA = D::NA, NAX = D::NAX;
    constexpr bool CROSS = D::CROSS != 0;
    const LfLineX* sl = wk.lines;
    // per-chunk tile [L][LANES]: the unit profile of every line at the lane's sample and, with CROSS, its velocity /
    // broadening partials as a bf16 pair in the same 8-byte slot (one 64-bit store per line, one 64-bit load back; three
    // significant digits are plenty for a term that only shapes the step)
    constexpr int GS = CROSS ? 2 : 1;
    float* gt = wk.gt;
    const float iw = 1.0f / sinc_w;
    float acc[NAX];
#pragma unroll
    for (int e = 0; e < NAX; ++e) acc[e] = 0.0f;
    double chi2 = 0.0;
    const bool tie = wk.tie->active != 0;
    for (int n0 = 0; n0 < nfit; n0 += LANES) {
        const int n = n0 + w.lane;
        const bool inwin = n < nfit;
        const float x = inwin ? xo[n] : 0.0f;
        const float yn = inwin ? wk.yw[n] : NAN;
        const bool live = yn == yn;                       // NaN samples are not part of the fit (LuciBase.py:358-360)
        float S = 0.0f, C = 1.0f;
        if (MODEL != LF_GAUSSIAN && inwin) { S = wk.sw[n]; C = wk.cw[n]; }
        float Jv[NV], Js[NS];
#pragma unroll
        for (int g = 0; g < NV; ++g) Jv[g] = 0.0f;
#pragma unroll
        for (int h = 0; h < NS; ++h) Js[h] = 0.0f;
        float m = cont;
#pragma unroll 1
        for (int k = 0; k < L; ++k) {
            const LfLineX& X = sl[k];
            const float dx = lf_dx(X.ln, x);
            if (MODEL == LF_GAUSSIAN) {
                float b = dx * X.ln.is2s;
                if (w.all(b * b > 100.0f)) {                                              // exp(-100) = 0 in FP32
                    if (CROSS) reinterpret_cast<LfF2*>(gt)[k * LANES + w.lane] = LfF2{0.0f, 0.0f};
                    else gt[k * LANES + w.lane] = 0.0f;
                    continue;
                }
            }
            float s = 0.0f, c = 1.0f;
            if (MODEL != LF_GAUSSIAN) { s = S * X.ck - C * X.sk; c = C * X.ck + S * X.sk; }
            float g, tv, tb;
            bool wing = false;
            if (MODEL == LF_SINCGAUSS) {
                const float b = dx * X.ln.is2s;
                const float r2 = b * b + X.a2;
                wing = w.all(b * b >= 40.0f && r2 >= LF_W_R2_ASY);
                if (wing) {
                    // the whole chunk is in the asymptotic region of w(z) and exp(-b^2) = 0: the wing formulas of
                    // lf_profile_sc with every per-line factor pre-multiplied (lf_fold_wing)
                    float T, T2, E2;
                    lf_wing_terms(b, X.ln.a, b * b, X.a2, r2, s, c, T, T2, E2);
                    const float Tb = -(2.0f * X.ln.a) * E2 - 2.0f * b * T2;
                    const float U = X.c_q * T - (X.c_ia2 * r2) * T2;
                    g = X.neae * T;
                    m += X.c_m * T;
                    tv = Tb * X.c_v1 + U * X.c_v2;
                    tb = U * X.c_b2;
                }
            }
            if (!wing) {
                float gp, gs;
                int hint = LF_HINT_LANE;
                if (MODEL == LF_SINCGAUSS) hint = lf_region_hint(w, X.ln, dx);
                lf_profile_sc<MODEL>(X.ln, dx, s, c, iw, hint, g, gp, gs);
                const float A = X.aeff;
                m += A * g;
                const float tp = A * gp, ts = A * gs;
                tv = tp * X.dpdv + ts * X.dsdv;
                tb = ts * X.dsdb;
            }
            // tv, tb: this line's share of dm/dv, dm/dbeta (amplitude included; the cross sums are divided by it at the end)
            if (CROSS) reinterpret_cast<LfF2*>(gt)[k * LANES + w.lane] = LfF2{g, lf_bits_float(lf_pack_bf16(tv, tb))};
            else gt[k * LANES + w.lane] = g;
#pragma unroll
            for (int gg = 0; gg < NV; ++gg) Jv[gg] += (NV == 1 || X.vg == gg) ? tv : 0.0f;
            if (MODEL != LF_SINC) {
#pragma unroll
                for (int hh = 0; hh < NS; ++hh) Js[hh] += (NS == 1 || X.sg == hh) ? tb : 0.0f;
            }
        }
        float J[P];
        float ub[CROSS ? L : 1];                           // the packed partials of every line, until the residual is known
        if (tie) {
            // amplitude columns: dm/dA_83 = g_83 + R g_48, the 6548 column is dead (fixed up in the lane's own tile
            // slots so that J keeps compile-time register indices)
            const float g48v = gt[(cfg.i48 * LANES + w.lane) * GS];
            gt[(cfg.i83 * LANES + w.lane) * GS] += wk.tie->R * g48v;
            gt[(cfg.i48 * LANES + w.lane) * GS] = 0.0f;
#pragma unroll
            for (int gg = 0; gg < NV; ++gg) Jv[gg] += g48v * wk.tie->cRv[gg];
            if (MODEL != LF_SINC) {
#pragma unroll
                for (int hh = 0; hh < NS; ++hh) Js[hh] += g48v * wk.tie->cRb[hh];
            }
        }
        if (CROSS) {
#pragma unroll
            for (int k = 0; k < L; ++k) { const LfF2 e = reinterpret_cast<const LfF2*>(gt)[k * LANES + w.lane]; J[k] = e.x; ub[k] = e.y; }
        } else {
#pragma unroll
            for (int k = 0; k < L; ++k) J[k] = gt[k * LANES + w.lane];
        }
#pragma unroll
        for (int gg = 0; gg < NV; ++gg) J[D::IV + gg] = Jv[gg];
#pragma unroll
        for (int hh = 0; hh < NS; ++hh) J[D::IS + hh] = (MODEL != LF_SINC) ? Js[hh] : 0.0f;
        J[D::IC] = 1.0f;
        if (live) {
            const float r = yn - m;
            chi2 += (double)r * (double)r;
#pragma unroll
            for (int i = 0; i < P; ++i) {
                acc[NH + i] += J[i] * r;
#pragma unroll
                for (int j = i; j < P; ++j) acc[LfSym<P>::idx(i, j)] += J[i] * J[j];
            }
            if (CROSS) {
#pragma unroll
                for (int k = 0; k < L; ++k) {
                    float uk, bk;
                    lf_unpack_bf16(lf_float_bits(ub[k]), uk, bk);
                    acc[NA + k] += r * uk;
                    acc[NA + L + k] += r * bk;
                }
            }
        }
    }
    chi2_out = w.sum(chi2);
    w.sync();
    LfReduceScatter<NAX, LANES / 2, LANES>::run(w, acc, 0, NAX, dst);
    w.sync();
    if (CROSS) {
        // sum r A_k dg_k/dv -> sum r dg_k/dv (an amplitude pinned at its lower bound has no cross term)
        for (int k = w.lane; k < L; k += LANES) {
            const float A = sl[k].aeff;
            const float ia = fabsf(A) > 1e-6f ? LF_RCP(A) : 0.0f;
            dst[NA + k] *= ia;
            dst[NA + L + k] *= ia;
        }
        w.sync();
        if (tie && w.lane == 0) {                          // the dead 6548 column: its sums ride on 6583 with weight R
            dst[NA + cfg.i83] += wk.tie->R * dst[NA + cfg.i48];
            dst[NA + L + cfg.i83] += wk.tie->R * dst[NA + L + cfg.i48];
            dst[NA + cfg.i48] = 0.0f;
            dst[NA + L + cfg.i48] = 0.0f;
        }
        w.sync();                                          // (the solver decides whether the term goes into the matrix)
    }
}

// Solve (D^-1/2 H D^-1/2 + lam I) s = D^-1/2 g with fixed columns pinned, delta = D^-1/2 s.  Returns false if the
// Cholesky factorisation meets a non-positive pivot.  `hg` = [H packed | g] in per-warp memory.
template <int P>
LF_CORE bool lf_solve(const float* hg, unsigned fixedmask, float lam, float (&delta)[P], float& pred)
{
    constexpr int NH = P * (P + 1) / 2;
    float sc[P], M[NH], b[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        float d = hg[LfSym<P>::idx(j, j)];
        bool fx = ((fixedmask >> j) & 1u) || !(d > 1e-30f);
        sc[j] = fx ? 0.0f : LF_RSQRT(d);
        b[j] = hg[NH + j] * sc[j];
    }
#pragma unroll
    for (int i = 0; i < P; ++i)
#pragma unroll
        for (int j = i; j < P; ++j) {
            float v = hg[LfSym<P>::idx(i, j)] * sc[i] * sc[j];
            if (i == j) v = (sc[i] == 0.0f) ? 1.0f : v + lam;
            M[LfSym<P>::idx(i, j)] = v;
        }
    // in-place Cholesky M = U^T U (upper), then forward/back substitution.  Every loop has the compile-time trip
    // count P with the triangular condition inside, so that the unroller resolves all indices (M stays in registers).
    bool ok = true;
#pragma unroll
    for (int i = 0; i < P; ++i) {
        float d = M[LfSym<P>::idx(i, i)];
#pragma unroll
        for (int k = 0; k < P; ++k) if (k < i) d -= M[LfSym<P>::idx(k, i)] * M[LfSym<P>::idx(k, i)];
        if (!(d > 1e-12f)) { ok = false; d = 1.0f; }
        float inv = LF_RSQRT(d);
        M[LfSym<P>::idx(i, i)] = inv;                      // store 1/U_ii
#pragma unroll
        for (int j = 0; j < P; ++j) {
            if (j > i) {
                float v = M[LfSym<P>::idx(i, j)];
#pragma unroll
                for (int k = 0; k < P; ++k) if (k < i) v -= M[LfSym<P>::idx(k, i)] * M[LfSym<P>::idx(k, j)];
                M[LfSym<P>::idx(i, j)] = v * inv;
            }
        }
    }
    float z[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {                          // U^T z = b
        float v = b[i];
#pragma unroll
        for (int k = 0; k < P; ++k) if (k < i) v -= M[LfSym<P>::idx(k, i)] * z[k];
        z[i] = v * M[LfSym<P>::idx(i, i)];
    }
    float s[P];
#pragma unroll
    for (int ii = 0; ii < P; ++ii) {                       // U s = z
        const int i = P - 1 - ii;
        float v = z[i];
#pragma unroll
        for (int k = 0; k < P; ++k) if (k > i) v -= M[LfSym<P>::idx(i, k)] * s[k];
        s[i] = v * M[LfSym<P>::idx(i, i)];
    }
    pred = 0.0f;                                           // predicted decrease of sum r^2: s.(lam s + b)
#pragma unroll
    for (int j = 0; j < P; ++j) { pred += s[j] * (lam * s[j] + b[j]); delta[j] = s[j] * sc[j]; }
    return ok;
}

#if defined(__CUDACC__)
// The same solve with the warp's lanes as the columns of the system (P <= 32): lane j keeps column j of the scaled,
// damped matrix in registers -- by symmetry also its row j -- and ends up with delta_j.  A right-looking Cholesky:
// at step i lane i broadcasts its (current) row i, every lane j > i updates its column, lane i keeps the finished row
// i of U, lanes j < i are done.  After the factorisation lane j therefore holds column j of U in m[0..j) and row j of
// U in m(j..P), which is exactly what the forward (U^T z = b, column-wise) and the backward (U s = z, row-wise)
// substitution need, one broadcast per unknown.  ~300 warp instructions for P = 8 instead of ~720 for the version
// above executed redundantly by all lanes (the LM kernel is issue-bound: profiles/r02_lm_ncu.txt).
template <int P>
__device__ __forceinline__ bool lf_solve_lanes(const float* hg, unsigned fixedmask, float lam, float& delta, float& pred)
{
    static_assert(P <= 32, "one lane per unknown");
    constexpr int NH = P * (P + 1) / 2;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int j = lane < P ? lane : P - 1;                   // lanes beyond P shadow the last column (never read)
    const int Cj = j * P - (j * (j - 1)) / 2 - j;            // packed index of (i, j), i <= j, is C_i + j
    const float d = hg[Cj + j];
    const bool fx = ((fixedmask >> j) & 1u) || !(d > 1e-30f);
    const float sc = fx ? 0.0f : LF_RSQRT(d);
    const float b = hg[NH + j] * sc;
    float m[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int Ci = i * P - (i * (i - 1)) / 2 - i;
        const float sci = __shfl_sync(FULL, sc, i);
        float v = hg[i <= j ? Ci + j : Cj + i] * sci * sc;
        if (i == j) v = fx ? 1.0f : v + lam;
        m[i] = v;
    }
    bool ok = true;
    float invd = 1.0f;                                       // 1 / U_jj
#pragma unroll
    for (int i = 0; i < P; ++i) {
        float piv = __shfl_sync(FULL, m[i], i);
        if (!(piv > 1e-12f)) { ok = false; piv = 1.0f; }
        const float inv = LF_RSQRT(piv);
        const float u = m[i] * inv;                          // lane j > i: U_ij
        if (lane == i) invd = inv;
        const float uu = lane > i ? u : 0.0f;
        if (lane > i) m[i] = u;
#pragma unroll
        for (int k = i + 1; k < P; ++k) {
            const float r = __shfl_sync(FULL, m[k], i) * inv;    // U_ik, from the lane that owns row i
            m[k] = lane == i ? r : fmaf(-r, uu, m[k]);
        }
    }
    float acc = b;                                           // U^T z = b
#pragma unroll
    for (int k = 0; k < P; ++k) {
        const float zk = __shfl_sync(FULL, acc * invd, k);
        acc = fmaf(-(k < lane ? m[k] : 0.0f), zk, acc);
    }
    acc *= invd;                                             // z_j;  U s = z
#pragma unroll
    for (int k = P - 1; k >= 0; --k) {
        const float sk = __shfl_sync(FULL, acc * invd, k);
        acc = fmaf(-(k > lane ? m[k] : 0.0f), sk, acc);
    }
    const float s = acc * invd;
    float t = lane < P ? s * (lam * s + b) : 0.0f;           // predicted decrease of sum r^2: s.(lam s + b)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
    pred = t;
    delta = s * sc;
    return ok;
}
#endif

// Structured secant term of the Hessian.  Gauss-Newton drops S = -sum r_i grad^2 m_i from the Hessian of F = 1/2 sum r^2;
// on noisy spectra (large residuals at the optimum) that makes the iteration contract only linearly, and for a line group
// lost in the noise it oscillates (tests/golden/fit_sincgauss_groups[9]: 50+ evaluations).  The gradients J^T r are
// accurate in FP32 long after sum r^2 itself has sunk into its rounding noise, so S is estimated from them: after every
// evaluation the pair  s = theta_new - theta_old,  y# = (grad F_new - grad F_old) - H s  (H = J^T J of the point the
// iteration stands on afterwards) must satisfy S s = y#; S is sized (Dennis-Gay-Welsch: shrunk by min(1, |s.y#| / |s.S s|),
// so that it fades with the residual) and corrected by the symmetric rank-one formula.  `g` here is the code's J^T r =
// -grad F.  S[P][P] is followed by two [P] scratch vectors.  All lanes call.
template <int P, int LANES>
LF_COLD void lf_secant_update(const LfWarp<LANES>& w, const float* h_old, const float* h_new,
                              const float* th_old, const float* th_new, float* S)
{
    constexpr int NH = P * (P + 1) / 2;
    float* rv = S + P * P;          // y#, then r = y# - tau S s
    float* sv = rv + P;             // s
    float* qv = sv + P;             // S s
    w.sync();
    for (int i = w.lane; i < P; i += LANES) sv[i] = th_new[i] - th_old[i];
    w.sync();
#pragma unroll 1
    for (int i = w.lane; i < P; i += LANES) {
        float Ss = 0.0f, Hs = 0.0f;
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const int a = i < j ? i : j, b = i < j ? j : i;
            Ss += S[i * P + j] * sv[j];
            const int e = a * P - (a * (a - 1)) / 2 + (b - a);
            Hs += 0.5f * (h_old[e] + h_new[e]) * sv[j];
        }
        rv[i] = -(h_new[NH + i] - h_old[NH + i]) - Hs;
        qv[i] = Ss;
    }
    w.sync();
    float sy = 0.0f, sSs = 0.0f;
#pragma unroll
    for (int i = 0; i < P; ++i) { sy += sv[i] * rv[i]; sSs += sv[i] * qv[i]; }
    const float tau = (fabsf(sSs) > fabsf(sy)) ? fabsf(sy) / fabsf(sSs) : 1.0f;
    float abs_sum = 0.0f;
#pragma unroll
    for (int i = 0; i < P; ++i) abs_sum += fabsf((rv[i] - tau * qv[i]) * sv[i]);
    const float rs = sy - tau * sSs;
    const bool use = abs_sum > 0.0f && fabsf(rs) > 1e-4f * abs_sum;
    const float inv = use ? 1.0f / rs : 0.0f;
    w.sync();
    for (int e = w.lane; e < P * P; e += LANES) {
        const int i = e / P, j = e - i * P;
        S[e] = tau * S[e] + (rv[i] - tau * qv[i]) * (rv[j] - tau * qv[j]) * inv;
    }
    w.sync();
}

// [H + S | g] for the solver (dst[NA], per-warp memory).  Returns false -- the caller then proposes a plain Gauss-Newton
// step -- when S would turn a diagonal element of the model Hessian non-positive.
template <int P, int LANES>
LF_COLD bool lf_secant_fold(const LfWarp<LANES>& w, const float* hg, const float* S, float* dst)
{
    constexpr int NH = P * (P + 1) / 2;
    w.sync();
    bool good = true;
    for (int e = w.lane; e < P * P; e += LANES) {
        const int i = e / P, j = e - i * P;
        if (j < i) continue;
        const int q = i * P - (i * (i - 1)) / 2 + (j - i);
        const float v = hg[q] + S[e];
        dst[q] = v;
        if (i == j && hg[q] > 0.0f && !(v > 0.25f * hg[q])) good = false;    // (a dead column has H_jj = 0)
    }
    for (int j = w.lane; j < P; j += LANES) dst[NH + j] = hg[NH + j];
    w.sync();
    return w.all(good);
}

// Box of the reduced parameters (LuciFit.py:614-633 amplitudes/continuum; :539-541 sigma of the last line): every lane
// writes the box of the unknowns it owns (j = lane, lane + LANES, ...) to per-warp memory; `th` must be visible to all
// lanes (callers sync before and after).
template <int MODEL, int L, int NV, int NS, int LANES>
LF_CORE void lf_bounds(const LfCfg& cfg, const LfWarp<LANES>& w, const float* th, float* lo, float* hi)
{
    typedef LfDims<MODEL, L, NV, NS> D;
    for (int j = w.lane; j < D::P; j += LANES) {
        float l = -1e-8f, h = 1.1f;                                      // amplitudes
        if (j >= D::IV && j < D::IS) { l = -INFINITY; h = INFINITY; }
        else if (j >= D::IS && j < D::IC) {
            // sigma box applies to the LAST line only (late-binding lambda in the reference); other groups: beta >= 0
            l = 0.0f; h = INFINITY;
            if (j - D::IS == cfg.sigbox_group) {
                const float vlast = th[D::IV + (NV == 1 ? 0 : cfg.vg[L - 1])];
                const float plast = cfg.p0[L - 1] / (1.0f + vlast * (1.0f / LF_C_KMS));
                h = 10.0f * LF_C_KMS / plast;
            }
        } else if (j == D::IC) { l = -1e-8f; h = 0.99f; }
        if (cfg.freeze) { l = -INFINITY; h = INFINITY; }                 // unconstrained SLSQP on amplitudes + continuum (:699-703)
        lo[j] = l; hi[j] = h;
    }
}

// ---------------------------------------------------------------------------------------------- stage: prep
// Reads the spaxel's spectrum, computes everything Fit.__init__ does plus the initial vector, writes the state.
template <int MODEL, int L, int NV, int NS, int LANES>
LF_CORE void lf_stage_prep(const LfCfg& cfg, const LfWarp<LANES>& w, const float* gspec, float theta_deg,
                           const float* pvel /* [nv] or [1] or null */, const float* pbroad /* [ns] or [1] or null */,
                           const LfWork& wk, float* state)
{
    typedef LfDims<MODEL, L, NV, NS> D;
    const int nz = cfg.nz;
    float* spec = wk.spec;
    float smax = -INFINITY, wmax = -INFINITY, tmax = -INFINITY;
    int nan_ct = 0;
    w.sync();
    // the spectrum is read once from HBM: LF_PREP_MLP independent loads per lane are issued before the first one is
    // consumed, otherwise the stage sits on DRAM latency (this loop is the only global traffic of the stage)
    const bool plain = !cfg.bkg && !cfg.trans;          // the common case keeps the streaming loop to load / store / max / NaN count
    for (int i0 = w.lane; i0 < nz; i0 += LANES * LF_PREP_MLP) {
        float q[LF_PREP_MLP];
#pragma unroll
        for (int j = 0; j < LF_PREP_MLP; ++j) { const int i = i0 + j * LANES; q[j] = i < nz ? gspec[i] : 0.0f; }
        if (plain) {
#pragma unroll
            for (int j = 0; j < LF_PREP_MLP; ++j) {
                const int i = i0 + j * LANES;
                if (i < nz) {
                    const float s = q[j];
                    spec[i] = s;
                    nan_ct += (s != s);
                    smax = fmaxf(smax, s);
                }
            }
            continue;
        }
#pragma unroll
        for (int j = 0; j < LF_PREP_MLP; ++j) {
            const int i = i0 + j * LANES;
            if (i < nz) {
                float s = q[j];
                if (cfg.bkg) s -= cfg.bkg[i];
                spec[i] = s;
                nan_ct += (s != s);
                smax = fmaxf(smax, s);
                float st = s;
                if (cfg.trans) { float t = cfg.trans[i]; if (t > 0.5f) st = s / t; }
                tmax = fmaxf(tmax, st);
            }
        }
    }
    if (plain) tmax = smax;
    w.sync();
    for (int i = cfg.fit_lo + w.lane; i < cfg.fit_hi; i += LANES) wmax = fmaxf(wmax, spec[i]);   // (own writes + sync: visible)
    smax = w.max(smax); wmax = w.max(wmax); tmax = w.max(tmax);
    nan_ct = w.sum(nan_ct);
    w.sync();
    // NaN samples are dropped like the reference's row worker does (LuciBase.py:358-360: sky[good], axis[good]): every
    // reduction below skips them and the later stages leave them out of the sums; a spectrum with nothing left is unusable
    int bad = 0;
    if (nan_ct >= nz) bad |= LF_ST_NAN_INPUT;
    if (!(smax > 0.0f) || !(wmax > 0.0f) || !isfinite(smax)) bad |= LF_ST_BAD_NORM;
    if (bad) {                                         // reference returns zeros for unusable spectra
        if (w.lane == 0) reinterpret_cast<int*>(state)[LF_SO_STATUS] = bad;
        return;
    }
    // window edges: argmin over the axis WITHOUT its NaN samples (only differs when an edge sample itself is NaN)
    int noise_lo = cfg.noise_lo, noise_hi = cfg.noise_hi, cont_lo = cfg.cont_lo, cont_hi = cfg.cont_hi, fit_cut = 0;
    if (nan_ct > 0) {
        auto is_nan = [&](int j) { return spec[j] != spec[j]; };
        auto edge = [&](int idx) { const int j = idx < nz ? idx : nz - 1; return lf_valid_near(cfg.axis, nz, j, cfg.axis[j], is_nan); };
        noise_lo = edge(cfg.noise_lo); noise_hi = edge(cfg.noise_hi);
        cont_lo = edge(cfg.cont_lo); cont_hi = edge(cfg.cont_hi);
        const int fh = edge(cfg.fit_hi);
        if (fh < cfg.fit_hi) { fit_cut = cfg.fit_hi - fh; if (fit_cut > 63) fit_cut = 63; }
        if (fit_cut > 0) {                              // the window maximum is taken over the reference's window
            wmax = -INFINITY;
            for (int i = cfg.fit_lo + w.lane; i < cfg.fit_hi - fit_cut; i += LANES) wmax = fmaxf(wmax, spec[i]);
            wmax = w.max(wmax);
            if (!(wmax > 0.0f)) {
                if (w.lane == 0) reinterpret_cast<int*>(state)[LF_SO_STATUS] = LF_ST_BAD_NORM;
                return;
            }
        }
    }
    // noise: population std of spectrum_normalized over the noise window (LuciFit.py:327-331)
    double noise;
    {
        int n = 0;
        double s1 = 0.0;
        for (int i = noise_lo + w.lane; i < noise_hi; i += LANES) { const float v = spec[i]; if (v == v) { s1 += (double)v; ++n; } }
        s1 = w.sum(s1);
        n = w.sum(n);
        double mean = n > 0 ? s1 / n : 0.0;
        double s2 = 0.0;
        for (int i = noise_lo + w.lane; i < noise_hi; i += LANES) { const float v = spec[i]; if (v == v) { double d = (double)v - mean; s2 += d * d; } }
        s2 = w.sum(s2);
        noise = n > 0 ? sqrt(s2 / n) / (double)smax : NAN;
    }
    // sigma_level: 1 in the free fit (LuciFit.py:661), 3 in the freeze branch (:690)
    const double cont0 = lf_cont_estimate(w, spec, cont_lo, cont_hi, cfg.freeze ? 3.0 : 1.0, wk.sortv, wk.npad) / (double)smax;
    // scale: ML-prior path max(spectrum_restricted) * max(spectrum_T) ; plain path max(spectrum_T)
    const double scale = cfg.ml_scale ? ((double)wmax / (double)smax) * (double)tmax : (double)tmax;
    const double ct = fabs(cos((double)theta_deg * 0.017453292519943295));
    const double corr = 1.0 / ct;
    const double sinc_w = corr * cfg.axis_k;
    // initial vector (LuciFit.py:663-672, 410-423): amplitude = max over +-5 channels around the expected
    // centre minus the continuum guess; centres from the prior velocity; sigma from the prior broadening.
    // priors: one (velocity, broadening) pair per spaxel like the reference's CNN output, or one per tie group
    // (cfg.prior_groups: what a double-component fit needs to leave the symmetric saddle, SURVEY hard part 7)
    float* th = state + LF_SO_TH;
    for (int k = w.lane; k < L; k += LANES) {
        float vel0 = pvel ? pvel[cfg.prior_groups ? cfg.vg[k] : 0] : 0.0f;
        if (vel0 != vel0) vel0 = 0.0f;                 // LuciFit.py:406-409
        double lam = cfg.line_nm[k];
        double pos = 1e7 / (((double)vel0 / 299792.0) * lam + lam);
        int ind = lf_nearest_axis_from(cfg.axis, nz, pos, lf_axis_guess(cfg, pos));
        float amp;
        if (ind + 5 > nz - 1) {
            amp = spec[ind];                           // IndexError branch (LuciFit.py:420-421)
        } else {
            amp = -INFINITY;
            for (int d = -5; d <= 5; ++d) { int i = ind + d; if (i < 0) i += nz; amp = fmaxf(amp, spec[i]); }
        }
        if (!(amp == amp) || amp == -INFINITY) amp = 0.0f;             // every sample around the line is NaN
        th[k] = (float)((double)amp / (double)smax - cont0);
    }
    // Start of the velocity: the prior (one value for all lines) is moved to the maximum of a matched filter within
    // +-LF_REFINE_HALF steps of LF_REFINE_STEP km/s around it -- sum over the lines of the squared positive part of the
    // smoothed, continuum-subtracted spectrum at the expected centres (lucifit_prior.cuh: the same score on a local
    // grid).  A CNN-like prior is good to ~15 km/s, the filter to ~1 km/s; the optimum does not depend on the start,
    // the number of LM iterations does.
    float v_refined = 0.0f;
    bool refined = false;
    if (cfg.refine_prior && NV == 1 && !cfg.prior_groups && !cfg.freeze) {
        float vc = pvel ? pvel[0] : 0.0f;                            // no prior: the reference starts at 0 (LuciFit.py:155)
        if (vc != vc) vc = 0.0f;
        const float contr = (float)(cont0 * (double)smax);
        // a maximum on the edge of the grid means the prior is further off than the grid reaches: the grid is moved
        // there and the search repeated (at most LF_REFINE_ROUNDS grids, +-108 km/s in all)
#pragma unroll 1
        for (int round = 0; round < LF_REFINE_ROUNDS && !refined; ++round) {
            w.sync();
            for (int j = w.lane; j <= 2 * LF_REFINE_HALF; j += LANES) {
                const float vt = (vc + (float)(j - LF_REFINE_HALF) * LF_REFINE_STEP) * (1.0f / LF_C_KMS);
                const float shift = vt / (1.0f + vt);
                float acc = 0.0f;
                for (int k = 0; k < L; ++k) {
                    const float t = (cfg.p0_off[k] - cfg.p0[k] * shift) * cfg.ichan;
                    const int i = cfg.fit_lo + (int)floorf(t);
                    if (i < 1 || i + 2 >= nz) continue;
                    const float f = t - floorf(t);
                    const float a = 0.25f * spec[i - 1] + 0.5f * spec[i] + 0.25f * spec[i + 1];
                    const float b = 0.25f * spec[i] + 0.5f * spec[i + 1] + 0.25f * spec[i + 2];
                    const float val = fmaxf(a + f * (b - a) - contr, 0.0f);          // (fmaxf drops a NaN)
                    acc += val * val;
                }
                wk.sortv[j] = acc;
            }
            w.sync();
            int jb = 0;
            for (int j = 1; j <= 2 * LF_REFINE_HALF; ++j) if (wk.sortv[j] > wk.sortv[jb]) jb = j;
            if (!(wk.sortv[jb] > 0.0f)) break;                       // nothing above the continuum: keep the prior
            if (jb > 0 && jb < 2 * LF_REFINE_HALF) {
                float dv = (float)(jb - LF_REFINE_HALF) * LF_REFINE_STEP;
                const float a = wk.sortv[jb - 1], b = wk.sortv[jb], d = wk.sortv[jb + 1];
                const float den = a - 2.0f * b + d;
                if (den < 0.0f) dv += LF_REFINE_STEP * fminf(fmaxf(0.5f * (a - d) / den, -0.5f), 0.5f);
                v_refined = vc + dv;
                refined = true;
            } else {
                vc += (float)(jb - LF_REFINE_HALF) * LF_REFINE_STEP;
            }
        }
        w.sync();
    }
    if (w.lane == 0) {
#pragma unroll
        for (int g = 0; g < NV; ++g) {
            float v = pvel ? pvel[(cfg.prior_groups && g < cfg.nv) ? g : 0] : 0.0f;
            th[D::IV + g] = (v != v) ? 0.0f : v;
        }
        if (refined) th[D::IV] = v_refined;
#pragma unroll
        for (int h = 0; h < NS; ++h) {
            float b = pbroad ? pbroad[(cfg.prior_groups && h < cfg.ns) ? h : 0] : 0.0f;
            th[D::IS + h] = (b != b) ? 1.0f : fabsf(b);
        }
        th[D::IC] = (float)cont0;
        double* sd = reinterpret_cast<double*>(state);
        sd[LF_SO_NOISE / 2] = noise; sd[LF_SO_SCALE / 2] = scale; sd[LF_SO_CORR / 2] = corr; sd[LF_SO_SINCW / 2] = sinc_w;
        state[LF_SO_INVW] = 1.0f / wmax;
        reinterpret_cast<int*>(state)[LF_SO_STATUS] = nan_ct > 0 ? (LF_ST_NAN_SAMPLES | (fit_cut << LF_ST_CUT_SHIFT)) : 0;
    }
    w.sync();
}

// Stages the normalised fit window and the phase table of one spaxel into per-warp memory:
//   yw[n] = (cube[fit_lo + n] - bkg) * mul, optionally divided by the transmission (the `out` stage).
template <int MODEL, int LANES>
LF_CORE void lf_stage_window(const LfCfg& cfg, const LfWarp<LANES>& w, const float* xo, const float* gspec, float mul,
                             bool use_trans, double sinc_w, const LfWork& wk, int fit_cut)
{
    const int nfit = cfg.fit_hi - cfg.fit_lo;
    w.sync();
    for (int n = w.lane; n < nfit; n += LANES) {
        int i = cfg.fit_lo + n;
        float s = n < nfit - fit_cut ? gspec[i] : NAN;        // samples beyond the reference's window count as dropped
        if (cfg.bkg) s -= cfg.bkg[i];
        if (use_trans && cfg.trans) { float t = cfg.trans[i]; if (t > 0.5f) s = s / t; }
        wk.yw[n] = s * mul;
    }
    if (MODEL != LF_GAUSSIAN) lf_phase_table(w, xo, nfit, 1.0 / sinc_w, wk.sw, wk.cw);
    w.sync();
}

// ------------------------------------------------------------------------------------------------ stage: lm
// Levenberg-Marquardt on sum r^2 (the noise is a common factor of the reference's nll, LuciFit.py:516, and does not
// move the optimum).  The accepted and the trial normal equations live in per-warp memory (wk.hb), theta / trial
// theta / box in wk.thb, so that the evaluation loop owns the registers.
//
// The stage runs as two kernels.  PHASE 1 (the bulk) is plain projected Gauss-Newton / LM and gives up after
// LF_FAST_ITERS accepted iterations, leaving theta, the damping and LF_ST_MAXITER in the state record; PHASE 2 picks
// exactly those spaxels up (a few per cent: faint spectra, line groups lost in the noise, blends) and finishes them
// with the structured secant term in the model (lf_secant_update), which costs code the bulk should not carry through
// the instruction cache.
template <int MODEL, int L, int NV, int NS, int LANES, int PHASE>
LF_CORE void lf_stage_lm(const LfCfg& cfg, const LfWarp<LANES>& w, const float* xo, const float* gspec, const LfWork& wk,
                         float* state)
{
    typedef LfDims<MODEL, L, NV, NS> D;
    constexpr int P = D::P, NA = D::NAX;              // (stride of the two normal-equation buffers, cross sums included)
    int status = reinterpret_cast<const int*>(state)[LF_SO_STATUS];
    if (status & LF_ST_BADMASK) return;
    if (PHASE == 2 && (status & (LF_ST_MAXITER | LF_ST_CONVERGED)) != LF_ST_MAXITER) return;
    const double sinc_w_d = reinterpret_cast<const double*>(state)[LF_SO_SINCW / 2];
    const float sinc_w = (float)sinc_w_d;
    const double inv_w = 1.0 / sinc_w_d;
    const int nfit = cfg.fit_hi - cfg.fit_lo;
    lf_stage_window<MODEL, LANES>(cfg, w, xo, gspec, state[LF_SO_INVW], false, sinc_w_d, wk,
                                  (status & LF_ST_CUT_MASK) >> LF_ST_CUT_SHIFT);
    float* th = wk.thb;
    float* tht = wk.thb + P;
    float* lo = wk.thb + 2 * P;
    float* hi = wk.thb + 3 * P;
    float* hcur = wk.hb;
    float* htri = wk.hb + NA;
    float* Ssec = wk.hb + 2 * NA;                      // secant term S[P][P] + scratch (lf_secant_update)
    const bool use_sec = PHASE == 2 && LF_SECANT && !cfg.freeze;           // (the freeze branch is linear: S = 0)
    if (PHASE == 2) for (int e = w.lane; e < P * P; e += LANES) Ssec[e] = 0.0f;
    for (int j = w.lane; j < P; j += LANES) th[j] = state[LF_SO_TH + j];
    w.sync();
    lf_bounds<MODEL, L, NV, NS, LANES>(cfg, w, th, lo, hi);
    w.sync();                                            // (the box of sigma reads the velocity: clamp only afterwards)
    for (int j = w.lane; j < P; j += LANES) { const float t = fminf(fmaxf(th[j], lo[j]), hi[j]); th[j] = t; tht[j] = t; }
    w.sync();
    int evals = 0, iters = 0, flags = status & (LF_ST_NAN_SAMPLES | LF_ST_CUT_MASK), tries = 0;
    float lam = LF_LM_LAM0, nu = 2.0f, pred = 0.0f, pred_prev = 0.0f;
    if (PHASE == 2) { iters = status & 0xff; evals = (status >> 8) & 0xff; flags |= status & LF_ST_SINGULAR; lam = state[LF_SO_LAM]; }
    const int iters0 = iters;
    bool sec_have = false;
    double chi2 = 0.0;
    const double ftol = (double)cfg.ftol;
    bool first = true, done = false, sec_applied = false, sec_skip = false, cross_off = false, cross_on = false, near_opt = false;
    unsigned fixedmask = 0;
    while (true) {
        // ---- evaluate the trial point
        double chi2t;
        lf_setup_lines<MODEL, L, NV, NS, LANES>(cfg, w, tht, sinc_w, inv_w, wk.lines, wk.tie);
        lf_eval<MODEL, L, NV, NS, LANES>(cfg, w, xo, wk, nfit, sinc_w, tht[D::IC], chi2t, htri);
        ++evals;
        bool accepted = false, learn = false;
        sec_skip = false;
        if (first) {
            first = false; accepted = true; chi2 = chi2t;
        } else {
            const bool finite = chi2t == chi2t;
            // An FP32 evaluation resolves sum r^2 to a few 1e-6 relative; once the predicted decrease is that small
            // the gradient (still accurate) is trusted and the step is taken even if chi2 appears to rise by noise.
            const double slack = ((double)pred <= LF_TUNE_SLACK_PRED * chi2) ? LF_TUNE_SLACK * chi2 : 0.0;
            if (finite && chi2t <= chi2 + slack) {
                double act = chi2 - chi2t;
                float rho = pred > 0.0f ? (float)(act / (double)pred) : 1.0f;
                // below the FP32 resolution of sum r^2 the measured gain is rounding noise: a negative one must not
                // raise the damping (that shortens the remaining steps and fakes convergence)
                if ((double)pred <= LF_TUNE_NOISE_PRED * chi2) rho = 1.0f;
                float f = 1.0f - (2.0f * rho - 1.0f) * (2.0f * rho - 1.0f) * (2.0f * rho - 1.0f);
                lam *= fmaxf(LF_LM_DEC, f);
                lam = fmaxf(lam, 1e-7f);
                nu = 2.0f;
                accepted = true;
                ++iters;
                learn = use_sec && (double)pred <= LF_SEC_LEARN * chi2;
                near_opt = rho > 0.25f && (double)pred <= LF_CROSS_GATE * chi2;
                chi2 = chi2t;
            } else {
                learn = use_sec && finite && !sec_applied && (double)pred <= LF_SEC_LEARN * chi2;
                // rejected: keep the accepted system and increase the damping; inside the FP32 noise floor of
                // sum r^2 (relative ~1e-6) a rejection only means "converged"
                lam = fmaxf(lam * nu, LF_LM_LAMREJ); nu *= 2.0f;
                ++tries;
                if (PHASE == 2 && sec_applied) {
                    // the rejected proposal came from the secant model: drop the term, propose a Gauss-Newton step instead
                    w.sync();
                    for (int e = w.lane; e < P * P; e += LANES) Ssec[e] = 0.0f;
                    w.sync();
                    sec_have = false;
                } else if (finite && (double)pred <= LF_TUNE_REJ_CONV * chi2) { done = true; flags |= LF_ST_CONVERGED; }
                if (D::CROSS && cross_on && !done) {
                    // the rejected proposal came from the matrix with the exact cross term: this point goes on with plain
                    // Gauss-Newton (the term only stays where it helps)
                    lf_cross_apply<MODEL, L, NV, NS, LANES>(w, hcur, 1.0f);
                    cross_on = false; cross_off = true;
                }
            }
        }
        if (PHASE == 2 && learn) { lf_secant_update<P, LANES>(w, hcur, htri, th, tht, Ssec); sec_have = true; }
        if (accepted) {
            float* t = hcur; hcur = htri; htri = t;       // the trial system becomes the accepted one
            w.sync();
            for (int j = w.lane; j < P; j += LANES) th[j] = tht[j];
            w.sync();
            lf_bounds<MODEL, L, NV, NS, LANES>(cfg, w, th, lo, hi);
            w.sync();
            tries = 0;
            cross_off = false; cross_on = false;           // (the accepted system arrives without the cross term)
            fixedmask = 0;
            for (int j = w.lane; j < P; j += LANES) {      // active set: at a bound with the gradient pushing outwards
                const float gj = hcur[D::NH + j], tj = th[j];
                const bool at_lo = tj <= lo[j] && gj <= 0.0f;
                const bool at_hi = tj >= hi[j] && gj >= 0.0f;
                if (at_lo || at_hi) fixedmask |= 1u << j;
            }
            fixedmask = w.or_all(fixedmask);
            if (cfg.freeze) fixedmask |= ((1u << D::IC) - 1u) & ~((1u << D::IV) - 1u);   // pin every v and beta column
            if (!(iters < cfg.max_iter && evals < 250)) { flags |= LF_ST_MAXITER; done = true; }
        }
        if (done) break;
        // ---- propose the next trial point
        bool go = false;
        while (!go && !done) {
            if (tries >= 12) { done = true; flags |= LF_ST_MAXITER; break; }
            // The exact (amplitude, velocity / broadening) term of the Hessian joins the model once the iteration is near the
            // optimum -- the step that led here was a good one (gain ratio > 1/4) and predicted less than LF_CROSS_GATE of
            // sum r^2: there the residuals are noise and the term is what Gauss-Newton lacks.  Further out they are model
            // mismatch, the corrected matrix can be indefinite or, worse, positive definite and misleading (a line group
            // walking to zero width on a handful of spaxels per cube), and plain Gauss-Newton is the safer model.
            if (D::CROSS && near_opt && !cross_on && !cross_off) {
                lf_cross_apply<MODEL, L, NV, NS, LANES>(w, hcur, -1.0f);
                cross_on = true;
            }
            constexpr int PL = (P + LANES - 1) / LANES;      // unknowns per lane: j = t LANES + lane
            float delta[PL];
            bool apply_sec = PHASE == 2 && use_sec && sec_have && !sec_skip && iters > iters0;
            if (PHASE == 2 && apply_sec) apply_sec = lf_secant_fold<P, LANES>(w, hcur, Ssec, htri);   // (the trial system is dead here)
            sec_applied = apply_sec;
            bool ok;
#if defined(__CUDA_ARCH__)
            if constexpr (LANES == 32) ok = lf_solve_lanes<P>(apply_sec ? htri : hcur, fixedmask, lam, delta[0], pred);
            else
#endif
                ok = lf_solve<P>(apply_sec ? htri : hcur, fixedmask, lam, delta, pred);
            if (!ok && apply_sec) {                          // J^T J + S is not positive definite: fall back to Gauss-Newton
                sec_skip = true;                             // (this proposal)
                continue;
            }
            if (!ok && D::CROSS && cross_on) {              // the matrix with the exact cross term is not positive definite here:
                lf_cross_apply<MODEL, L, NV, NS, LANES>(w, hcur, 1.0f);   // take the term out, plain Gauss-Newton for this point
                cross_on = false; cross_off = true;
                continue;
            }
            if (!ok) { lam *= 10.0f; flags |= LF_ST_SINGULAR; ++tries; continue; }
            if (D::CROSS && cross_on) {
                // a step of the cross-corrected model that runs a free parameter into its box (a broadening to 0, an amplitude
                // to its floor) is where the few bad basins of that model lie: this point goes on with plain Gauss-Newton
                bool clipped = false;
#pragma unroll
                for (int t = 0; t < PL; ++t) {
                    const int j = t * LANES + w.lane;
                    if (j < P && !((fixedmask >> j) & 1u)) { const float v = th[j] + delta[t]; clipped |= v < lo[j] || v > hi[j]; }
                }
                if (w.any(clipped)) {
                    lf_cross_apply<MODEL, L, NV, NS, LANES>(w, hcur, 1.0f);
                    cross_on = false; cross_off = true;
                    continue;
                }
            }
            float rel = 0.0f;
            w.sync();
#pragma unroll
            for (int t = 0; t < PL; ++t) {
                const int j = t * LANES + w.lane;
                if (j < P) {
                    const float tj = th[j];
                    const float tn = fminf(fmaxf(tj + delta[t], lo[j]), hi[j]);
                    const float sc = (j >= D::IV && j < D::IS) ? 1.0f : fmaxf(fabsf(tj), LF_CONV_AMP_FLOOR);   // velocities: km/s absolute
                    rel = fmaxf(rel, fabsf(tn - tj) * LF_RCP(sc));
                    tht[j] = tn;
                }
            }
            rel = w.max(rel);
            w.sync();
            // The predicted decrease is below what an FP32 evaluation of sum r^2 can resolve: take the (Newton-like)
            // step and stop -- evaluating it would only return rounding noise.  What is left AFTER this step is the
            // predicted decrease of the next one; where the iteration contracts fast (two consecutive accepted steps,
            // the second predicted to gain at most LF_CONV_CONTRACTION of the first) that is extrapolated as
            // pred^2 / pred_prev, and the step is the last one once the extrapolation is a tenth of the tolerance.
            // Slowly contracting fits (flat valleys, blended lines) never qualify and run down to the tolerance itself.
#ifdef LF_TRACE
            printf("it %d ev %d chi2 %.10e pred %.3e pred/chi2 %.3e rel %.3e lam %.2e tries %d | th", iters, evals, chi2, pred, pred / chi2, rel, lam, tries);
            for (int j = 0; j < P; ++j) printf(" %.7g", tht[j]);
            printf("\n");
#endif
            bool conv = (double)pred <= ftol * chi2;
            if (!conv && !apply_sec && tries == 0 && pred_prev > 0.0f && pred <= LF_CONV_CONTRACTION * pred_prev &&
                (double)pred <= LF_CONV_PRED_CAP * chi2)
                conv = (double)pred * (double)(pred / pred_prev) <= 0.1 * ftol * chi2;
            if (conv) {
                // The step about to be taken blindly leaves an error of about rho/(1 - rho) times its own size when the
                // iteration contracts linearly at rate rho (Gauss-Newton on noisy spectra does; rho is measured from the
                // predicted gains of this step and the previous one, |step| ~ sqrt(pred)).  The weak lines of a spectrum
                // sit in flat directions of sum r^2 where a gain of 1e-8 still means 1e-3 of their amplitude, so the
                // objective test alone is not enough for the 1e-3 flux bar: the estimated remaining error in units of
                // the parameters (amplitudes relative with a floor of 1e-3 of the peak, velocities in km/s) must be below
                // LF_CONV_REMAIN as well.
                float rho = (tries == 0 && pred_prev > 0.0f) ? sqrtf(pred / pred_prev) : LF_CONV_RHO0;
                rho = fminf(rho, 0.9f);
                // ... unless the predicted gain is already LF_CONV_FLAT times below the tolerance: a direction that flat
                // (the width of a line group lost in the noise) is not determined by the data to 1e-3, iterating on
                // would only follow the FP32 rounding of the residuals
                // (with the secant term in the model the steps no longer contract linearly: the step itself must be small)
                if (apply_sec) rho = 0.5f;
                conv = rel * rho < LF_CONV_REMAIN * (1.0f - rho) || (double)pred <= LF_CONV_FLAT * ftol * chi2;
            }
            if (conv && rel < 1e-2f) {
                for (int j = w.lane; j < P; j += LANES) th[j] = tht[j];
                w.sync();
                done = true; flags |= LF_ST_CONVERGED;
                ++iters;
                break;
            }
            // not converged after LF_FAST_ITERS iterations: the second kernel goes on from the accepted point
            if (PHASE == 1 && iters >= LF_FAST_ITERS && !cfg.freeze) { flags |= LF_ST_MAXITER; done = true; break; }
            pred_prev = tries == 0 ? pred : 0.0f;                 // a damped retry breaks the chain of plain steps
            go = true;
        }
        if (done) break;
    }
    bool at_bound = false;
    for (int j = w.lane; j < P; j += LANES) {
        const float tj = th[j];
        at_bound |= tj <= lo[j] || tj >= hi[j];
        state[LF_SO_TH + j] = tj;
    }
    if (w.any(at_bound)) flags |= LF_ST_BOUND;
    if (w.lane == 0) {
        reinterpret_cast<int*>(state)[LF_SO_STATUS] = (iters & 0xff) | ((evals & 0xff) << 8) | flags;
        if (PHASE == 1) state[LF_SO_LAM] = lam;
    }
    w.sync();
}

// Expand the reduced solution to the reference's fit_sol layout [A, p, sigma]*L + [continuum] in normalised units:
// lane k writes line k of apd = [A | p | sigma] (float64).
template <int MODEL, int L, int NV, int NS, int LANES>
LF_CORE void lf_expand(const LfCfg& cfg, const LfWarp<LANES>& w, float sinc_w, const float* th, double* apd)
{
    typedef LfDims<MODEL, L, NV, NS> D;
    w.sync();
    for (int k = w.lane; k < L; k += LANES) {
        float v = th[D::IV + (NV == 1 ? 0 : cfg.vg[k])], b = th[D::IS + (NS == 1 ? 0 : cfg.sg[k])];
        double p = 1e7 / (cfg.line_nm[k] * (1.0 + (double)v / 299792.0));
        double sg = p * (double)b / 299792.0;
        double A = (double)th[k];
        bool tie = cfg.i48 >= 0 && k == cfg.i48;
        if (tie) {
            float v83 = th[D::IV + (NV == 1 ? 0 : cfg.vg[cfg.i83])], b83 = th[D::IS + (NS == 1 ? 0 : cfg.sg[cfg.i83])];
            if (MODEL == LF_SINC && (b == 0.0f || b83 == 0.0f)) tie = false;
            if (tie) {
                double p83 = 1e7 / (cfg.line_nm[cfg.i83] * (1.0 + (double)v83 / 299792.0));
                double sg83 = p83 * (double)b83 / 299792.0;
                LfLine t;
                lf_line_setup(MODEL, 0.f, 0.f, (float)sg83, sinc_w, t);
                float s83 = t.sig;
                lf_line_setup(MODEL, 0.f, 0.f, (float)sg, sinc_w, t);
                float s48 = t.sig;
                bool same = (NS == 1) || cfg.sg[cfg.i83] == cfg.sg[cfg.i48];
                float R = lf_tie_ratio(MODEL, (float)p83, b83, s83, (float)p, b, s48, same, sinc_w);
                A = (double)th[cfg.i83] * (double)R;
            }
        }
        apd[k] = A; apd[L + k] = p; apd[2 * L + k] = sg;
    }
    w.sync();
}

// ----------------------------------------------------------------------------------------------- stage: out
// Row layout of the per-spaxel result (K = 7 L + 5 floats), the 12 quantities Luci.fit_calc collects
// (LuciBase.py:381-392, 406):
//   [0,L) amplitudes | [L,2L) fluxes | [2L,3L) flux errors | [3L,4L) velocities | [4L,5L) velocity errors |
//   [5L,6L) broadenings | [6L,7L) broadening errors | chi2 | corr | axis_step | continuum | continuum error
template <int MODEL, int L, int NV, int NS, int LANES, bool STAGE_ROW = false>
LF_CORE void lf_stage_out(const LfCfg& cfg, const LfWarp<LANES>& w, const float* xo, const float* gspec, const LfWork& wk,
                          const float* state, const double* unc /* [3L+1] normalised or null */,
                          float* out_row, int32_t* out_status, float* out_sol, float* out_unc)
{
    typedef LfDims<MODEL, L, NV, NS> D;
    const int status = reinterpret_cast<const int*>(state)[LF_SO_STATUS];
    if (status & LF_ST_BADMASK) {
        for (int j = w.lane; j < 7 * L + 5; j += LANES) out_row[j] = 0.0f;
        if (out_sol) for (int j = w.lane; j < 3 * L + 1; j += LANES) out_sol[j] = 0.0f;
        if (out_unc) for (int j = w.lane; j < 3 * L + 1; j += LANES) out_unc[j] = 0.0f;
        if (w.lane == 0) *out_status = status;
        return;
    }
    const double* sd = reinterpret_cast<const double*>(state);
    const double noise = sd[LF_SO_NOISE / 2], scale = sd[LF_SO_SCALE / 2], corr = sd[LF_SO_CORR / 2], sinc_w_d = sd[LF_SO_SINCW / 2];
    const float sinc_w = (float)sinc_w_d;
    const double inv_w = 1.0 / sinc_w_d;
    const float* th = state + LF_SO_TH;
    const int nfit = cfg.fit_hi - cfg.fit_lo;
    lf_stage_window<MODEL, LANES>(cfg, w, xo, gspec, (float)(1.0 / scale), true, sinc_w_d, wk,
                                  (status & LF_ST_CUT_MASK) >> LF_ST_CUT_SHIFT);
    double* apd = wk.apd;
    lf_expand<MODEL, L, NV, NS, LANES>(cfg, w, sinc_w, th, apd);
    const double cont = (double)th[D::IC];
    // chi2 with every line centre snapped to the nearest axis sample (Model.plot, LuciFunctions.py:113-115 etc.;
    // LuciFit.py:799-801, 1065-1068).  The scale cancels: chi2 = sum (s_T/scale - m)^2 / (noise (N - P_full)).
    LfLineX* sl = wk.lines;
    for (int k = w.lane; k < L; k += LANES) {
        int idx = lf_nearest_axis_from(cfg.axis, cfg.nz, apd[L + k], lf_axis_guess(cfg, apd[L + k]));
        if (status & LF_ST_NAN_SAMPLES)                 // Model.plot snaps on the axis the NaN samples were removed from
            idx = lf_valid_near(cfg.axis, cfg.nz, idx, apd[L + k], [&](int j) {
                float v = gspec[j];
                if (cfg.bkg) v -= cfg.bkg[j];
                return v != v;
            });
        LfLineX x;
        float pps, ppl;
        lf_split(cfg.axis[idx] - cfg.x_ref, pps, ppl);
        lf_line_setup(MODEL, pps, ppl, (float)apd[2 * L + k], sinc_w, x.ln);
        x.aeff = (float)apd[k];
        x.dpdv = x.dsdv = x.dsdb = 0.0f; x.vg = x.sg = 0; x.p = x.beta = 0.0f; x.Gk = x.Ek = 0.0f;
        x.ck = 1.0f; x.sk = 0.0f;
        if (MODEL != LF_GAUSSIAN) lf_phase(((double)x.ln.pp + (double)x.ln.dp) * inv_w, x.sk, x.ck);
        lf_fold_wing(x);
        sl[k] = x;
    }
    w.sync();
    const float iw = 1.0f / sinc_w;
    double S = 0.0;
    int nvalid = 0;
    for (int n0 = 0; n0 < nfit; n0 += LANES) {
        const int n = n0 + w.lane;
        const bool live = n < nfit;
        const float x = live ? xo[n] : 0.0f;
        float Sx = 0.0f, Cx = 1.0f;
        if (MODEL != LF_GAUSSIAN && live) { Sx = wk.sw[n]; Cx = wk.cw[n]; }
        float m = (float)cont;
#pragma unroll 1
        for (int k = 0; k < L; ++k) {
            const LfLineX& X = sl[k];
            const float dx = lf_dx(X.ln, x);
            int hint = LF_HINT_LANE;
            float s = 0.0f, c = 1.0f;
            if (MODEL != LF_GAUSSIAN) { s = Sx * X.ck - Cx * X.sk; c = Cx * X.ck + Sx * X.sk; }
            if (MODEL == LF_SINCGAUSS) {
                const float b = dx * X.ln.is2s;
                const float r2 = b * b + X.a2;
                if (w.all(b * b >= 40.0f && r2 >= LF_W_R2_ASY)) {      // whole chunk in the wing of this line
                    float T, T2, E2;
                    lf_wing_terms(b, X.ln.a, b * b, X.a2, r2, s, c, T, T2, E2);
                    m += X.c_m * T;
                    continue;
                }
                hint = lf_region_hint(w, X.ln, dx);
            }
            float g, gp, gs;
            lf_profile_sc<MODEL>(X.ln, dx, s, c, iw, hint, g, gp, gs);
            m += X.aeff * g;
        }
        if (live) {
            const float r = wk.yw[n] - m;
            if (r == r) { S += (double)r * (double)r; ++nvalid; }       // NaN samples were dropped (LuciBase.py:358-360)
        }
    }
    S = w.sum(S);
    nvalid = w.sum(nvalid);
    const double chi2 = S / (noise * (double)(nvalid - (3 * L + 1)));
    // STAGE_ROW (rows that land in another GPU's memory through a peer mapping, see lucifit_fit_stages): the row is assembled
    // in the warp's scratch (the window is dead by now; yw | sw | cw hold >= 9L+6 floats) and leaves as whole 128-byte
    // stores of consecutive lanes -- 2 NVLink writes per row instead of 12 scattered 4..20-byte ones.  Into local memory
    // the direct stores are faster (L2 merges them), so that instance of the kernel is left exactly as it was.
    float* const dst_row = out_row;
    if (STAGE_ROW) { w.sync(); out_row = wk.yw; }
    const double FWHM_COEFF = 2.3548200450309493;
    const double SQ2PI = 2.5066282746310002, PI = 3.141592653589793;
    for (int k = w.lane; k < L; k += LANES) {
        double Ak = apd[k] * scale, s = apd[2 * L + k], pk = apd[L + k];
        double uA = unc ? unc[3 * k] * scale : 0.0, up = unc ? unc[3 * k + 1] : 0.0, us = unc ? unc[3 * k + 2] : 0.0;
        double flux, ferr;
        if (MODEL == LF_GAUSSIAN) {
            flux = (1.20671 / FWHM_COEFF) * SQ2PI * Ak * s;
        } else if (MODEL == LF_SINC) {
            flux = PI * sqrt(PI) * Ak * sinc_w_d;
        } else {
            // (erf in FP32: 1e-7 of a flux that is stored as float32; the float64 routine costs ~120 instructions)
            flux = (1.20671 / (PI * FWHM_COEFF)) * Ak * (SQ2PI * s / (double)erff((float)(s / (1.4142135623730951 * sinc_w_d))));
        }
        double lam = cfg.line_nm[k];
        double v = 299792.0 * ((1e7 / pk - lam) / lam);                         // LuciFitParameters.py:34-36
        double broad = pk > 0 ? fabs(299792.0 * s / pk) : 0.0;                  // :82-88
        double verr, berr;
        if (unc) {
            if (MODEL == LF_SINCGAUSS) {
                double c0 = 1.4142135623730951 * sinc_w_d;
                double t = erf(s / c0) - (2.0 * PI) / (sqrt(PI) * c0) * exp(-s * s / (c0 * c0));
                ferr = flux * sqrt((uA / Ak) * (uA / Ak) + (us / s) * (us / s) * t * t);
            } else {
                ferr = flux * sqrt((uA / Ak) * (uA / Ak) + (us / s) * (us / s));
            }
            verr = 299792.0 * up * (1e7 / (lam * pk * pk));                     // :57
            berr = (pk > 0 && s > 0) ? (299792.0 * s / pk) * sqrt((us / s) * (us / s) + (up / pk) * (up / pk)) : 0.0;
        } else {
            // uncertainty_bool = False: every 1-sigma input is 0 and the expressions above reduce to 0 -- or to the
            // 0/0 = NaN the reference produces when an amplitude or a width collapsed to 0 (kept: same maps)
            ferr = (Ak == 0.0 || s == 0.0 || flux != flux) ? NAN : flux * 0.0;
            verr = 0.0;
            berr = 0.0;
        }
        out_row[k] = (float)Ak; out_row[L + k] = (float)flux; out_row[2 * L + k] = (float)ferr;
        out_row[3 * L + k] = (float)v; out_row[4 * L + k] = (float)verr;
        out_row[5 * L + k] = (float)broad; out_row[6 * L + k] = (float)berr;
        if (out_sol) { out_sol[3 * k] = (float)Ak; out_sol[3 * k + 1] = (float)pk; out_sol[3 * k + 2] = (float)s; }
    }
    if (w.lane == 0) {
        out_row[7 * L] = (float)chi2;
        out_row[7 * L + 1] = (float)corr;
        out_row[7 * L + 2] = (float)sinc_w_d;          // axis_step == sinc_width (LuciFit.py:214-215, 222-223)
        out_row[7 * L + 3] = (float)(cont * scale);
        out_row[7 * L + 4] = (float)(unc ? unc[3 * L] * scale : 0.0);
        if (out_sol) out_sol[3 * L] = (float)(cont * scale);
        *out_status = status & ~LF_ST_CUT_MASK;
    }
    if (STAGE_ROW) {
        w.sync();
        for (int j = w.lane; j < 7 * L + 5; j += LANES) dst_row[j] = out_row[j];
    }
    if (out_unc) {
        for (int j = w.lane; j < 3 * L + 1; j += LANES) {
            double u = unc ? unc[j] : 0.0;
            if (j == 3 * L || (j % 3) == 0) u *= scale;
            out_unc[j] = (float)u;
        }
    }
    w.sync();
}

// ----------------------------------------------------------------------------------------------- stage: unc
// 1-sigma errors exactly as the reference defines them (LuciFit.py:713-722, LuciUtility.py:409-434): finite
// difference Hessian of the nll with delta = 0.1 on ALL 3L+1 unconstrained parameters [A, p, sigma]*L + [c]
// (normalised units), unc = sqrt(|diag(-(H/2)^-1)|).  The 4 (3L+1)^2 objective calls of hessianComp are
// restructured using the additivity of the model over lines: for i, j in different lines
// H_ij = sum D_i D_j / noise^2 with D_i the central difference of the line profile; inside a line the exact
// second-difference expressions are accumulated.  13 unit-profile evaluations per line and sample.
// variant order: 0 (p,s) 1 (p+d,s) 2 (p-d,s) 3 (p,s+d) 4 (p,s-d) 5 (p+2d,s) 6 (p-2d,s) 7 (p,s+2d) 8 (p,s-2d)
//                9 (p+d,s+d) 10 (p+d,s-d) 11 (p-d,s+d) 12 (p-d,s-d)
LF_HD float lf_var_dp(int v)
{
    return (v == 1 || v == 9 || v == 10) ? 0.1f : (v == 2 || v == 11 || v == 12) ? -0.1f : v == 5 ? 0.2f : v == 6 ? -0.2f : 0.0f;
}
LF_HD float lf_var_ds(int v)
{
    return (v == 3 || v == 9 || v == 11) ? 0.1f : (v == 4 || v == 10 || v == 12) ? -0.1f : v == 7 ? 0.2f : v == 8 ? -0.2f : 0.0f;
}

template <int MODEL, int L, int NV, int NS, int LANES>
LF_CORE void lf_stage_unc(const LfCfg& cfg, const LfWarp<LANES>& w, const float* xo, const float* gspec, const LfWork& wk,
                          const float* state, double* unc)
{
    typedef LfDims<MODEL, L, NV, NS> D;
    constexpr int PF = 3 * L + 1;
    constexpr int TS = LANES + 4;                        // row stride of the D tile [PF][TS]: 16-byte rows, skewed banks
    constexpr int NE = PF * (PF + 1) / 2;
    constexpr int SLOTS = (NE + LANES - 1) / LANES;
    const int status = reinterpret_cast<const int*>(state)[LF_SO_STATUS];
    if (status & LF_ST_BADMASK) return;
    const double* sd = reinterpret_cast<const double*>(state);
    const double noise = sd[LF_SO_NOISE / 2], sinc_w_d = sd[LF_SO_SINCW / 2];
    const float sinc_w = (float)sinc_w_d;
    const double inv_w = 1.0 / sinc_w_d;
    const float* th = state + LF_SO_TH;
    const float dl = 0.1f;
    const int nfit = cfg.fit_hi - cfg.fit_lo;
    // The sample loop below visits every chunk of the window ONCE, so the data and the phase of a chunk are formed when
    // the chunk is reached (the next chunk's data are requested one chunk ahead) instead of being staged for the whole
    // window: 3 nfit floats of per-warp shared memory less, which is what limits the resident warps of this stage.
    const float invw_norm = state[LF_SO_INVW];
    const int nlive = nfit - ((status & LF_ST_CUT_MASK) >> LF_ST_CUT_SHIFT);   // the reference's window (see lf_stage_prep)
    double* apd = wk.apd;
    lf_expand<MODEL, L, NV, NS, LANES>(cfg, w, sinc_w, th, apd);
    // line set-ups of all variants, built lane-parallel
    LfLineV* lines = wk.vlines;
    for (int e = w.lane; e < 13 * L; e += LANES) {
        int k = e / 13, v = e % 13;
        float pp, pl;
        lf_split(apd[L + k] - cfg.x_ref, pp, pl);
        float s = (float)apd[2 * L + k] + lf_var_ds(v);
        LfLineV x;
        lf_line_setup(MODEL, pp, pl - lf_var_dp(v), s, sinc_w, x.ln);
        x.aeff = (float)apd[k]; x.pad = 0.0f;
        x.ck = 1.0f; x.sk = 0.0f;
        if (MODEL != LF_GAUSSIAN) lf_phase(((double)x.ln.pp + (double)x.ln.dp) * inv_w, x.sk, x.ck);
        x.a2 = x.ln.a * x.ln.a;
        x.neae = -x.ln.ea * x.ln.ierf;
        lines[e] = x;
    }
    w.sync();
    const float iw = 1.0f / sinc_w;
    const float cont = th[D::IC];
    float gram[SLOTS];
    int gi[SLOTS], gj[SLOTS];                            // (row, column) of the Gram entries this lane owns
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        gram[s] = 0.0f;
        int e = s * LANES + w.lane, i = 0, rem = e < NE ? e : 0;
        while (rem >= PF - i) { rem -= PF - i; ++i; }
        gi[s] = i; gj[s] = i + rem;
    }
    // per-line second-difference sums: [5][L] per lane, kept in per-warp memory (rolled line loop)
    float* sums = wk.gt + 13 * LANES;                  // [5 L][LANES]
    for (int e = 0; e < 5 * L; ++e) sums[e * LANES + w.lane] = 0.0f;
    float* tile = wk.tile;
    float* gvt = wk.gt;                                // [13][LANES]
    float y_ahead = 0.0f;
    if (w.lane < nfit) {
        const int i = cfg.fit_lo + w.lane;
        y_ahead = gspec[i];
        if (cfg.bkg) y_ahead -= cfg.bkg[i];
    }
    for (int n0 = 0; n0 < nfit; n0 += LANES) {
        const int n = n0 + w.lane;
        const bool inwin = n < nfit;
        const bool live = n < nlive && (y_ahead == y_ahead);   // NaN samples are not part of the objective (LuciBase.py:358-360)
        const float x = inwin ? xo[n] : 0.0f;
        const float yn = live ? y_ahead * invw_norm : 0.0f;
        if (n + LANES < nfit) {
            const int i = cfg.fit_lo + n + LANES;
            y_ahead = gspec[i];
            if (cfg.bkg) y_ahead -= cfg.bkg[i];
        }
        float Sx = 0.0f, Cx = 1.0f;
        if (MODEL != LF_GAUSSIAN && inwin) lf_phase((double)x * inv_w, Sx, Cx);
        float* trow = tile + w.lane;                       // D_i of this lane's sample: trow[i * TS]
        // pass 1: model value at the solution (residual r0) and the unit profiles g0
        float m = cont;
#pragma unroll 1
        for (int k = 0; k < L; ++k) {
            const LfLineV& X = lines[13 * k];
            const float dx = lf_dx(X.ln, x);
            int hint = LF_HINT_LANE;
            if (MODEL == LF_SINCGAUSS) hint = lf_region_hint(w, X.ln, dx);
            float s = 0.0f, c = 1.0f;
            if (MODEL != LF_GAUSSIAN) { s = Sx * X.ck - Cx * X.sk; c = Cx * X.ck + Sx * X.sk; }
            float g, gp, gs;
            lf_profile_sc<MODEL>(X.ln, dx, s, c, iw, hint, g, gp, gs);
            trow[(3 * k) * TS] = live ? g : 0.0f;
            m += X.aeff * g;
        }
        const float r0 = yn - m;
        // pass 2: the 12 shifted variants of every line
#pragma unroll 1
        for (int k = 0; k < L; ++k) {
            const float Ak = lines[13 * k].aeff;
            const float g0 = trow[(3 * k) * TS];
            gvt[w.lane] = g0;
            // The 12 shifted variants of a line differ by at most 2 delta in centre and width, so one conservative test
            // (closest centre, widest / narrowest width) tells whether ALL of them are in the wing for this chunk.  Then
            // they run through straight-line code without a vote or a branch per variant, which lets the scheduler
            // overlap the dependent chains of neighbouring variants.
            bool wing_all = false;
            if (MODEL == LF_SINCGAUSS) {
                const LfLine& B = lines[13 * k].ln;
                const float adx = fmaxf(fabsf(lf_dx(B, x)) - 2.0f * dl, 0.0f);
                const float bmin = adx * LF_RCP(LF_SQRT2_F * (B.sig + 2.0f * dl));
                const float smin = fmaxf(B.sig - 2.0f * dl, 0.02f * sinc_w * (LF_SQRT2_F / LF_PI_F));
                const float amin = smin * (LF_PI_F / LF_SQRT2_F) * iw;
                wing_all = w.all(!live || (bmin * bmin >= 40.0f && bmin * bmin + amin * amin >= LF_W_R2_ASY));
            }
            if (wing_all) {
                // one full evaluation of w(z) per sigma value; the shifted-centre siblings of a sigma value are a
                // second-order Taylor step in z with their own phase rotation (lf_wing_w)
                const LfLineV* Lk = lines + 13 * k;
                auto rotate = [&](const LfLineV& X, float& s, float& c) { s = Sx * X.ck - Cx * X.sk; c = Cx * X.ck + Sx * X.sk; };
                auto group = [&](int vb, int v1, int v2, int v3, int v4) {      // base variant, siblings at -/+ d [, -/+ 2d]
                    const LfLineV& B = Lk[vb];
                    const float b = lf_dx(B.ln, x) * B.ln.is2s;
                    float wr, wi, d1r, d1i, d2r, d2i;
                    lf_wing_w(b, B.ln.a, b * b + B.a2, wr, wi);
                    lf_wing_w_taylor(b, B.ln.a, wr, wi, d1r, d1i, d2r, d2i);
                    float s, c;
                    if (vb != 0) { rotate(B, s, c); gvt[vb * LANES + w.lane] = B.neae * (c * wr + s * wi); }
                    const int sib[4] = {v1, v2, v3, v4};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int v = sib[q];
                        if (v < 0) continue;
                        const float dz = -lf_var_dp(v) * B.ln.is2s;          // dx_v = dx + dp_v: z_v = z - dp_v / (sqrt2 sigma)
                        const float X_ = wr + dz * (d1r + dz * d2r), Y_ = wi + dz * (d1i + dz * d2i);
                        rotate(Lk[v], s, c);
                        gvt[v * LANES + w.lane] = B.neae * (c * X_ + s * Y_);
                    }
                };
                auto single = [&](int v) {
                    const LfLineV& X = Lk[v];
                    const float b = lf_dx(X.ln, x) * X.ln.is2s;
                    float wr, wi, s, c;
                    lf_wing_w(b, X.ln.a, b * b + X.a2, wr, wi);
                    rotate(X, s, c);
                    gvt[v * LANES + w.lane] = X.neae * (c * wr + s * wi);
                };
                group(0, 1, 2, 5, 6);
                group(3, 9, 11, -1, -1);
                group(4, 10, 12, -1, -1);
                single(7);
                single(8);
            } else
#pragma unroll 1
            for (int v = 1; v < 13; ++v) {
                float gv;
                if (MODEL == LF_SINC && lf_var_ds(v) != 0.0f) {
                    gv = (lf_var_dp(v) == 0.0f) ? g0 : 0.0f;            // filled from variants 1 / 2 below
                } else {
                    const LfLineV& X = lines[13 * k + v];
                    const float dx = lf_dx(X.ln, x);
                    float s = 0.0f, c = 1.0f;
                    if (MODEL != LF_GAUSSIAN) { s = Sx * X.ck - Cx * X.sk; c = Cx * X.ck + Sx * X.sk; }
                    bool wing = false;
                    if (MODEL == LF_SINCGAUSS) {
                        const float b = dx * X.ln.is2s;
                        const float r2 = b * b + X.a2;
                        wing = w.all(b * b >= 40.0f && r2 >= LF_W_R2_ASY);
                        if (wing) {                                 // value-only wing form (see lf_eval)
                            float T, T2, E2;
                            lf_wing_terms(b, X.ln.a, b * b, X.a2, r2, s, c, T, T2, E2);
                            gv = X.neae * T;
                        }
                    }
                    if (!wing) {
                        int hint = LF_HINT_LANE;
                        if (MODEL == LF_SINCGAUSS) hint = lf_region_hint(w, X.ln, dx);
                        float gp, gs;
                        lf_profile_sc<MODEL>(X.ln, dx, s, c, iw, hint, gv, gp, gs);
                    }
                }
                gvt[v * LANES + w.lane] = gv;
            }
            float gv[13];
#pragma unroll
            for (int v = 0; v < 13; ++v) gv[v] = gvt[v * LANES + w.lane];
            if (MODEL == LF_SINC) { gv[9] = gv[10] = gv[1]; gv[11] = gv[12] = gv[2]; }
            const float Dp = Ak * (gv[1] - gv[2]) * (0.5f / dl);
            const float Ds = Ak * (gv[3] - gv[4]) * (0.5f / dl);
            trow[(3 * k + 1) * TS] = live ? Dp : 0.0f;
            trow[(3 * k + 2) * TS] = live ? Ds : 0.0f;
            if (live) {
                const float base = Ak * g0;
                float* sk_ = sums + (5 * k) * LANES + w.lane;
                // (A, p): D^{s,t} = (A + s d) g(p + t d) - A g0
                {
                    float a = (Ak + dl) * gv[1] - base, b = (Ak + dl) * gv[2] - base;
                    float c = (Ak - dl) * gv[1] - base, d = (Ak - dl) * gv[2] - base;
                    sk_[0] += a * a - b * b - c * c + d * d - 2.0f * r0 * (a - b - c + d);
                }
                {
                    float a = (Ak + dl) * gv[3] - base, b = (Ak + dl) * gv[4] - base;
                    float c = (Ak - dl) * gv[3] - base, d = (Ak - dl) * gv[4] - base;
                    sk_[LANES] += a * a - b * b - c * c + d * d - 2.0f * r0 * (a - b - c + d);
                }
                {
                    float a = Ak * (gv[5] - g0), b = Ak * (gv[6] - g0);
                    sk_[2 * LANES] += a * a + b * b - 2.0f * r0 * (a + b);
                }
                {
                    float a = Ak * (gv[7] - g0), b = Ak * (gv[8] - g0);
                    sk_[3 * LANES] += a * a + b * b - 2.0f * r0 * (a + b);
                }
                {
                    float a = Ak * (gv[9] - g0), b = Ak * (gv[10] - g0);
                    float c = Ak * (gv[11] - g0), d = Ak * (gv[12] - g0);
                    sk_[4 * LANES] += a * a - b * b - c * c + d * d - 2.0f * r0 * (a - b - c + d);
                }
            }
        }
        trow[(3 * L) * TS] = live ? 1.0f : 0.0f;
        w.sync();
        // Gram matrix of the D vectors: entry e = (i, j), i <= j, owned by lane e % LANES
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            if (s * LANES + w.lane < NE) {
                const float* ti = tile + gi[s] * TS;
                const float* tj = tile + gj[s] * TS;
                float acc = 0.0f;
#if defined(__CUDA_ARCH__)
                if (LANES == 32) {                          // rows are 16-byte aligned: four samples per load
#pragma unroll
                    for (int q = 0; q < LANES / 4; ++q) {
                        const float4 a = *reinterpret_cast<const float4*>(ti + 4 * q);
                        const float4 b = *reinterpret_cast<const float4*>(tj + 4 * q);
                        acc += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
                    }
                } else
#endif
                for (int t = 0; t < LANES; ++t) acc += ti[t] * tj[t];
                gram[s] += acc;
            }
        }
        w.sync();
    }
    // assemble H (double) in scratch
    double noise2 = noise * noise;
    if (!(noise2 == noise2)) noise2 = 1e-2;                       // LuciFit.py:514-515
    double* Hm = wk.hess;
    w.sync();
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        if (s * LANES + w.lane < NE) {
            const int i = gi[s], j = gj[s];
            double v = (double)gram[s] / noise2;
            Hm[i * PF + j] = v; Hm[j * PF + i] = v;
        }
    }
    w.sync();
    const double f8 = 1.0 / (8.0 * (double)dl * (double)dl * noise2);
#pragma unroll 1
    for (int k = 0; k < L; ++k) {
        const float* sk_ = sums + (5 * k) * LANES + w.lane;
        double vAp = (double)w.sum(sk_[0]) * f8, vAs = (double)w.sum(sk_[LANES]) * f8;
        double vpp = (double)w.sum(sk_[2 * LANES]) * f8, vss = (double)w.sum(sk_[3 * LANES]) * f8, vps = (double)w.sum(sk_[4 * LANES]) * f8;
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
    double* uv = Hm + PF * PF;                                     // PF doubles, then PF ints of pivots
    int* piv = reinterpret_cast<int*>(uv + PF);
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
        // eliminate column c from all other rows at once: the lanes share the PF^2 elements (not one row after the other
        // with PF lanes busy); column c is stashed first because its elements are overwritten by their owners
        for (int r = w.lane; r < PF; r += LANES) uv[r] = Hm[r * PF + c];
        w.sync();
        for (int e = w.lane; e < PF * PF; e += LANES) {
            const int r = e / PF, j = e - r * PF;
            if (r != c) Hm[e] = (j == c ? 0.0 : Hm[e]) - uv[r] * Hm[c * PF + j];
        }
        w.sync();
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

// ------------------------------------------------------------------------------------ per-warp scratch layout
// Byte offsets inside one warp's scratch area, per stage (computed on the host by lf_make_plan; the test harness
// uses the same layout with LANES = 1).  Offset 0 is `spec` (prep) or `yw` (lm / unc / out).
struct LfLayout {
    int lanes;
    int npad;                                              // continuum window padded to a power of two
    int prep_sortv, prep_bytes;
    int lm_sw, lm_cw, lm_lines, lm_tie, lm_gt, lm_hb, lm_thb, lm_bytes;
    int un_sw, un_cw, un_lines, un_apd, un_gt, un_tile, un_hess, un_bytes;
    int out_sw, out_cw, out_lines, out_apd, out_bytes;
};

enum { LF_STAGE_PREP = 0, LF_STAGE_LM = 1, LF_STAGE_UNC = 2, LF_STAGE_OUT = 3 };

LF_CORE LfWork lf_work_at(int stage, const LfLayout& ly, unsigned char* base)
{
    LfWork wk;
    wk.spec = reinterpret_cast<float*>(base);
    wk.yw = reinterpret_cast<float*>(base);
    wk.sortv = 0; wk.npad = ly.npad;
    wk.sw = 0; wk.cw = 0; wk.lines = 0; wk.vlines = 0; wk.tie = 0; wk.gt = 0; wk.hb = 0; wk.thb = 0; wk.apd = 0; wk.tile = 0; wk.hess = 0;
    if (stage == LF_STAGE_PREP) {
        wk.sortv = reinterpret_cast<float*>(base + ly.prep_sortv);
    } else if (stage == LF_STAGE_LM) {
        wk.sw = reinterpret_cast<float*>(base + ly.lm_sw);
        wk.cw = reinterpret_cast<float*>(base + ly.lm_cw);
        wk.lines = reinterpret_cast<LfLineX*>(base + ly.lm_lines);
        wk.tie = reinterpret_cast<LfTieX*>(base + ly.lm_tie);
        wk.gt = reinterpret_cast<float*>(base + ly.lm_gt);
        wk.hb = reinterpret_cast<float*>(base + ly.lm_hb);
        wk.thb = reinterpret_cast<float*>(base + ly.lm_thb);
    } else if (stage == LF_STAGE_UNC) {
        wk.sw = reinterpret_cast<float*>(base + ly.un_sw);
        wk.cw = reinterpret_cast<float*>(base + ly.un_cw);
        wk.vlines = reinterpret_cast<LfLineV*>(base + ly.un_lines);
        wk.apd = reinterpret_cast<double*>(base + ly.un_apd);
        wk.gt = reinterpret_cast<float*>(base + ly.un_gt);
        wk.tile = reinterpret_cast<float*>(base + ly.un_tile);
        wk.hess = reinterpret_cast<double*>(base + ly.un_hess);   // aliases the variant records (dead by then)
    } else {
        wk.sw = reinterpret_cast<float*>(base + ly.out_sw);
        wk.cw = reinterpret_cast<float*>(base + ly.out_cw);
        wk.lines = reinterpret_cast<LfLineX*>(base + ly.out_lines);
        wk.apd = reinterpret_cast<double*>(base + ly.out_apd);
    }
    return wk;
}
