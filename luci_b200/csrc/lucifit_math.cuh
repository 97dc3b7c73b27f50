// lucifit_math.cuh -- line-profile models of LUCI with analytic partial derivatives, FP32.
//
// Replaces (value only, float64, SciPy) LUCI/LuciFunctions.py:51-55 (Gaussian.function), :132-137
// (Sinc.function), :219-232 (SincGauss.function, Dawson form) and supplies the derivatives the reference
// obtains by finite differences inside SLSQP (LUCI/LuciFit.py:676-681, jac='3-point').
//
// The same source is compiled by nvcc for sm_100a (fast intrinsics) and by g++ for the CPU test harness
// under tests/hostemu (libm); the harness is test infrastructure, never loaded by the package.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define LF_HD __host__ __device__ __forceinline__
#else
#define LF_HD inline
#endif

#if defined(__CUDA_ARCH__)
// raw MUFU forms: the arguments here are never denormal / huge, so the range guards of __expf / __fdividef
// (5-8 extra issue slots each) buy nothing
__device__ __forceinline__ float lf_mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lf_mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#define LF_EXP(x) lf_mufu_ex2((x) * 1.4426950408889634f)
#define LF_RCP(x) lf_mufu_rcp(x)
#define LF_SIN(x) __sinf(x)
#define LF_COS(x) __cosf(x)
#define LF_RINT(x) rintf(x)
#define LF_RSQRT(x) rsqrtf(x)
// per-line constants of one model evaluation (lf_line_setup, lf_setup_lines: five lanes busy, a dozen divisions and two
// exponentials each): the MUFU forms (1 ulp / 2 ulp) scale a whole line's profile by 1e-7 at most
#define LF_SETUP_RCP(x) lf_mufu_rcp(x)
#define LF_SETUP_EXP(x) lf_mufu_ex2((x) * 1.4426950408889634f)
#else
#define LF_SETUP_RCP(x) (1.0f / (x))
#define LF_SETUP_EXP(x) expf(x)
#define LF_RSQRT(x) (1.0f / sqrtf(x))
#define LF_EXP(x) expf(x)
#define LF_RCP(x) (1.0f / (x))
#define LF_SIN(x) sinf(x)
#define LF_COS(x) cosf(x)
#define LF_RINT(x) rintf(x)
#endif

#ifndef LF_PHASE_MUFU
#define LF_PHASE_MUFU 1
#endif
#define LF_PI_F 3.14159265358979f
#define LF_ISQRTPI_F 0.564189583547756f   /* 1/sqrt(pi) */
#define LF_2ISQRTPI_F 1.128379167095513f  /* 2/sqrt(pi) */
#define LF_SQRT2_F 1.414213562373095f

enum { LF_GAUSSIAN = 0, LF_SINC = 1, LF_SINCGAUSS = 2 };

// sin(pi d), cos(pi d) with an exact range reduction (d is a channel offset, |d| < ~1e3).
LF_HD void lf_sincospi(float d, float& s, float& c)
{
    float r = d - 2.0f * LF_RINT(0.5f * d);   // r in [-1, 1], exact in FP32
    float t = LF_PI_F * r;
    s = LF_SIN(t);
    c = LF_COS(t);
}

// Faddeeva function w(z) = exp(-z^2) erfc(-iz) for z = x + i y, y >= 0, |error| ~ 2e-7 (FP32).
//   |z|^2 >= 36 : Laplace continued fraction, 5 levels, as a rational function of u = z^2
//                 (w ~ i/sqrt(pi) * z (u^2 - 7u + 8.25) / (u^3 - 7.5u^2 + 11.25u - 1.875));
//   |z|^2 >= 1e4: first three terms of the asymptotic series (avoids the FP32 overflow of |zD|^2 further out; the
//                 series is already good to 7e-8 from |z|^2 = 300 = LF_W_R2_ASY, which is where the fit kernels'
//                 warp-uniform wing path starts -- per lane the fraction is kept up to 1e4 so that a mixed chunk
//                 does not pay for a third divergent branch);
//   otherwise   : Weideman's N = 16 rational expansion (SIAM J. Numer. Anal. 31 (1994) 1497).
#define LF_W_R2_CF 36.0f
#define LF_W_R2_ASY 300.0f
#define LF_W_R2_ASY_LANE 1.0e4f
// Far region, split as w(z) = (i/sqrt pi) w1 + w2 with w1 = 1/z and w2 = O(z^-3) evaluated without
// cancellation; the split is what keeps the wing derivatives of the sincgauss profile accurate in FP32.
// For the 5-level fraction  w2 = (i/sqrt pi) (u^2/2 - 3u + 1.875) / (z D(u)),  D = u^3 - 7.5u^2 + 11.25u - 1.875.
LF_HD void lf_faddeeva_far_split(float x, float y, float r2, float& w1r, float& w1i, float& w2r, float& w2i)
{
    if (r2 >= LF_W_R2_ASY_LANE) {
        float inv = LF_RCP(r2);
        w1r = x * inv; w1i = -y * inv;                               // 1/z
        float sr = w1r * w1r - w1i * w1i, si = 2.0f * w1r * w1i;     // s = 1/z^2
        float fr = 0.5f * sr * (1.0f + 1.5f * sr) - 0.75f * si * si, fi = 0.5f * si + 1.5f * sr * si;   // s/2 + 3 s^2/4
        float qr = w1r * fr - w1i * fi, qi = w1r * fi + w1i * fr;
        w2r = -qi * LF_ISQRTPI_F;
        w2i = qr * LF_ISQRTPI_F;
        return;
    }
    float ur = x * x - y * y, ui = 2.0f * x * y;
    float er = ur - 7.5f;
    float dr = er * ur - ui * ui + 11.25f, di = er * ui + ui * ur;
    float Dr = dr * ur - di * ui - 1.875f, Di = dr * ui + di * ur;    // D(u)
    float Zr = x * Dr - y * Di, Zi = x * Di + y * Dr;                 // z D
    float inv = LF_RCP(Zr * Zr + Zi * Zi);
    float ir = Zr * inv, ii = -Zi * inv;                              // 1/(z D)
    w1r = Dr * ir - Di * ii; w1i = Dr * ii + Di * ir;                 // 1/z = D/(zD)
    float tr = 0.5f * ur - 3.0f;
    float nr = tr * ur - 0.5f * ui * ui + 1.875f, ni = tr * ui + 0.5f * ui * ur;   // u^2/2 - 3u + 1.875
    float qr = nr * ir - ni * ii, qi = nr * ii + ni * ir;
    w2r = -qi * LF_ISQRTPI_F;
    w2i = qr * LF_ISQRTPI_F;
}

LF_HD void lf_faddeeva_far(float x, float y, float r2, float& wr, float& wi)
{
    float w1r, w1i, w2r, w2i;
    lf_faddeeva_far_split(x, y, r2, w1r, w1i, w2r, w2i);
    wr = -LF_ISQRTPI_F * w1i + w2r;
    wi = LF_ISQRTPI_F * w1r + w2i;
}

LF_HD void lf_faddeeva_near(float x, float y, float& wr, float& wi)
{
    const float Lw = 3.363585661014858f;
    // den = L - i z = (L + y) - i x ; Z = (L + i z)/den = ((L - y) + i x)/den
    float dr = Lw + y, di = -x;
    float inv = LF_RCP(dr * dr + di * di);
    float ir = dr * inv, ii = -di * inv;                          // 1/den
    float nr = Lw - y, ni = x;
    float Zr = nr * ir - ni * ii, Zi = nr * ii + ni * ir;
    const float c[16] = {9.939322536e-07f, 3.981287575e-06f, -5.584233411e-06f, -2.734640462e-05f,
                         2.170986793e-05f, 2.107105640e-04f, 8.703158428e-05f, -1.527659740e-03f,
                         -3.881015189e-03f, 3.682567317e-03f, 5.182240243e-02f, 1.912417267e-01f,
                         4.692909009e-01f, 8.864478302e-01f, 1.362240822e+00f, 1.748395886e+00f};
    float pr = c[0], pi_ = 0.0f;
#pragma unroll
    for (int k = 1; k < 16; ++k) {
        float t = pr * Zr - pi_ * Zi + c[k];
        pi_ = pr * Zi + pi_ * Zr;
        pr = t;
    }
    float i2r = ir * ir - ii * ii, i2i = 2.0f * ir * ii;          // 1/den^2
    wr = 2.0f * (pr * i2r - pi_ * i2i) + LF_ISQRTPI_F * ir;
    wi = 2.0f * (pr * i2i + pi_ * i2r) + LF_ISQRTPI_F * ii;
}

LF_HD void lf_faddeeva(float x, float y, float& wr, float& wi)
{
    float r2 = x * x + y * y;
    if (r2 >= LF_W_R2_CF)
        lf_faddeeva_far(x, y, r2, wr, wi);
    else
        lf_faddeeva_near(x, y, wr, wi);
}

// sin(pi t), cos(pi t) for a phase t carried in double (t = offset / sinc width, |t| up to ~1e3 half periods): the
// range reduction is exact, the result is good to ~1e-7 absolute.  Used for the per-sample phase table and the
// per-line phase constants of the fit kernels (sin(pi (x - p)/w) = S_x C_p - C_x S_p).
LF_HD void lf_phase(double t, float& s, float& c)
{
    double r = t - 2.0 * rint(0.5 * t);               // r in [-1, 1]
#if defined(__CUDA_ARCH__)
#if LF_PHASE_MUFU
    const float a = LF_PI_F * (float)r;               // |a| <= pi: MUFU.SIN / MUFU.COS are good to 4e-7 absolute there
    s = __sinf(a); c = __cosf(a);
#else
    sincospif((float)r, &s, &c);
#endif
#else
    s = (float)sin(3.141592653589793 * r);
    c = (float)cos(3.141592653589793 * r);
#endif
}

// ---------------------------------------------------------------------------------------------------
// Per-line constants, refreshed once per model evaluation (warp-uniform).
struct LfLine {
    float pp;      // line centre minus x_ref, leading part              [cm^-1]
    float dp;      // small remainder: offset from centre is (x - pp) - dp (keeps FP32 resolution ~1e-6 cm^-1)
    float sig;     // gaussian sigma                                      [cm^-1]
    float isig;    // 1/sigma
    float is2s;    // 1/(sqrt(2) sigma)
    float a;       // sincgauss: sigma pi / (sqrt(2) w)
    float ea;      // exp(-a^2)
    float ierf;    // 1/erf(a)
    float q;       // (2/sqrt(pi)) exp(-a^2)/erf(a)
    float ia;      // 1/a
};

LF_HD void lf_line_setup(int model, float pp, float dp, float sig, float sinc_w, LfLine& ln)
{
    ln.pp = pp;
    ln.dp = dp;
    float s = fabsf(sig);                     // every profile is even in sigma
    if (model == LF_SINCGAUSS) s = fmaxf(s, 0.02f * sinc_w * (LF_SQRT2_F / LF_PI_F));   // a >= 0.02 (cancellation guard)
    else s = fmaxf(s, 1e-6f);
    ln.sig = s;
    ln.isig = LF_SETUP_RCP(s);
    ln.is2s = ln.isig * (1.0f / LF_SQRT2_F);
    if (model == LF_SINCGAUSS) {
        float a = s * (LF_PI_F / LF_SQRT2_F) * LF_SETUP_RCP(sinc_w);
        float ef = erff(a);
        ln.a = a;
        ln.ea = LF_SETUP_EXP(-a * a);
        ln.ierf = LF_SETUP_RCP(ef);
        ln.q = LF_2ISQRTPI_F * ln.ea * ln.ierf;
        ln.ia = LF_SETUP_RCP(a);
    } else {
        ln.a = 0.0f; ln.ea = 0.0f; ln.ierf = 1.0f; ln.q = 0.0f; ln.ia = 0.0f;
    }
}

// Unit-amplitude profile value g and its partials with respect to the line centre (g_p) and sigma (g_s)
// at offset dx = x - centre.
//   gaussian : g = exp(-b^2),                       b = dx/(sqrt2 sigma)
//   sinc     : g = sin(pi d)/(pi d),                d = dx/w               (sigma is not a parameter)
//   sincgauss: g = exp(-b^2) Re erf(a + i b)/erf(a) = [exp(-b^2) - exp(-a^2) Re(e^{-2iab} w(ia - b))]/erf(a)
//              dg/db = -2 b g + q sin(2ab),  dg/da = q (cos(2ab) - g),  q = (2/sqrt pi) exp(-a^2)/erf(a)
// which is algebraically the reference's Dawson expression (LuciFunctions.py:229-232) without its e^{+a^2}.
// `s`, `c` = sin, cos(pi dx / w): the caller supplies them (the fit kernels rotate a per-sample phase table by a
// per-line constant instead of spending two MUFU ops per line and sample); unused by the gaussian.
// `hint` (warp-uniform, from lf_region_hint in the fit kernels): LF_HINT_LANE every lane picks its own branch of w(z);
// LF_HINT_FAR all lanes are beyond b^2 = 40 (far branch, exp(-b^2) = 0); LF_HINT_NEAR all lanes take the Weideman
// branch -- a chunk that holds the core of a line would otherwise execute BOTH branches (the lanes of the core and the
// lanes 7+ samples away diverge), and the N = 16 expansion is good to 3e-7 relative out to |z| = 30 (value and d/dp to
// 1e-7 of the peak; d/dsigma to 1e-5 of its peak for a >= 1, which is why narrower lines keep the per-lane choice).
enum { LF_HINT_LANE = 0, LF_HINT_FAR = 1, LF_HINT_NEAR = 2 };
#define LF_NEAR_A_MIN 1.0f
#define LF_NEAR_R2_MAX 6400.0f

template <int MODEL>
LF_HD void lf_profile_sc(const LfLine& ln, float dx, float s, float c, float iw /* 1/w */, int hint,
                         float& g, float& g_p, float& g_s)
{
    const bool all_far = hint == LF_HINT_FAR;
    if (MODEL == LF_GAUSSIAN) {
        float b = dx * ln.is2s;
        float b2 = b * b;
        g = LF_EXP(-b2);
        g_p = g * 2.0f * b * ln.is2s;
        g_s = g * 2.0f * b2 * ln.isig;
    } else if (MODEL == LF_SINC) {
        float d = dx * iw;
        float ad = fabsf(d);
        float gd;                                     // dg/dd
        if (ad < 2e-2f) {
            float t = (LF_PI_F * d) * (LF_PI_F * d);
            g = 1.0f - t * (1.0f / 6.0f) * (1.0f - t * (1.0f / 20.0f));
            gd = -(LF_PI_F * LF_PI_F) * d * (1.0f / 3.0f) * (1.0f - t * (1.0f / 10.0f));
        } else {
            float id = LF_RCP(d);
            g = s * id * (1.0f / LF_PI_F);
            gd = (c - g) * id;
        }
        g_p = -gd * iw;
        g_s = 0.0f;
    } else {
        float b = dx * ln.is2s;                       // 2ab = pi dx / w is the angle of (s, c)
        const float a = ln.a;
        float r2 = b * b + a * a;
        bool far = all_far || (hint != LF_HINT_NEAR && r2 >= LF_W_R2_CF);
        float eb = all_far ? 0.0f : LF_EXP(-b * b);
        float gb, gal;                                // dg/db at fixed a ; dg/da at fixed 2ab
        if (far) {
            // wings: w = (i/sqrt pi)/z + w2;  T_b and T_alpha written so the O(b) terms cancel analytically
            float w1r, w1i, w2r, w2i;
            lf_faddeeva_far_split(-b, a, r2, w1r, w1i, w2r, w2i);
            float T2 = c * w2r + s * w2i;
            float T = LF_ISQRTPI_F * (s * w1r - c * w1i) + T2;
            float Tb = (LF_2ISQRTPI_F * a) * (c * w1r + s * w1i) - 2.0f * b * T2;
            float eae = ln.ea * ln.ierf;
            g = eb * ln.ierf - eae * T;
            gb = -2.0f * b * eb * ln.ierf - eae * Tb;
            gal = (2.0f * b * b * ln.ia + 2.0f * a) * eb * ln.ierf
                  - (2.0f * a + ln.q) * g - 2.0f * r2 * ln.ia * eae * T2;
        } else {
            float wr, wi;
            lf_faddeeva_near(-b, a, wr, wi);
            float T = c * wr + s * wi;
            g = (eb - ln.ea * T) * ln.ierf;
            gb = -2.0f * b * g + ln.q * s;
            gal = ln.q * (c - g) - gb * b * ln.ia;
        }
        g_p = -gb * ln.is2s;
        g_s = gal * a * ln.isig;
    }
}

// Stand-alone form: computes the phase itself (synthetic-cube generator, unit tests).
template <int MODEL>
LF_HD void lf_profile(const LfLine& ln, float dx, float piw /* pi/w */, float iw /* 1/w */, bool all_far,
                      float& g, float& g_p, float& g_s)
{
    float s = 0.0f, c = 1.0f;
    if (MODEL != LF_GAUSSIAN) lf_sincospi(dx * iw, s, c);
    lf_profile_sc<MODEL>(ln, dx, s, c, iw, all_far ? LF_HINT_FAR : LF_HINT_LANE, g, g_p, g_s);
}

// "far" test used to pick the warp-uniform wing path of the sincgauss evaluation: continued-fraction region
// of w(z) and exp(-b^2) below FP32 resolution of the result (b^2 > 40 -> exp(-b^2) < 4.3e-18).
LF_HD float lf_dx(const LfLine& ln, float x) { return (x - ln.pp) - ln.dp; }

// Split a double offset into the (pp, dp) pair of LfLine.
LF_HD void lf_split(double d, float& hi, float& lo) { hi = (float)d; lo = (float)(d - (double)hi); }

LF_HD bool lf_is_far(const LfLine& ln, float dx)
{
    float b = dx * ln.is2s;
    return b * b >= 40.0f;
}

// d ln G / d sigma of the area factor G that enters the [NII] doublet constraint (LuciFit.py:584-595):
//   gaussian, sinc: G = sigma;   sincgauss: G = sigma / erf(sigma/(sqrt2 w)).
LF_HD float lf_dlnG(int model, float sig, float sinc_w)
{
    if (model != LF_SINCGAUSS) return 1.0f / sig;
    float u = sig / (LF_SQRT2_F * sinc_w);
    return 1.0f / sig - LF_2ISQRTPI_F * expf(-u * u) / (erff(u) * LF_SQRT2_F * sinc_w);
}
LF_HD float lf_G(int model, float sig, float sinc_w)
{
    if (model != LF_SINCGAUSS) return sig;
    return sig / erff(sig / (LF_SQRT2_F * sinc_w));
}
