"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy float64 + SciPy SLSQP) of LUCI's per-spaxel fit path.

This is the *oracle* of the B200 implementation: it restates, function by function, what the reference
does for one spectrum and for one row of spaxels, so that the CUDA path can be checked against it on the
GPU box (where ``/root/reference`` does not exist).  It is pinned against the unmodified reference run
under ``oracle/ref_shim.py`` in the build container: ``oracle/make_golden.py`` writes the reference's
outputs to ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` checks this module against them.

Third-party arithmetic the reference delegates to (not under ``/root/reference``; SciPy is unpinned upstream,
``requirements.txt:3``; container has SciPy 1.18.1): ``scipy.optimize.minimize(method='SLSQP')``
(``LUCI/LuciFit.py:676-681``) and ``scipy.special.dawsn / erf`` (``LUCI/LuciFunctions.py:229-231``,
``LUCI/LuciFitParameters.py:139,178``).  The oracle calls the same SciPy routines.

Reference anchors (``/root/reference`` relative):
  Fit.__init__ ............... LUCI/LuciFit.py:49-173
  restrict_wavelength ........ LUCI/LuciFit.py:225-280
  calculate_noise ............ LUCI/LuciFit.py:282-331
  interpolate_spectrum ....... LUCI/LuciFit.py:364-380   (scale only; the CNN itself is out of scope)
  line_vals_estimate ......... LUCI/LuciFit.py:382-430
  cont_estimate .............. LUCI/LuciFit.py:432-486
  log_likelihood ............. LUCI/LuciFit.py:488-520
  constraints ................ LUCI/LuciFit.py:523-633
  calculate_params ........... LUCI/LuciFit.py:635-759
  fit (tail) ................. LUCI/LuciFit.py:761-835
  calc_chisquare ............. LUCI/LuciFit.py:1050-1069
  profiles ................... LUCI/LuciFunctions.py:40-310
  derived quantities ......... LUCI/LuciFitParameters.py:14-185
  hessianComp ................ LUCI/LuciUtility.py:409-434
  fit_calc (row worker) ...... LuciBase.py:258-406
  create_deep_image .......... LuciBase.py:147-215
  create_snr_map / SNR_calc .. LuciBase.py:1048-1191
  bin_cube_function .......... LUCI/LuciUtility.py:319-355
  LuciSim.Spectrum ........... LUCI/LuciSim.py:12-203
"""
import math

import numpy as np
from scipy import special as sps
from scipy.optimize import minimize

from oracle.astro_stats import mad_std, sigma_clip

SPEED_OF_LIGHT = 299792  # km/s (integer in the reference: LuciFit.py:26)
FWHM_COEFF = 2.0 * math.sqrt(2.0 * math.log(2.0))

LINE_DICT = {  # nm, LuciFit.py:90-99
    'Halpha': 656.280, 'NII6583': 658.341, 'NII6548': 654.803,
    'SII6716': 671.647, 'SII6731': 673.085, 'OII3726': 372.603,
    'OII3729': 372.882, 'OIII4959': 495.891, 'OIII5007': 500.684,
    'Hbeta': 486.133, 'OH': 649.873, 'HalphaC4': 807.881, 'NII6583C4': 810.417,
    'NII6548C4': 804.7,
    'OIII5007C2': 616.342, 'OIII4959C2': 610.441821, 'HbetaC2': 598.429723,
    'OII3729C1': 459.017742, 'OII3726C1': 458.674293, 'OI6364': 636.3776,
    'FeXIV5303': 530.286, 'NI5200': 520.026, 'FeVII5158': 515.89, 'HeII5411': 541.152,
    'FeXI6624': 662.43, 'NiXV6703': 670.332,
}

# (fit window, noise window, continuum window) in cm^-1: LuciFit.py:233-241, 291-299, 452-460
FILTER_WINDOWS = {
    'SN3': ((14750, 15400), (15600, 15800), (14850, 14950)),
    'SN2': ((19500, 20750), (18600, 19000), (19500, 19550)),
    'SN1': ((26000, 28000), (26000, 26200), (26000, 26250)),
}


def _argmin_abs(axis, value):
    return int(np.argmin(np.abs(np.asarray(axis) - value)))


# ------------------------------------------------------------------------------------------------ profiles
def gaussian_profile(x, A, p, s):
    """LuciFunctions.py:51-55"""
    return A * np.exp((-(x - p) ** 2) / (2 * s ** 2))


def sinc_profile(x, A, p, w):
    """LuciFunctions.py:132-137 (np.sinc is the normalised sinc; u == 0 -> A)"""
    return A * np.sinc((x - p) / w)


def sincgauss_profile(x, A, p, s, w):
    """LuciFunctions.py:219-232, Dawson-function form evaluated exactly as the reference does."""
    p2 = w / np.pi
    a = np.asarray(s / (np.sqrt(2) * p2), dtype=float)
    b = np.asarray((x - p) / (np.sqrt(2) * s), dtype=float)
    d1 = sps.dawsn(1j * a + b) * np.exp(2. * 1j * a * b)
    d2 = sps.dawsn(1j * a - b) * np.exp(-2. * 1j * a * b)
    d3 = 2. * sps.dawsn(1j * a)
    return np.real(A * (d1 + d2) / d3)


def model_sum(model, x, theta, L, w):
    """``*.evaluate``: sum of L profiles, parameters [A, p, sigma] per line (LuciFunctions.py:57-76 etc.)"""
    f = 0.0
    for k in range(L):
        A, p, s = theta[3 * k:3 * k + 3]
        if model == 'gaussian':
            f = f + gaussian_profile(x, A, p, s)
        elif model == 'sinc':
            f = f + sinc_profile(x, A, p, w)
        else:
            f = f + sincgauss_profile(x, A, p, s, w)
    return f


def model_plot(model, x, theta, L, w):
    """``*.plot``: like ``evaluate`` but every line centre snaps to the nearest axis sample
    (LuciFunctions.py:97-118, 181-203, 288-310); sincgauss maps NaN -> 0 (:309)."""
    f = 0.0
    for k in range(L):
        A, p, s = theta[3 * k:3 * k + 3]
        pos = x[_argmin_abs(x, p)]
        if model == 'gaussian':
            f = f + gaussian_profile(x, A, pos, s)
        elif model == 'sinc':
            f = f + sinc_profile(x, A, pos, w)
        else:
            f = f + np.nan_to_num(sincgauss_profile(x, A, pos, s, w))
    return f


# -------------------------------------------------------------------------------------- derived quantities
def calculate_vel(p, lam):
    """LuciFitParameters.py:15-37"""
    return SPEED_OF_LIGHT * ((1e7 / p - lam) / lam)


def calculate_vel_err(p, lam, u_p):
    """LuciFitParameters.py:41-57"""
    return SPEED_OF_LIGHT * u_p * (1e7 / (lam * p ** 2))


def calculate_broad(p, s):
    """LuciFitParameters.py:61-88"""
    return abs(SPEED_OF_LIGHT * s / p) if p > 0 else 0.0


def calculate_broad_err(p, s, u_p, u_s):
    """LuciFitParameters.py:93-115"""
    if p > 0 and s > 0:
        return (SPEED_OF_LIGHT * s / p) * np.sqrt((u_s / s) ** 2 + (u_p / p) ** 2)
    return 0.0


def calculate_flux(A, s, model, w):
    """LuciFitParameters.py:119-142"""
    if model == 'gaussian':
        return (1.20671 / FWHM_COEFF) * np.sqrt(2 * np.pi) * A * s
    if model == 'sinc':
        return np.pi * np.sqrt(np.pi) * A * w
    return (1.20671 / (np.pi * FWHM_COEFF)) * A * ((np.sqrt(2 * np.pi) * s) / (sps.erf(s / (np.sqrt(2) * w))))


def calculate_flux_err(A, s, u_A, u_s, model, w):
    """LuciFitParameters.py:146-185"""
    c0 = np.sqrt(2) * w
    flux = calculate_flux(A, s, model, w)
    with np.errstate(all='ignore'):
        if model in ('gaussian', 'sinc'):
            return flux * np.sqrt((u_A / A) ** 2 + (u_s / s) ** 2)
        return flux * np.sqrt((u_A / A) ** 2 + (u_s / s) ** 2 *
                              (sps.erf(s / c0) - (2 * np.pi) / (np.sqrt(np.pi) * c0) * np.exp(-s ** 2 / c0 ** 2)) ** 2)


def hessian_stencil(f, x0, delta=1e-1):
    """LuciUtility.py:409-434 (4-point central second differences, same delta for every parameter)."""
    x0 = np.array(x0, dtype=float)
    n = len(x0)
    H = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            ei = np.zeros(n); ei[i] = 1
            ej = np.zeros(n); ej[j] = 1
            H[i, j] = (f(x0 + delta * ei + delta * ej) - f(x0 + delta * ei - delta * ej)
                       - f(x0 - delta * ei + delta * ej) + f(x0 - delta * ei - delta * ej)) / (4 * delta * delta)
    return H


# ------------------------------------------------------------------------------------------- the fit itself
class OracleFit:
    """One spectrum.  Mirrors ``LUCI.LuciFit.Fit`` for the SLSQP (non-Bayesian, non-sky-line) branch.

    ``prior=(vel, broad)`` is what the reference's CNN would have produced (``vel_ml``, ``broad_ml``,
    LuciFit.py:356-361); with ``ml_path=True`` the scale follows ``interpolate_spectrum`` (:374-375), else
    ``fit():782``.  ``initial_values=[v, b]`` selects the reference's *freeze* branch (:159-162, 688-709).
    """

    def __init__(self, spectrum, axis, model_type, lines, vel_rel, sigma_rel, trans_filter=None, theta=0.0,
                 delta_x=2943, n_steps=842, zpd_index=169, filter='SN3', uncertainty_bool=False, nii_cons=True,
                 initial_values=(False,), spec_min=None, spec_max=None, obj_redshift=0.0, n_stoch=1,
                 prior=None, ml_path=True, rng=None, group_prior=None):
        if model_type == 'gauss':
            model_type = 'gaussian'
        if model_type not in ('gaussian', 'sinc', 'sincgauss'):
            raise Exception('Please submit a fitting function name in the available list')
        if not set(lines).issubset(LINE_DICT):
            raise Exception('Please submit a line name in the available list')
        if len(vel_rel) != len(lines) or len(sigma_rel) != len(lines):
            raise Exception('vel_rel / sigma_rel must have one entry per line')
        self.model = model_type
        self.lines = list(lines)
        self.L = len(lines)
        self.vel_rel = list(vel_rel)
        self.sigma_rel = list(sigma_rel)
        self.nii_cons = nii_cons
        self.filter = filter
        self.uncertainty_bool = uncertainty_bool
        self.n_stoch = n_stoch
        self.rng = rng if rng is not None else np.random
        spectrum = np.asarray(spectrum, dtype=float)
        self.spectrum = spectrum
        self.spectrum_normalized = spectrum / np.max(spectrum)                      # :110-112
        self.axis = np.asarray(axis) * (1 + obj_redshift)                            # :113
        if trans_filter is not None:                                                 # :124-125, 197-207
            t = np.asarray(trans_filter, dtype=float)
            self.spectrum = np.where(t > 0.5, spectrum / np.where(t > 0.5, t, 1.0), spectrum)
        if spec_min is None or spec_max is None:
            (spec_min, spec_max) = FILTER_WINDOWS[filter][0]
        self.spec_min, self.spec_max = spec_min, spec_max
        self.lo = _argmin_abs(self.axis, spec_min)                                   # :273-279
        self.hi = _argmin_abs(self.axis, spec_max)
        self.spectrum_restricted = self.spectrum_normalized[self.lo:self.hi]
        self.axis_restricted = self.axis[self.lo:self.hi]
        self.spectrum_restricted_norm = self.spectrum_restricted / np.max(self.spectrum_restricted)
        self.cos_theta = float(np.abs(np.cos(np.deg2rad(theta))))                    # :132
        self.correction_factor = 1 / self.cos_theta                                 # :214-215
        self.axis_step = self.correction_factor / (2 * float(delta_x) * (int(n_steps) - int(zpd_index))) * 1e7
        nlo = _argmin_abs(self.axis, FILTER_WINDOWS[filter][1][0])                   # :327-331
        nhi = _argmin_abs(self.axis, FILTER_WINDOWS[filter][1][1])
        with np.errstate(all='ignore'):
            self.noise = np.nanstd(self.spectrum_normalized[nlo:nhi])
        mpd = self.cos_theta * float(delta_x) * (int(n_steps) - int(zpd_index)) / 1e7  # :222-223
        self.sinc_width = 1 / (2 * mpd)
        self.vel_ml, self.broad_ml = (0.0, 0.0) if prior is None else (float(prior[0]), float(prior[1]))
        # group_prior = (velocity per velocity group, broadening per sigma group), groups in order of first appearance:
        # the per-group start of oracle/make_golden.py's ``per_group`` fixtures (line_vals_estimate wrapped so that
        # vel_ml / broad_ml are those of the line's group; not expressible through the reference's CNN)
        self.group_prior = group_prior
        self.ml_path = bool(ml_path) and prior is not None
        self.initial_values = list(initial_values)
        self.freeze = self.initial_values[0] is not False
        self.fit_sol = np.zeros(3 * self.L + 1)
        self.uncertainties = np.zeros(3 * self.L + 1)
        self.spectrum_scale = 0.0
        self.fit_vector = None
        self.slsqp = None

    # -- a7
    def line_vals_estimate(self, line, k=None):
        lam = LINE_DICT[line]
        if self.group_prior is not None and k is not None:
            dense = lambda rel: list(dict.fromkeys(rel)).index(rel[k])
            self.vel_ml = float(self.group_prior[0][dense(self.vel_rel)])
            self.broad_ml = float(self.group_prior[1][dense(self.sigma_rel)])
        if self.freeze:
            self.vel_ml, self.broad_ml = self.initial_values[0], self.initial_values[1]
        if np.isnan(self.broad_ml):
            self.broad_ml = 1.0
        if np.isnan(self.vel_ml):
            self.vel_ml = 0.0
        pos = 1e7 / ((self.vel_ml / SPEED_OF_LIGHT) * lam + lam)
        ind = _argmin_abs(self.axis, pos)
        sn = self.spectrum_normalized
        try:
            amp = np.max([sn[ind + d] for d in range(-5, 6)])   # negative indices wrap, like the reference
        except IndexError:
            amp = sn[ind]
        self.broad_ml = np.abs(self.broad_ml)
        return amp, pos, (pos * self.broad_ml) / SPEED_OF_LIGHT

    # -- a8
    def cont_estimate(self, sigma_level=3):
        lo = _argmin_abs(self.axis, FILTER_WINDOWS[self.filter][2][0])
        hi = _argmin_abs(self.axis, FILTER_WINDOWS[self.filter][2][1])
        win = self.spectrum_normalized[lo:hi]
        clipped = sigma_clip(win, sigma=sigma_level, masked=False, copy=False, maxiters=3, stdfunc=mad_std)
        if len(clipped) < 1:
            clipped = win
        return np.nanmedian(clipped)

    # -- a10
    def log_likelihood(self, theta):
        if self.freeze:
            full = np.zeros(3 * self.L)
            for k in range(self.L):
                full[3 * k] = theta[k]
                full[3 * k + 1] = self._frozen_pos[k]
                full[3 * k + 2] = self._frozen_sig[k]
            model = model_sum(self.model, self.axis_restricted, full, self.L, self.sinc_width)
        else:
            model = model_sum(self.model, self.axis_restricted, theta, self.L, self.sinc_width)
        model = model + np.real(theta[-1])
        sigma2 = np.real(self.noise ** 2)
        if np.isnan(sigma2):
            sigma2 = 1e-2
        res = -0.5 * np.nansum((self.spectrum_restricted_norm - model) ** 2 / sigma2) + np.log(2 * np.pi * sigma2)
        return -1e44 if np.isnan(res) else res

    # -- a9
    def constraints(self):
        C = SPEED_OF_LIGHT
        cons = []
        for g in np.unique(self.sigma_rel):                                           # :523-538
            inds = [i for i, e in enumerate(self.sigma_rel) if e == g]
            for j in inds[1:]:
                cons.append({'type': 'eq', 'fun': lambda x, j=j, i0=inds[0]:
                             (C * x[3 * i0 + 2]) / x[3 * i0 + 1] - (C * x[3 * j + 2]) / x[3 * j + 1]})
        last = self.L - 1                                                            # late binding, :539-541
        for _ in range(self.L):
            cons.append({'type': 'ineq', 'fun': lambda x: x[3 * last + 2]})
            cons.append({'type': 'ineq', 'fun': lambda x: -x[3 * last + 2] + 10})
        for g in np.unique(self.vel_rel):                                             # :545-568
            inds = [i for i, e in enumerate(self.vel_rel) if e == g]
            for j in inds[1:]:
                lj, l0 = LINE_DICT[self.lines[j]], LINE_DICT[self.lines[inds[0]]]
                cons.append({'type': 'eq', 'fun': lambda x, j=j, i0=inds[0], lj=lj, l0=l0:
                             C * ((1e7 / x[3 * j + 1] - lj) / lj) - C * ((1e7 / x[3 * i0 + 1] - l0) / l0)})
        for j in range(1, self.L):                                                    # :600-612
            cons.append({'type': 'ineq', 'fun': lambda x, j=j: x[3 * j + 1] + x[1] + 1})
        for j in range(self.L):                                                       # :614-633
            cons.append({'type': 'ineq', 'fun': lambda x, j=j: -x[3 * j] + 1.1})
            cons.append({'type': 'ineq', 'fun': lambda x, j=j: x[3 * j] + 1e-8})
        cons.append({'type': 'ineq', 'fun': lambda x: -x[-1] + 0.99})
        cons.append({'type': 'ineq', 'fun': lambda x: x[-1] + 1.e-8})
        if 'NII6548' in self.lines and 'NII6583' in self.lines and self.nii_cons is True:   # :570-598
            i48 = self.lines.index('NII6548')
            i83 = self.lines.index('NII6583')
            w = self.sinc_width
            if self.model in ('gaussian', 'sinc'):
                fn = lambda x: (1 / 3) * (x[3 * i83] * x[3 * i83 + 2]) - x[3 * i48] * x[3 * i48 + 2]
            else:
                g = lambda s: (np.sqrt(2 * np.pi) * s) / sps.erf(s / (np.sqrt(2) * w))
                fn = lambda x: (1 / 3) * (x[3 * i83] * g(x[3 * i83 + 2])) - x[3 * i48] * g(x[3 * i48 + 2])
            cons.append({'type': 'eq', 'fun': fn})
        return cons

    # -- a12
    def calculate_params(self):
        nll = lambda th: -self.log_likelihood(th)
        L = self.L
        best_fit, best_loss = None, 1e46
        positions = np.ones(L)
        sigmas = np.ones(L)
        if not self.freeze:
            cons = self.constraints()
            cont_est = self.cont_estimate(sigma_level=1)
            for st in range(self.n_stoch):
                initial = np.ones(3 * L + 1)
                initial[-1] = cont_est
                for k in range(L):
                    amp, pos, sig = self.line_vals_estimate(self.lines[k], k)
                    initial[3 * k] = amp - initial[-1]
                    if st == 0:
                        initial[3 * k + 1] = pos
                        initial[3 * k + 2] = sig
                    else:
                        initial[3 * k + 1] = self.rng.normal(pos, np.abs((1e7 / float(LINE_DICT[self.lines[k]]) - pos)) / 50)
                        initial[3 * k + 2] = self.rng.normal(sig, np.abs(sig / 10))
                if st == 0:
                    self.initial = initial.copy()
                soln = minimize(nll, initial, method='SLSQP', options={'disp': False, 'maxiter': 500},
                                tol=1e-8, jac="3-point", args=(), constraints=cons)
                if st == 0 or soln.fun < best_loss:
                    best_loss, best_fit = soln.fun, soln.x
                    self.slsqp = soln
        else:
            initial = np.ones(L + 1)
            initial[-1] = self.cont_estimate(sigma_level=3)
            for k in range(L):
                amp, pos, sig = self.line_vals_estimate(self.lines[k], k)
                initial[k] = amp - initial[-1]
                positions[k], sigmas[k] = pos, sig
            self._frozen_pos, self._frozen_sig = positions, sigmas
            self.initial = initial.copy()
            soln = minimize(nll, initial, method='SLSQP', options={'disp': False, 'maxiter': 30},
                            tol=1e-8, jac="3-point", args=())
            best_loss, best_fit = soln.fun, soln.x
            self.slsqp = soln
        parameters = np.array(best_fit, dtype=float)
        if self.uncertainty_bool:                                                     # :713-722
            if self.freeze:
                # the reference then indexes an (L+1)-vector as if it had 3L+1 entries (LuciFit.py:818-822)
                # and raises IndexError for L >= 2; not restated.
                raise NotImplementedError('freeze + uncertainty_bool is broken in the reference')
            H = 0.5 * hessian_stencil(nll, parameters)
            try:
                cov = -np.linalg.inv(H)
            except np.linalg.LinAlgError:
                cov = -np.linalg.pinv(H)
            self.uncertainties = np.sqrt(np.abs(np.diagonal(cov)))
        for k in range(L):                                                            # :724-734
            if self.freeze:
                parameters[k] *= self.spectrum_scale
                self.uncertainties[k] *= self.spectrum_scale
            else:
                parameters[3 * k] *= self.spectrum_scale
                self.uncertainties[3 * k] *= self.spectrum_scale
        parameters[-1] *= self.spectrum_scale
        self.uncertainties[-1] *= self.spectrum_scale
        if self.freeze:
            full = np.zeros(3 * L + 1)
            for k in range(L):
                full[3 * k], full[3 * k + 1], full[3 * k + 2] = parameters[k], positions[k], sigmas[k]
            full[-1] = parameters[-1]
            parameters = full
        self.fit_sol = parameters
        self.fit_vector = model_plot(self.model, self.axis, parameters[:-1], L, self.sinc_width) + parameters[-1]

    # -- a14
    def fit(self):
        if self.ml_path and not self.freeze:                                         # :776-780, 374-375
            zeros = np.zeros_like(np.asarray(self.spectrum, dtype=float))
            zeros[self.lo:self.hi] = self.spectrum_restricted
            self.spectrum_scale = np.max(zeros * np.max(self.spectrum))
        else:
            self.spectrum_scale = np.max(self.spectrum)                              # :782
        self.calculate_params()
        z = self.spectrum[self.lo:self.hi] - self.fit_vector[self.lo:self.hi]        # :1065-1068
        chi2 = np.sum((z ** 2) / (self.noise * self.spectrum_scale))
        red = chi2 / (len(self.fit_vector[self.lo:self.hi]) - (3 * self.L + 1))
        red /= self.spectrum_scale                                                   # :801
        s, u, w = self.fit_sol, self.uncertainties, self.sinc_width
        out = {'fit_sol': s, 'fit_uncertainties': u, 'amplitudes': [], 'fluxes': [], 'flux_errors': [],
               'chi2': red, 'velocities': [], 'sigmas': [], 'vels_errors': [], 'sigmas_errors': [],
               'axis_step': self.axis_step, 'corr': self.correction_factor, 'continuum': s[-1],
               'continuum_error': u[-1], 'scale': self.spectrum_scale, 'fit_vector': self.fit_vector,
               'fit_axis': self.axis, 'vel_ml': self.vel_ml, 'broad_ml': self.broad_ml, 'noise': self.noise}
        for k, line in enumerate(self.lines):
            lam = LINE_DICT[line]
            A, p, sg = s[3 * k], s[3 * k + 1], s[3 * k + 2]
            out['amplitudes'].append(A)
            out['fluxes'].append(calculate_flux(A, sg, self.model, w))
            out['velocities'].append(calculate_vel(p, lam))
            out['sigmas'].append(calculate_broad(p, sg))
            out['vels_errors'].append(calculate_vel_err(p, lam, u[3 * k + 1]))
            out['sigmas_errors'].append(calculate_broad_err(p, sg, u[3 * k + 1], u[3 * k + 2]))
            out['flux_errors'].append(calculate_flux_err(A, sg, u[3 * k], u[3 * k + 2], self.model, w))
        return out


ROW_KEYS = ('amplitudes', 'fluxes', 'flux_errors', 'velocities', 'vels_errors', 'sigmas', 'sigmas_errors',
            'chi2', 'corr', 'axis_step', 'continuum', 'continuum_error')


def fit_calc(i, x_min, x_max, y_min, fit_function, lines, vel_rel, sigma_rel, cube_slice, spectrum_axis,
             transmission_interpolated, interferometer_theta, hdr_dict, step_nb, zpd_index, mask=None,
             uncertainty_bool=False, nii_cons=False, bkg=None, bkgType=None, binning=None, spec_min=None,
             spec_max=None, initial_values=(False,), obj_redshift=0.0, n_stoch=1, prior_maps=None, ml_path=True):
    """Row worker, LuciBase.py:258-406.  ``prior_maps=(vel0[ny,nx], broad0[ny,nx])`` replaces the CNN."""
    y_pix = y_min + i
    L = len(lines)
    out = [[] for _ in ROW_KEYS]
    for j in range(x_max - x_min):
        x_pix = x_min + j
        do_fit = True if mask is None else bool(mask[x_pix, y_pix])
        sky = np.array(cube_slice[x_pix, :], dtype=float)
        if bkgType == 'standard':                                                     # :325-330
            sky = sky - (bkg * binning ** 2 if binning else bkg)
        good = ~np.isnan(sky)
        sky, axis = sky[good], np.asarray(spectrum_axis)[good]
        if initial_values[0] is not False:
            iv = [initial_values[0][i][j], initial_values[1][i][j]]
        else:
            iv = list(initial_values)
        if len(sky) > 0 and do_fit:
            prior = None if prior_maps is None else (prior_maps[0][i][j], prior_maps[1][i][j])
            trans = None if transmission_interpolated is None else np.asarray(transmission_interpolated)[good]
            d = OracleFit(sky, axis, fit_function, lines, vel_rel, sigma_rel, trans_filter=trans,
                          theta=interferometer_theta[x_pix, y_pix], delta_x=hdr_dict['STEP'], n_steps=step_nb,
                          zpd_index=zpd_index, filter=hdr_dict['FILTER'], uncertainty_bool=uncertainty_bool,
                          nii_cons=nii_cons, initial_values=iv, spec_min=spec_min, spec_max=spec_max,
                          obj_redshift=obj_redshift, n_stoch=n_stoch, prior=prior, ml_path=ml_path).fit()
            for lst, key in zip(out, ROW_KEYS):
                lst.append(d[key])
        else:
            for lst, key in zip(out, ROW_KEYS):
                lst.append([0] * L if key in ROW_KEYS[:7] else 0)
    return (i,) + tuple(out)


# --------------------------------------------------------------------------------------------- reductions
def deep_image(cube, strict_reference=True):
    """LuciBase.py:171-178: nansum over z in 10 chunks of int(nx/10) x-rows; the trailing ``nx mod 10`` rows
    stay zero (reference quirk); returned transposed (ny, nx)."""
    nx = cube.shape[0]
    deep = np.zeros((cube.shape[0], cube.shape[1]))
    if strict_reference:
        step = int(nx / 10)
        for i in range(10):
            deep[step * i:step * (i + 1)] = np.nansum(cube[step * i:step * (i + 1)], axis=2)
    else:
        deep[:] = np.nansum(cube, axis=2)
    return deep.T


SNR_WINDOWS = {  # (flux window, noise window) LuciBase.py:1083-1101
    'SN3': ((15150, 15300), (14500, 14600)),
    'SN2': ((1e7 / 486, 1e7 / 482), (19000, 19500)),
    'SN2_OIII': ((1e7 / 505, 1e7 / 495), (19000, 19500)),
    'SN1': ((26550, 27550), (25700, 26300)),
}


def snr_spaxel(sky, flo, fhi, nlo, nhi, method):
    """LuciBase.py:1142-1172 for one spectrum."""
    with np.errstate(all='ignore'):
        flux = np.nansum(sky[flo:fhi])
        clipped = sigma_clip(sky, sigma=1, masked=False, copy=False, maxiters=10)
        try:
            cont = np.nanmin(clipped)
        except ValueError:
            cont = np.nanmin(sky)
        flux -= cont * (fhi - flo)
        out = sky[nlo:nhi]
        std_out = np.nanstd(out)
        if method == 1:
            signal = np.nanmax(sky) - np.nanmean(sky)
            snr = float(signal / np.sqrt(np.abs(std_out)))
            snr = 0 if snr < 0 else snr / (np.sqrt(np.nanmean(np.abs(sky))))
        else:
            snr = float(flux / std_out)
            snr = 0 if snr < 0 else snr
    return snr


def snr_map(cube, axis, filter_key, method=1):
    """LuciBase.py:1078-1179, returned (ny, nx) float32."""
    (f0, f1), (n0, n1) = SNR_WINDOWS[filter_key]
    flo, fhi = _argmin_abs(axis, f0), _argmin_abs(axis, f1)
    nlo, nhi = _argmin_abs(axis, n0), _argmin_abs(axis, n1)
    nx, ny = cube.shape[:2]
    snr = np.zeros((nx, ny), dtype=np.float32).T
    for y in range(ny):
        for x in range(nx):
            snr[y, x] = snr_spaxel(np.asarray(cube[x, y, :], dtype=float), flo, fhi, nlo, nhi, method)
    return snr


def bin_cube(cube, binning, x_min, x_max, y_min, y_max):
    """LuciUtility.py:319-355: b x b nansum, no division (header edits not restated)."""
    nxb = int((x_max - x_min) / binning)
    nyb = int((y_max - y_min) / binning)
    out = np.zeros((nxb, nyb, cube.shape[2]))
    for i in range(nxb):
        for j in range(nyb):
            blk = cube[x_min + int(i * binning):x_min + int((i + 1) * binning),
                       y_min + int(j * binning):y_min + int((j + 1) * binning), :]
            out[i, j] = np.nansum(np.nansum(blk, axis=0), axis=0)
    return out


def bin_mask(mask, binning, x_min, x_max, y_min, y_max):
    """LuciUtility.py:358-388 (any-true, then divided by b^2 -- truthiness is what fit_calc reads)."""
    nxb = int((x_max - x_min) / binning)
    nyb = int((y_max - y_min) / binning)
    out = np.zeros((nxb, nyb))
    for i in range(nxb):
        for j in range(nyb):
            blk = mask[x_min + int(i * binning):x_min + int((i + 1) * binning),
                       y_min + int(j * binning):y_min + int((j + 1) * binning)]
            out[i, j] = bool(np.any(blk))
    return out / (binning ** 2)


# ----------------------------------------------------------------------------------------------- LuciSim
SIM_FILTERS = {'SN3': (2943, 8), 'SN2': (1680, 6), 'SN1': (1647, 8), 'C1': (570, 2), 'C2': (1680, 5),
               'C3': (1778, 6), 'C4': (5272, 12)}   # LuciSim.py:53-73


def sim_axis(filter_, resolution, theta=11.96):
    """LuciSim.py:90-96, 169-175 -> (axis[n] float64, n_steps, delta_x, sinc_width)"""
    delta_x, order = SIM_FILTERS[filter_]
    temp = (order + 0.5) / (2 * delta_x)
    n = int(np.ceil(1.20671 * resolution / (2 * temp * delta_x)))
    corr = 1 / np.cos(np.deg2rad(theta))
    x_min = order / (2 * delta_x) * corr * 1e7
    x_max = (order + 1) / (2 * delta_x) * corr * 1e7
    step_ = x_max - x_min
    axis = np.array([x_min + j * step_ / n for j in range(n)])
    mpd = np.cos(np.deg2rad(theta)) * delta_x * n / 1e7
    return axis, n, delta_x, 1 / (2 * mpd)


def sim_spectrum(lines, fit_function, ampls, velocity, broadening, filter_, resolution, snr, redshift=0.0,
                 theta=11.96, noise=None, rng=None):
    """LuciSim.Spectrum.create_spectrum, LuciSim.py:159-203.  ``noise`` (unit-variance draws) may be supplied
    to make the result reproducible; otherwise drawn from ``rng`` (default: global numpy RNG like the reference)."""
    axis, n, delta_x, w = sim_axis(filter_, resolution, theta)
    corr = 1 / np.cos(np.deg2rad(theta))
    if not isinstance(velocity, (list, tuple, np.ndarray)):
        velocity = [velocity] * len(lines)
    if not isinstance(broadening, (list, tuple, np.ndarray)):
        broadening = [broadening] * len(lines)
    spectrum = np.zeros_like(axis)
    for k, line in enumerate(lines):
        lam_cm = (1e7 / LINE_DICT[line]) / (1 + redshift)
        pos = (velocity[k] / 3e5) * lam_cm + lam_cm
        pos = axis[_argmin_abs(axis, pos)]
        sigma = (broadening[k] * pos) / (3e5 * corr)
        if fit_function == 'gaussian':
            spectrum += gaussian_profile(axis, ampls[k], pos, sigma)
        elif fit_function == 'sinc':
            spectrum += sinc_profile(axis, ampls[k], pos, w)
        else:
            spectrum += sincgauss_profile(axis, ampls[k], pos, sigma, w)
    if noise is None:
        noise = (rng if rng is not None else np.random).normal(0.0, 1.0, spectrum.shape)
    if snr is not None:
        spectrum = spectrum + np.max(spectrum) * noise / snr
    return axis, spectrum
