"""TEST INFRASTRUCTURE ONLY -- a *tight* float64 solve of the reference's fitting problem.

The reference minimises its negative log-likelihood with SLSQP (``tol=1e-8``, 3-point finite-difference gradient,
``LUCI/LuciFit.py:676-681``) under the equality ties of ``LuciFit.py:523-598``.  SLSQP stops on a change of the
objective, so in a flat direction (the amplitude of a weak line) its answer is only as good as that test.  To decide
which of two nearby solutions (the reference's, the CUDA path's) is closer to the true optimum of the reference's own
objective, this module solves the same problem to machine precision:

  * the ties are eliminated analytically exactly as ``luci_b200/csrc/lucifit_core.cuh`` does (one velocity per
    velocity group, one broadening per sigma group, ``A_6548 = A_6583 G(sigma_6583) / (3 G(sigma_6548))``), which
    leaves a box-constrained non-linear least-squares problem in ``[A_free..., v_g..., beta_h..., continuum]``;
  * that problem is solved by ``scipy.optimize.least_squares`` (trust-region reflective, float64, analytic residuals,
    2-point Jacobian with tiny steps replaced by a complex-step-free central difference) with
    ``xtol = ftol = gtol = 1e-15``, started from a given solution (normally the reference's).

Nothing under ``luci_b200/`` imports this file.
"""
import numpy as np
from scipy import special as sps
from scipy.optimize import least_squares

from oracle import luci_oracle as O

C_KMS = float(O.SPEED_OF_LIGHT)


def _G(model, s, w):
    """Area factor of the [NII] tie (LuciFit.py:584-595)."""
    if model == 'sincgauss':
        return (np.sqrt(2 * np.pi) * s) / sps.erf(s / (np.sqrt(2) * w))
    return s


class TightProblem:
    """Tie-eliminated least-squares form of one ``OracleFit`` (non-freeze branch)."""

    def __init__(self, fit: O.OracleFit):
        self.f = fit
        self.L = fit.L
        self.model = fit.model
        self.lam = np.array([O.LINE_DICT[l] for l in fit.lines], dtype=float)
        self.vg = self._dense(fit.vel_rel)
        self.sg = self._dense(fit.sigma_rel)
        self.nv, self.ns = max(self.vg) + 1, max(self.sg) + 1
        self.i48 = self.i83 = -1
        # sinc with no prior broadening (the reference's ML_bool=False path): sigma = 0 makes the tie
        # A_83 sigma_83 / 3 = A_48 sigma_48 (LuciFit.py:587-589) vacuous, both amplitudes are free
        vacuous = fit.model == 'sinc' and fit.broad_ml == 0.0
        if 'NII6548' in fit.lines and 'NII6583' in fit.lines and fit.nii_cons is True and not vacuous:
            self.i48, self.i83 = fit.lines.index('NII6548'), fit.lines.index('NII6583')
        self.free_amp = [k for k in range(self.L) if k != self.i48]
        self.x = fit.axis_restricted.astype(float)
        self.y = fit.spectrum_restricted_norm.astype(float)
        self.w = fit.sinc_width

    @staticmethod
    def _dense(labels):
        seen, out = [], []
        for v in labels:
            if v not in seen:
                seen.append(v)
            out.append(seen.index(v))
        return out

    # reduced vector: [A_free | v | beta | c]
    def expand(self, red):
        """Reduced parameters -> the reference's [A, p, sigma]*L + [c] vector (normalised units)."""
        na = len(self.free_amp)
        A = np.zeros(self.L)
        A[self.free_amp] = red[:na]
        v = np.asarray(red[na:na + self.nv])[self.vg]
        beta = np.asarray(red[na + self.nv:na + self.nv + self.ns])[self.sg]
        p = 1e7 / (self.lam * (1.0 + v / C_KMS))
        s = p * beta / C_KMS
        if self.i48 >= 0:
            A[self.i48] = A[self.i83] * _G(self.model, s[self.i83], self.w) / (3.0 * _G(self.model, s[self.i48], self.w))
        full = np.zeros(3 * self.L + 1)
        full[0:3 * self.L:3], full[1:3 * self.L:3], full[2:3 * self.L:3] = A, p, s
        full[-1] = red[-1]
        return full

    def reduce(self, full):
        """The reference's vector (normalised units) -> reduced parameters (group values from each group's first line)."""
        A, p, s = full[0:3 * self.L:3], full[1:3 * self.L:3], full[2:3 * self.L:3]
        v = C_KMS * ((1e7 / p - self.lam) / self.lam)
        beta = np.abs(C_KMS * s / p)
        if self.model != 'sinc':                     # a zero-width spike (sigma = 0) is not a point of these models
            beta = np.maximum(beta, 1e-3)
        vfirst = [v[self.vg.index(g)] for g in range(self.nv)]
        bfirst = [beta[self.sg.index(h)] for h in range(self.ns)]
        return np.concatenate([A[self.free_amp], vfirst, bfirst, [full[-1]]])

    def residual(self, red):
        full = self.expand(red)
        with np.errstate(all='ignore'):
            m = O.model_sum(self.model, self.x, full, self.L, self.w) + full[-1]
        r = self.y - m
        return np.where(np.isfinite(r), r, 1e3)

    def objective(self, red):
        r = self.residual(red)
        return float(r @ r)

    def objective_full(self, full):
        with np.errstate(all='ignore'):
            m = O.model_sum(self.model, self.x, np.asarray(full, float), self.L, self.w) + full[-1]
        r = self.y - m
        return float(r @ r) if np.isfinite(r).all() else float('inf')

    def bounds(self):
        na = len(self.free_amp)
        lo = np.concatenate([np.full(na, -1e-8), np.full(self.nv, -np.inf), np.full(self.ns, 0.0 if self.model == 'sinc' else 1e-6), [-1e-8]])
        hi = np.concatenate([np.full(na, 1.1), np.full(self.nv, np.inf), np.full(self.ns, np.inf), [0.99]])
        # 0 <= sigma <= 10 cm^-1 binds the LAST line only (late-binding lambda, LuciFit.py:539-541): as a box on its
        # group's broadening (the centre moves by < 1e-3 relative over any plausible velocity)
        hi[na + self.nv + self.sg[-1]] = 10.0 * C_KMS / (1e7 / self.lam[-1])
        return lo, hi

    def solve(self, start_full, rounds=3):
        """Tight optimum started from ``start_full`` (reference layout, normalised units).  Returns
        (full vector, objective, scipy result)."""
        x0 = self.reduce(np.asarray(start_full, float))
        lo, hi = self.bounds()
        x0 = np.minimum(np.maximum(x0, lo + 1e-14), np.where(np.isfinite(hi), hi - 1e-14, hi))
        free = np.ones(len(x0), bool)
        if self.model == 'sinc':                     # sigma is not a parameter of the sinc profile (LuciFunctions.py:132-137)
            na = len(self.free_amp)
            free[na + self.nv:na + self.nv + self.ns] = False
        base = x0.copy()

        def fun(z):
            x = base.copy()
            x[free] = z
            return self.residual(x)

        res = None
        z0 = x0[free]
        for _ in range(rounds):                      # restarts re-estimate the Jacobian scaling at the new point
            res = least_squares(fun, z0, jac='3-point', bounds=(lo[free], hi[free]), method='trf', x_scale='jac',
                                xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=400)
            z0 = res.x
        x = base.copy()
        x[free] = res.x
        return self.expand(x), float(2.0 * res.cost), res


def derived(prob: TightProblem, full, scale):
    """velocities / broadenings / fluxes (physical units) of a normalised solution vector."""
    L = prob.L
    out = {'velocities': [], 'sigmas': [], 'fluxes': [], 'amplitudes': []}
    for k in range(L):
        A, p, s = full[3 * k] * scale, full[3 * k + 1], full[3 * k + 2]
        out['amplitudes'].append(A)
        out['velocities'].append(O.calculate_vel(p, prob.lam[k]))
        out['sigmas'].append(O.calculate_broad(p, s))
        out['fluxes'].append(O.calculate_flux(A, s, prob.model, prob.w))
    return {k: np.array(v) for k, v in out.items()}


def problem_from_golden(g, i, **kw):
    """The TightProblem of spaxel ``i`` of a golden fixture (tests/golden/*.npz as loaded by tests.common.load_golden)."""
    import warnings
    noml = bool(g['noml']) if 'noml' in g else False
    sky = g['cube'][i].astype(np.float64)
    good = ~np.isnan(sky)                                   # the row worker drops NaN samples (LuciBase.py:358-360)
    trans = g['trans'] if 'trans' in g else None
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        f = O.OracleFit(sky[good], np.asarray(g['axis32'])[good], g['model'], g['lines'],
                        [int(v) for v in g['vel_rel']], [int(v) for v in g['sigma_rel']],
                        trans_filter=None if trans is None else np.asarray(trans)[good],
                        theta=float(g['theta'][i]), delta_x=float(g['delta_x']), n_steps=int(g['n_steps']), zpd_index=0,
                        filter=g['filter'], nii_cons=True,
                        prior=None if noml else (float(np.ravel(g['v0'][i])[0]), float(np.ravel(g['b0'][i])[0])), **kw)
    return TightProblem(f)
