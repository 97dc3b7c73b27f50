"""The device fit pipeline (luci_b200/csrc/lucifit_core.cuh) compiled for the CPU with LANES = 1, against the
reference outputs in tests/golden/.  This exercises prep, the tie-eliminated LM, the FD-stencil uncertainties and
the derived outputs without a GPU; the GPU tier repeats the same comparison through the C ABI."""
import numpy as np
import pytest

from tests.common import (FIT_CASES, TOL_VEL_KMS, UNC_CASES, assert_fixture_parity, assert_uncertainty_parity,
                          config_from_golden, golden_priors, load_golden, optimum_report, parity_report, run_hostemu,
                          split_rows)


COMPONENT_OFFSETS = (-300.0, -200.0, -100.0, 100.0, 200.0, 300.0)   # luci_b200.luci.COMPONENT_OFFSETS


def multistart_double(fit, g):
    """The API's component search for a double-component fit given ONE shared prior (luci_b200/luci.py): the second
    velocity group is started at the prior + each offset, the fit with the lowest chi2 wins per spaxel.
    ``fit(cfg, v0 [n, 2], b0 [n, 2]) -> (rows, status)``."""
    L = len(g['lines'])
    cfg = config_from_golden(g, uncertainty=False, prior_groups=True)
    best = best_st = None
    for d in COMPONENT_OFFSETS:
        v = np.stack([g['v0'], g['v0'] + d], axis=1).astype(np.float32)
        b = np.stack([g['b0'], g['b0']], axis=1).astype(np.float32)
        rows, st = fit(cfg, v, b)
        if best is None:
            best, best_st = rows.copy(), st.copy()
        else:
            m = (((st >> 16) & 1) == 1) & (rows[:, 7 * L] < best[:, 7 * L])
            best[m], best_st[m] = rows[m], st[m]
    return best, best_st


@pytest.mark.parametrize('name', FIT_CASES)
def test_hostemu_matches_reference(hostemu, name):
    g = load_golden(name)
    if name == 'fit_sincgauss_double':
        # BASELINE config 4 started the only way the reference's API can (ONE prior pair for both components): an exactly
        # symmetric start.  The reference leaves it through the noise of its finite-difference gradient and lands on the
        # right answer in 2 of these 10 spectra ("the results are not always correct", docs/source/example_double_fit.rst);
        # the API here searches the second component on a velocity grid.  Every spaxel must be the reference's optimum
        # or a strictly lower converged minimum of the same objective.
        fit = lambda cfg, v, b: (lambda r: (r[0], r[3]))(run_hostemu(hostemu, cfg, g['cube'], g['theta'], v, b, want_unc=False))
        rows, st = multistart_double(fit, g)
        opt = optimum_report(rows, g)
        assert opt['ok'].all() and opt['lower'].sum() >= 5, (opt['same'], opt['lower'])
        assert (opt['dchi'] <= 1e-4).all()
        vel = split_rows(rows, 6)['velocities']
        assert (np.abs(vel[:, 5] - vel[:, 0] - 200.0) < 50.0).all()   # the simulated second component at +200 km/s (each snapped to a channel)
        return
    cfg = config_from_golden(g)
    v0, b0 = golden_priors(g)
    rows, sol, unc, st = run_hostemu(hostemu, cfg, g['cube'], g['theta'], v0, b0)
    rep, opt, direct = assert_fixture_parity(name, rows, st, g)
    # how often the direct comparison with the reference is possible at all (it is at its own optimum): documented floor
    floor = {'fit_sinc_noml': 0.25, 'fit_sinc_sn3': 0.8, 'fit_sincgauss_groups': 0.75, 'fit_sincgauss_groups_hi': 0.85}.get(name, 0.95)
    assert direct.mean() >= floor, (name, direct.mean())
    model = g['model']
    if model == 'sinc' and not bool(g['noml']):
        # sigma is not a parameter of the sinc model: the broadening map returns the prior (the reference returns
        # the prior plus an SLSQP feasibility-restoration drift of 0.1-20 %, which is not a property of the data)
        np.testing.assert_allclose(rep['q']['sigmas'], np.repeat(np.abs(g['b0'])[:, None], 5, axis=1), rtol=1e-6)
    # geometry outputs are exact formulas
    np.testing.assert_allclose(rep['q']['corr'], g['ref_corr'], rtol=1e-6)
    np.testing.assert_allclose(rep['q']['axis_step'], g['ref_axis_step'], rtol=1e-6)
    np.testing.assert_allclose(rep['q']['continuum'][direct], g['ref_continuum'][direct], rtol=2e-3)


@pytest.mark.parametrize('name', UNC_CASES)
def test_hostemu_uncertainties(hostemu, name):
    g = load_golden(name)
    cfg = config_from_golden(g, uncertainty=True)
    v0, b0 = golden_priors(g)
    rows, sol, unc, st = run_hostemu(hostemu, cfg, g['cube'], g['theta'], v0, b0)
    rep = parity_report(rows, g, g['model'])
    sel = rep['ok']                       # every spaxel where the two fits agree (all of them but the reference's run-aways)
    assert sel.mean() >= (0.85 if name == 'fit_sincgauss_groups_hi' else 1.0)
    assert_uncertainty_parity(name, rows, unc, g, sel)


def test_hostemu_eval_count_is_small(hostemu):
    g = load_golden('fit_sincgauss_sn3')
    cfg = config_from_golden(g, uncertainty=False)
    rows, sol, unc, st = run_hostemu(hostemu, cfg, g['cube'], g['theta'], g['v0'], g['b0'])
    evals = (st >> 8) & 0xff
    assert evals.mean() < 10 and evals.max() <= 30


def test_hostemu_flags_bad_spectra(hostemu):
    g = load_golden('fit_sincgauss_sn3')
    cfg = config_from_golden(g, uncertainty=False)
    cube = g['cube'][:3].copy()
    cube[0, :] = np.nan
    cube[1, :] = -1e-18
    cube[2, 50] = np.nan                     # a NaN sample is dropped and the rest fitted (LuciBase.py:358-360)
    rows, sol, unc, st = run_hostemu(hostemu, cfg, cube, g['theta'][:3], g['v0'][:3], g['b0'][:3])
    assert st[0] & (1 << 20) and np.all(rows[0] == 0)
    assert st[1] & (1 << 22) and np.all(rows[1] == 0)
    assert st[2] & (1 << 16) and st[2] & (1 << 24) and np.isfinite(rows[2]).all()
    np.testing.assert_allclose(rows[2][15], g['ref_velocities'][2][0], atol=0.05)


def test_profiles_against_float64(hostemu):
    """FP32 unit profiles and analytic partials vs the float64 oracle profiles (LuciFunctions.py restated)."""
    from oracle import luci_oracle as O
    from tests.common import ptr
    w = 2.446
    x = np.linspace(0, 650, 1301).astype(np.float32)
    for model, name in ((0, 'gaussian'), (1, 'sinc'), (2, 'sincgauss')):
        for sig in (1.0, 1.3, 2.0, 3.5, 9.0):
            pp = np.float32(481.3217)
            g = np.zeros_like(x); gp = np.zeros_like(x); gs = np.zeros_like(x)
            hostemu.hostemu_profile(model, pp, np.float32(sig), np.float32(w), len(x), ptr(x), ptr(g), ptr(gp), ptr(gs))
            xd, p, s, wd = x.astype(np.float64), float(pp), float(np.float32(sig)), float(np.float32(w))
            f = {0: lambda p_, s_: O.gaussian_profile(xd, 1, p_, s_), 1: lambda p_, s_: O.sinc_profile(xd, 1, p_, wd),
                 2: lambda p_, s_: O.sincgauss_profile(xd, 1, p_, s_, wd)}[model]
            h = 1e-5
            assert np.abs(g - f(p, s)).max() < 5e-7, (name, sig)
            assert np.abs(gp - (f(p + h, s) - f(p - h, s)) / (2 * h)).max() < 2e-6, (name, sig)
            if model != 1:
                assert np.abs(gs - (f(p, s + h) - f(p, s - h)) / (2 * h)).max() < 4e-6, (name, sig)


def test_faddeeva_against_scipy(hostemu):
    from scipy.special import wofz
    from tests.common import ptr
    rng = np.random.default_rng(0)
    x = rng.uniform(-60, 60, 20000).astype(np.float32)
    y = np.exp(rng.uniform(np.log(0.02), np.log(12), 20000)).astype(np.float32)
    wr = np.zeros_like(x); wi = np.zeros_like(x)
    hostemu.hostemu_faddeeva(len(x), ptr(x), ptr(y), ptr(wr), ptr(wi))
    ref = wofz(x.astype(np.float64) + 1j * y.astype(np.float64))
    assert np.abs(wr - ref.real).max() < 6e-7 and np.abs(wi - ref.imag).max() < 6e-7


def _double_component_case(n=6, seed=3):
    """SN3 5 lines + a second Halpha component (vel_rel = sigma_rel = [1,1,1,1,1,2], BASELINE config 4): truth, cube,
    and per-group priors (group order of first appearance: main, second component)."""
    from oracle import luci_oracle as O
    rng = np.random.default_rng(seed)
    axis, n_steps, delta_x, _ = O.sim_axis('SN3', 5000)
    ax32 = axis.astype(np.float32)
    a64 = ax32.astype(np.float64)
    lines = ['Halpha', 'NII6548', 'NII6583', 'SII6716', 'SII6731', 'Halpha']
    theta = 11.96
    corr = 1 / np.cos(np.deg2rad(theta))
    w = corr * 1e7 / (2 * delta_x * n_steps)
    cube, truth = [], []
    for i in range(n):
        v1, v2 = rng.uniform(-80, 80), 0.0
        v2 = v1 + 200.0 + rng.uniform(-20, 20)
        b1, b2 = rng.uniform(30, 50), rng.uniform(30, 50)
        a83 = 0.3
        amps = [1.0, a83 / 3, a83, 0.15, 0.12, 0.5]
        spec = np.full(len(a64), 0.1)
        for k, line in enumerate(lines):
            v, b = (v2, b2) if k == 5 else (v1, b1)
            pos = 1e7 / (O.LINE_DICT[line] * (1 + v / 299792.0))
            spec += O.sincgauss_profile(a64, amps[k], pos, pos * b / 299792.0, w)
        spec += rng.normal(0, 1.0 / 60.0, len(a64))
        cube.append((spec * 1e-17).astype(np.float32))
        truth.append((v1, v2, b1, b2))
    truth = np.array(truth)
    v0 = (truth[:, :2] + rng.normal(0, 10, (n, 2))).astype(np.float32)
    b0 = (truth[:, 2:] * (1 + rng.normal(0, 0.1, (n, 2)))).astype(np.float32)
    return dict(lines=lines, axis32=ax32, delta_x=delta_x, n_steps=n_steps, theta=np.full(n, theta, np.float32),
                cube=np.stack(cube), truth=truth, v0=v0, b0=b0)


def test_hostemu_double_component_with_per_group_priors(hostemu):
    """Two velocity components started apart (one prior per tie group) separate and recover the truth; started from one
    shared prior the fit can stay on the symmetric saddle the reference sits on (SURVEY hard part 7)."""
    from luci_b200._capi import FitConfig
    c = _double_component_case()
    rel = [1, 1, 1, 1, 1, 2]
    mk = lambda pg: FitConfig('sincgauss', c['lines'], rel, rel, c['axis32'], filter='SN3', delta_x=c['delta_x'],
                              n_steps=c['n_steps'], zpd_index=0, uncertainty_bool=False, nii_cons=True, prior_groups=pg)
    rows, sol, unc, st = run_hostemu(hostemu, mk(True), c['cube'], c['theta'], np.ascontiguousarray(c['v0']),
                                     np.ascontiguousarray(c['b0']), want_unc=False)
    L = 6
    vel = rows[:, 3 * L:4 * L]
    assert ((st >> 16) & 1).all()
    np.testing.assert_allclose(vel[:, 0], c['truth'][:, 0], atol=3.0)          # main component
    np.testing.assert_allclose(vel[:, 5], c['truth'][:, 1], atol=6.0)          # second Halpha component
    np.testing.assert_allclose(vel[:, 1:5], np.repeat(vel[:, :1], 4, axis=1), atol=1e-3)   # tied to the main group
    amp2 = rows[:, 5] / rows[:, 0]
    assert np.all(np.abs(amp2 - 0.5) < 0.08)
    rows1, _, _, _ = run_hostemu(hostemu, mk(False), c['cube'], c['theta'], np.ascontiguousarray(c['v0'][:, 0]),
                                 np.ascontiguousarray(c['b0'][:, 0]), want_unc=False)
    ratio = rows[:, 7 * L] / rows1[:, 7 * L]                                    # never worse, and clearly better where
    assert (ratio < 1 + 1e-4).all() and ratio.min() < 0.7                       # the shared prior stays on the saddle


@pytest.mark.parametrize('name', ['parity_sincgauss_cfg2', 'parity_sincgauss_heldout'])
def test_hostemu_large_sample_parity_config2(hostemu, name):
    """1000 spaxels of the benchmark configuration (nz = 841, 5-line sincgauss, ties + [NII] ratio) against the
    unmodified reference and against the machine-precision optimum of its objective: the north_star's 99.9 % criterion.
    The second fixture (400 spaxels, another seed) is held out: no constant of the device code was tuned on it."""
    from tests.common import check_large_sample_parity, load_tight
    g = load_golden(name)
    cfg = config_from_golden(g, uncertainty=False)
    rows, sol, unc, st = run_hostemu(hostemu, cfg, g['cube'], g['theta'], g['v0'], g['b0'], want_unc=False)
    check_large_sample_parity(rows, st, g, load_tight(name))
    assert (((st >> 8) & 0xff).mean()) < 7


def test_hostemu_freeze_mode_matches_reference(hostemu):
    """initial_values = frozen (velocity, broadening): amplitudes and continuum only, no ties, no boxes
    (LuciFit.py:159-162, 688-709) against the reference's own freeze branch."""
    g = load_golden('fit_sincgauss_freeze')
    cfg = config_from_golden(g)
    rows, sol, unc, st = run_hostemu(hostemu, cfg, g['cube'], g['theta'], g['v0'], g['b0'], want_unc=False)
    rep = parity_report(rows, g, g['model'])
    assert rep['ok'].all(), (rep['dv'], rep['db'], rep['df'], rep['dchi'])
    assert rep['dv'].max() < 1e-2 and rep['db'].max() < 1e-5 and rep['df'].max() < 1e-3 and np.median(rep['df']) < 1e-5
    # frozen: velocity and broadening are the inputs, the [NII] amplitudes are no longer tied 1:3
    np.testing.assert_allclose(rep['q']['velocities'][:, 0], g['v0'], atol=1e-3)
    np.testing.assert_allclose(rep['q']['sigmas'][:, 0], np.abs(g['b0']), rtol=1e-5)
    assert np.abs(rep['q']['fluxes'][:, 1] * 3 / rep['q']['fluxes'][:, 2] - 1).max() > 1e-3
    assert (((st >> 8) & 0xff) <= 4).all()                  # a linear problem: one Gauss-Newton step and its check
    # the model and the uncertainty stage are refused exactly where the reference itself fails
    import ctypes as C
    from luci_b200 import _capi
    bad = config_from_golden(g, uncertainty=True)
    sz = C.c_size_t(0)
    assert _capi.lib().lucifit_workspace_bytes(bad.byref(), 10, C.byref(sz)) != 0
    assert b'freeze' in _capi.lib().lucifit_last_error()


@pytest.mark.parametrize('name', ['parity_sincgauss_cfg2', 'fit_gaussian_sn3'])
def test_analytic_gradient_matches_finite_differences(hostemu, name):
    """J^T r of the tie-eliminated model (analytic partials, [NII] ratio chain rule, folded wing path) equals the
    central finite difference of the evaluated sum of squares in every free parameter, off the optimum."""
    import ctypes as C
    from tests.common import ptr
    g = load_golden(name)
    cfg = config_from_golden(g, uncertainty=False)
    L, P, NH = 5, 8, 36
    for i in (0, 1):
        spec = np.ascontiguousarray(g['cube'][i], dtype=np.float32)
        scale = float(g['ref_scale'][i])
        th = np.zeros(P, np.float32)
        th[:L] = g['ref_amplitudes'][i] / scale
        th[5], th[6], th[7] = g['ref_velocities'][i][0] + 3.0, g['ref_sigmas'][i][0] * 1.05, g['ref_continuum'][i] / scale
        th[0] *= 0.97

        def ev(t):
            t = np.ascontiguousarray(t, dtype=np.float32)
            chi, hg = C.c_double(0), np.zeros(NH + P, np.float32)
            assert hostemu.hostemu_eval(cfg.byref(), ptr(spec), C.c_float(float(g['theta'][i])), ptr(t), C.byref(chi), ptr(hg)) == 0
            return chi.value, hg

        _, hg = ev(th)
        for j, h in zip(range(P), [1e-3] * 5 + [0.05, 0.05, 1e-3]):
            if j == 1:
                assert hg[NH + j] == 0.0                     # the tied [NII]6548 column is dead
                continue
            tp, tm = th.copy(), th.copy()
            tp[j] += h; tm[j] -= h
            fd = -(ev(tp)[0] - ev(tm)[0]) / (2.0 * (float(tp[j]) - float(tm[j])))      # -1/2 d(sum r^2)/d theta_j
            np.testing.assert_allclose(hg[NH + j], fd, rtol=3e-2, atol=2e-6 * abs(hg[NH:]).max(), err_msg='%s param %d' % (name, j))


def test_hostemu_region_strip_config1(hostemu):
    """BASELINE config 1 as a strip (100 x 8 spaxels, five-line sinc, 70 % mask; every y-row went through the reference's
    row worker): the 562 fitted spaxels against the optimum of the reference's objective.  The reference's SLSQP misses
    its own optimum on the multi-modal sinc objective in ~15 % of them (``ref_off``)."""
    g = load_golden('region_sinc_strip')
    cfg = config_from_golden(g, uncertainty=False)
    rows, sol, unc, st = run_hostemu(hostemu, cfg, g['cube'], g['theta'], g['v0'], g['b0'], want_unc=False)
    rep, opt, direct = assert_fixture_parity('region_sinc_strip', rows, st, g)
    assert opt['same'].mean() >= 0.99 and direct.mean() >= 0.8, (opt['same'].mean(), direct.mean())
