"""lucifit_estimate_prior: the device stand-in for the reference's CNN prior (LuciFit.py:335-362).

The reference's network cannot run here (TensorFlow absent, R5000-SN3 weights not in the checkout), so the estimator
is pinned on what a prior is for: it must put every fit inside the basin of the optimum the unmodified reference
reached (tests/golden: reference velocities / broadenings), and the fits started from it must land on that optimum.
CPU tier: the same source compiled with LANES = 1 (tests/hostemu).  GPU tier: through the C ABI and the Luci API."""
import ctypes as C

import numpy as np
import pytest

from tests.common import (check_large_sample_parity, config_from_golden, load_golden, parity_report, ptr, run_hostemu)

HALF_CHANNEL_KMS = 24.0          # R = 5000 SN3 channel = 48 km/s: the basin of the optimum is about half of it


def hostemu_prior(lib, cfg, cube, theta, v_range=(-500.0, 500.0), v_step=12.0):
    n = cube.shape[0]
    cube = np.ascontiguousarray(cube, dtype=np.float32)
    theta = np.ascontiguousarray(theta, dtype=np.float32)
    vel, broad = np.zeros(n, np.float32), np.zeros(n, np.float32)
    assert lib.hostemu_estimate_prior(cfg.byref(), n, ptr(cube), ptr(theta), None, v_range[0], v_range[1], v_step,
                                      ptr(vel), ptr(broad)) == 0
    return vel, broad


def check_prior_against_reference(vel, broad, g, model):
    dv = np.abs(vel - g['ref_velocities'][:, 0])
    assert dv.max() < HALF_CHANNEL_KMS, dv.max()
    assert np.median(dv) < 10.0
    if model != 'sinc':                      # the sinc model has no broadening parameter
        ratio = broad / g['ref_sigmas'][:, 0]
        assert 0.3 < ratio.min() and ratio.max() < 3.0, (ratio.min(), ratio.max())


@pytest.mark.parametrize('name', ['parity_sincgauss_cfg2', 'fit_sincgauss_sn3', 'fit_gaussian_sn3', 'fit_sinc_sn3',
                                  'fit_gaussian_sn2', 'fit_gaussian_sn1', 'fit_sincgauss_single', 'fit_sincgauss_trans'])
def test_hostemu_prior_lands_in_the_basin_of_the_reference_optimum(hostemu, name):
    g = load_golden(name)
    cfg = config_from_golden(g, uncertainty=False)
    vel, broad = hostemu_prior(hostemu, cfg, g['cube'], g['theta'])
    check_prior_against_reference(vel, broad, g, g['model'])


def test_hostemu_fits_from_estimated_prior_meet_the_parity_bar(hostemu):
    """1000 spaxels of the benchmark configuration: fits started from the estimated prior reach the reference's optimum
    (north_star tolerances) and need fewer evaluations than from the CNN-like prior of the fixture (truth + N(0, 15 km/s))."""
    from tests.common import load_tight
    g = load_golden('parity_sincgauss_cfg2')
    cfg = config_from_golden(g, uncertainty=False)
    vel, broad = hostemu_prior(hostemu, cfg, g['cube'], g['theta'])
    rows, sol, unc, st = run_hostemu(hostemu, cfg, g['cube'], g['theta'], vel, broad, want_unc=False)
    check_large_sample_parity(rows, st, g, load_tight('parity_sincgauss_cfg2'))
    assert ((st >> 8) & 0xff).mean() < 5.5


def test_hostemu_prior_flags_unusable_spectra(hostemu):
    g = load_golden('fit_sincgauss_sn3')
    cfg = config_from_golden(g, uncertainty=False)
    cube = g['cube'][:3].copy()
    cube[0, cfg.c.fit_lo:cfg.c.fit_hi] = np.nan
    cube[1] = 0.0
    cube[2, cfg.c.fit_lo + 10] = np.nan                   # one NaN sample: left out, like the fit does (LuciBase.py:358-360)
    vel, broad = hostemu_prior(hostemu, cfg, cube, g['theta'][:3])
    assert vel[0] == 0 and broad[0] == 0                 # nothing left of the window
    assert np.isfinite(vel).all() and np.isfinite(broad).all()
    assert abs(vel[2] - g['ref_velocities'][2, 0]) < HALF_CHANNEL_KMS


def test_capi_rejects_bad_velocity_grids():
    from luci_b200 import _capi
    g = load_golden('fit_sincgauss_sn3')
    cfg = config_from_golden(g, uncertainty=False)
    lib = _capi.lib()
    one = C.c_void_p(16)                                  # never dereferenced: argument checks come first
    assert lib.lucifit_estimate_prior(cfg.byref(), 4, one, None, one, None, -500.0, 500.0, 0.0, one, one, None) != 0
    assert b'v_step' in lib.lucifit_last_error()
    assert lib.lucifit_estimate_prior(cfg.byref(), 4, one, None, one, None, -5000.0, 5000.0, 1.0, one, one, None) != 0
    assert b'512' in lib.lucifit_last_error()
    assert lib.lucifit_estimate_prior(cfg.byref(), 0, None, None, None, None, -500.0, 500.0, 6.0, None, None, None) == 0


# ----------------------------------------------------------------------------------------------------- GPU tier
@pytest.mark.gpu
@pytest.mark.parametrize('name', ['parity_sincgauss_cfg2', 'fit_gaussian_sn2', 'fit_sinc_sn3'])
def test_gpu_prior_lands_in_the_basin_of_the_reference_optimum(name):
    import torch
    from luci_b200 import engine
    g = load_golden(name)
    cfg = config_from_golden(g, uncertainty=False)
    dev = torch.device('cuda', 0)
    cube = torch.from_numpy(np.ascontiguousarray(g['cube'], dtype=np.float32)).to(dev)
    theta = torch.from_numpy(np.ascontiguousarray(g['theta'], dtype=np.float32)).to(dev)
    vel, broad = engine.estimate_prior(cfg, cube, theta)
    check_prior_against_reference(vel.cpu().numpy(), broad.cpu().numpy(), g, g['model'])
    # compacted index list and background: same answers for the selected rows
    idx = torch.arange(cube.shape[0] - 1, -1, -2, dtype=torch.int64, device=dev)
    v2, b2 = engine.estimate_prior(cfg, cube, theta[idx].contiguous(), index=idx)
    if name != 'fit_sincgauss_trans':
        np.testing.assert_array_equal(v2.cpu().numpy(), vel[idx].cpu().numpy())
    bkg = torch.full((cube.shape[1],), float(cube.median()) * 0.5, dtype=torch.float32, device=dev)
    v3, b3 = engine.estimate_prior(cfg, cube, theta, bkg=bkg)
    np.testing.assert_allclose(v3.cpu().numpy(), vel.cpu().numpy(), atol=0.5)       # a constant offset leaves the median-subtracted window alone


@pytest.mark.gpu
def test_gpu_fits_from_estimated_prior_meet_the_parity_bar():
    import torch
    from luci_b200 import engine
    g = load_golden('parity_sincgauss_cfg2')
    cfg = config_from_golden(g, uncertainty=False)
    dev = torch.device('cuda', 0)
    cube = torch.from_numpy(np.ascontiguousarray(g['cube'], dtype=np.float32)).to(dev)
    theta = torch.from_numpy(np.ascontiguousarray(g['theta'], dtype=np.float32)).to(dev)
    vel, broad = engine.estimate_prior(cfg, cube, theta)
    res = engine.fit_batch(cfg, cube, theta, vel, broad)
    rows, st = res['rows'].cpu().numpy(), res['status'].cpu().numpy()
    rep = parity_report(rows, g, g['model'])
    assert ((st >> 16) & 1).all()
    assert (rep['dv'] <= 0.5).all() and (rep['db'] <= 1e-3).all() and (rep['dchi'] <= 1e-4).all()
    assert (rep['df'] <= 1e-3).mean() >= 0.995
    assert ((st >> 8) & 0xff).mean() < 5.5


@pytest.mark.gpu
def test_gpu_luci_api_uses_the_estimator_when_no_prior_is_given(tmp_path):
    """ML_bool=True without prior_values: same maps as with the CNN-like priors of the simulation (same optimum)."""
    import torch
    from luci_b200.luci import Luci
    from luci_b200.sim import SN3_LINES, SyntheticCube
    dev = torch.device('cuda', 0)
    sc = SyntheticCube(16, 12, model='sincgauss', resolution=5000, seed=11, snr_range=(15.0, 100.0))
    cube = sc.synthesize(dev)
    luci = Luci.from_arrays(cube, sc.hdr_dict, str(tmp_path), 'AUTO', spectrum_axis=sc.axis32, device=dev)
    v0, b0 = sc.priors()
    vel_p, broad_p, flux_p, _ = luci.fit_cube(SN3_LINES, 'sincgauss', [1] * 5, [1] * 5, 0, 16, 0, 12, prior_values=(v0, b0))
    vel_a, broad_a, flux_a, _ = luci.fit_cube(SN3_LINES, 'sincgauss', [1] * 5, [1] * 5, 0, 16, 0, 12)
    pv, pb = (t.cpu().numpy() for t in luci.last_prior)
    assert np.abs(pv.reshape(12, 16) - vel_p[:, :, 0]).max() < HALF_CHANNEL_KMS
    same = (np.abs(vel_a - vel_p).max(axis=2) < 0.05) & (np.abs(flux_a / flux_p - 1).max(axis=2) < 2e-3)
    assert same.mean() >= 0.99, same.mean()
    assert luci.last_fit_summary['n_converged'] == 16 * 12
    # the integrated-spectrum entry point takes the same route
    mask = np.zeros((16, 12), bool)
    mask[4:9, 3:8] = True
    _, _, fd = luci.fit_spectrum_region(SN3_LINES, 'sincgauss', [1] * 5, [1] * 5, mask)
    assert abs(fd['vel_ml'] - fd['velocities'][0]) < HALF_CHANNEL_KMS and fd['chi2'] > 0
    # ML_bool=False keeps the reference's behaviour: start at zero (and say so for models that need a prior)
    luci.ML_bool = False
    with pytest.warns(UserWarning):
        luci.fit_cube(SN3_LINES, 'sincgauss', [1] * 5, [1] * 5, 0, 2, 0, 2)
