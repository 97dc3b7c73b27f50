"""Sky-line calibration (Luci.skyline_calibration, LuciBase.py:1205-1284; Fit.fit(sky_line=True), LuciFit.py:837-928):
a single sinc at the OH 6498.729 A line started at 80 km/s.  Golden vectors: tests/golden/skyline_sn3.npz, produced by
the unmodified reference (oracle/make_golden_skyline.py)."""
import os

import numpy as np
import pytest

from luci_b200._capi import FitConfig
from tests.common import load_golden, run_hostemu, split_rows


def skyline_config(g):
    return FitConfig('sinc', ['OH_0'], [1], [1], g['axis32'], filter='SN3', delta_x=float(g['delta_x']),
                     n_steps=int(g['n_steps']), zpd_index=0, uncertainty_bool=True, nii_cons=False, ml_scale=False,
                     extra_lines={'OH_0': float(g['oh_nm'])})


def check_against_reference(rows, g):
    q = split_rows(rows, 1)
    dv = np.abs(q['velocities'][:, 0] - g['ref_velocity'])
    assert dv.max() < 0.05, dv                                  # km/s (bar of the fit path: 0.5)
    # the sky-line branch inverts the Hessian, the regular fit half of it (LuciFit.py:873-876 vs :718): sqrt(2) apart
    rel = np.abs(q['vels_errors'][:, 0] / np.sqrt(2.0) / g['ref_velocity_error'] - 1)
    assert np.median(rel) < 5e-3 and rel.max() < 5e-2, rel
    assert np.abs(q['velocities'][:, 0] - g['truth']).max() < 3.0


def test_hostemu_skyline_fit_matches_reference(hostemu):
    g = load_golden('skyline_sn3')
    cfg = skyline_config(g)
    n = g['cube'].shape[0]
    # like Luci.skyline_calibration: the reference's 80 km/s guess is moved to the matched-filter maximum within +-200 km/s
    from tests.test_prior_estimator import hostemu_prior
    v0, _ = hostemu_prior(hostemu, cfg, g['cube'], g['theta'], v_range=(-120.0, 280.0))
    assert np.abs(v0 - g['ref_velocity']).max() < 24.0          # half a channel: inside the basin of the optimum
    rows, sol, unc, st = run_hostemu(hostemu, cfg, g['cube'], g['theta'], v0, np.zeros(n, np.float32))
    assert ((st >> 16) & 1).all()
    check_against_reference(rows, g)
    # fit_sol layout (amplitude, position, sigma), continuum: position against the reference's
    np.testing.assert_allclose(sol[:, 1], g['ref_fit_sol'][:, 1], rtol=2e-7)


def test_unknown_line_names_still_raise():
    g = load_golden('skyline_sn3')
    with pytest.raises(Exception, match='line name'):
        FitConfig('sinc', ['OH_0'], [1], [1], g['axis32'])


@pytest.mark.gpu
def test_gpu_skyline_fit_matches_reference():
    import torch
    from luci_b200 import engine
    g = load_golden('skyline_sn3')
    cfg = skyline_config(g)
    dev = torch.device('cuda', 0)
    n = g['cube'].shape[0]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    v0, _ = engine.estimate_prior(cfg, t(g['cube']), t(g['theta']), v_range=(-120.0, 280.0))
    res = engine.fit_batch(cfg, t(g['cube']), t(g['theta']), v0, t(np.zeros(n)))
    assert ((res['status'].cpu().numpy() >> 16) & 1).all()
    check_against_reference(res['rows'].cpu().numpy(), g)


@pytest.mark.gpu
def test_gpu_skyline_calibration_grid(tmp_path):
    """The API end to end: a 460 x 440 cube whose OH line moves smoothly across the field; every region of the
    n_grid x n_grid set recovers its velocity, and the reference's diagonal region sum is reproduced."""
    import torch
    from luci_b200.fitsio import read_fits
    from luci_b200.luci import Luci
    g = load_golden('skyline_sn3')
    dev = torch.device('cuda', 0)
    axis = g['axis32'].astype(np.float64)
    nz = len(axis)
    nx, ny, n_grid, bin_size = 460, 440, 3, 4
    sinc_w = (1.0 / np.cos(np.deg2rad(11.96))) * 1e7 / (2 * float(g['delta_x']) * int(g['n_steps']))
    xs, ys = np.meshgrid(np.arange(nx), np.arange(ny), indexing='ij')
    vtrue = 60.0 + 0.08 * xs + 0.05 * ys                                    # km/s
    p = 1e7 / (float(g['oh_nm']) * (1 + vtrue / 299792.0))
    rng = np.random.default_rng(7)
    cube = torch.empty((nx, ny, nz), dtype=torch.float32, device=dev)
    ax_t = torch.from_numpy(axis).to(dev)
    for x0 in range(0, nx, 64):                                             # built on the device in slabs
        d = (ax_t[None, None, :] - torch.from_numpy(p[x0:x0 + 64]).to(dev)[:, :, None]) / sinc_w
        prof = torch.sinc(d)
        noise = torch.from_numpy(rng.normal(0, 0.02, prof.shape)).to(dev)
        cube[x0:x0 + 64] = ((prof + 0.08 + noise) * 3e-17).to(torch.float32)
    os.makedirs(tmp_path / 'L' / 'Data')
    with open(tmp_path / 'L' / 'Data' / 'sky_lines.dat', 'w') as fh:
        fh.write('#### SKYLINE DATA\n#### Hanuschik, A&A, 407, 3 (2003)\nWavelength,Sigma,Strength\n6498.729, 1, 1\n')
    hdr = {'FILTER': 'SN3', 'STEP': float(g['delta_x']), 'STEPNB': int(g['n_steps']), 'ZPDINDEX': 0}
    luci = Luci.from_arrays(cube, hdr, str(tmp_path), 'SKY', spectrum_axis=g['axis32'], device=dev)
    vel, fit_vector, sky, vel_grid, vel_unc, ax = luci.skyline_calibration(str(tmp_path / 'L'), n_grid, bin_size=bin_size)
    assert vel_grid.shape == (n_grid, n_grid) and vel_unc.shape == (n_grid, n_grid) and fit_vector.shape == (nz,)
    x_step, y_step = int((nx - 400) / n_grid), int((ny - 400) / n_grid)
    for xg in range(n_grid):
        for yg in range(n_grid):
            xc, yc = 200 + int(0.5 * x_step * (xg + 1)), 200 + int(0.5 * y_step * (yg + 1))
            want = np.mean([vtrue[xc + i, yc + i] for i in range(bin_size)])        # the reference's diagonal region
            assert abs(vel_grid[xg, yg] - want) < 1.5, (xg, yg, vel_grid[xg, yg], want)
            assert 0 < vel_unc[xg, yg] < 2.0
    host = cube.cpu().numpy().astype(np.float64)
    xc, yc = 200 + int(0.5 * x_step * n_grid), 200 + int(0.5 * y_step * n_grid)
    np.testing.assert_allclose(sky, bin_size * sum(host[xc + i, yc + i] for i in range(bin_size)), rtol=1e-6)
    assert vel == vel_grid[-1, -1]
    d, _ = read_fits(os.path.join(str(tmp_path), 'Luci_outputs', 'velocity_correction.fits'))
    np.testing.assert_allclose(d, vel_grid)
    luci.strict_reference = False                                                    # full bin_size x bin_size boxes
    _, _, sky2, vel_grid2, _, _ = luci.skyline_calibration(str(tmp_path / 'L'), n_grid, bin_size=bin_size)
    np.testing.assert_allclose(sky2, host[xc:xc + bin_size, yc:yc + bin_size].sum(axis=(0, 1)), rtol=1e-6)
    assert np.abs(vel_grid2 - vel_grid).max() < 2.0
    # two sky lines in the file (the second one has no signal in this cube): per-group priors, one broadening group --
    # the first line's grid is unchanged
    with open(tmp_path / 'L' / 'Data' / 'sky_lines.dat', 'w') as fh:
        fh.write('#### SKYLINE DATA\nWavelength,Sigma,Strength\n6498.729, 1, 1\n6553.617, 1, 1\n')
    luci.strict_reference = True
    _, _, _, vel_grid3, vel_unc3, _ = luci.skyline_calibration(str(tmp_path / 'L'), n_grid, bin_size=bin_size)
    assert np.abs(vel_grid3 - vel_grid).max() < 0.5 and (vel_unc3 > 0).all()
