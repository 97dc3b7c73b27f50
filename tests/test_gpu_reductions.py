"""GPU tests of the HBM-bound kernels (deep image, SNR map, binning, synthetic cube) against the reference's
outputs in tests/golden/reductions_sn3.npz and against the oracle."""
import numpy as np
import pytest
import torch

from tests.common import load_golden

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.device('cuda', 0))


def test_deep_image_matches_reference_including_zero_tail():
    from luci_b200 import engine
    g = load_golden('reductions_sn3')
    cube = g['cube']
    nx, ny, nz = cube.shape
    deep = engine.deep_image(_t(cube).view(nx * ny, nz), n_zero_tail=(nx % 10) * ny).view(nx, ny).t().cpu().numpy()
    np.testing.assert_allclose(deep, g['ref_deep'], rtol=2e-6, atol=1e-24)    # FP32 summation-order tolerance
    assert np.all(deep[:, 20:] == 0)
    full = engine.deep_image(_t(cube).view(nx * ny, nz)).view(nx, ny).cpu().numpy()
    np.testing.assert_allclose(full, np.nansum(cube.astype(np.float64), axis=2), rtol=2e-6, atol=1e-24)


@pytest.mark.parametrize('method', [1, 2])
def test_snr_map_matches_reference(method):
    from luci_b200 import engine
    from luci_b200._capi import argmin_abs
    from luci_b200.constants import snr_windows
    g = load_golden('reductions_sn3')
    cube, ax = g['cube'], g['axis32']
    nx, ny, nz = cube.shape
    (f0, f1), (n0, n1) = snr_windows('SN3')
    snr, deep = engine.snr_map(_t(cube).view(nx * ny, nz), method, argmin_abs(ax, f0), argmin_abs(ax, f1),
                               argmin_abs(ax, n0), argmin_abs(ax, n1), with_deep=True)
    snr = snr.view(nx, ny).t().cpu().numpy()
    ref = g['ref_snr%d' % method]
    finite = np.isfinite(ref)
    assert finite.sum() >= ref.size - 1                       # the all-NaN spaxel
    np.testing.assert_allclose(snr[finite], ref[finite], rtol=2e-4 if method == 1 else 2e-3, atol=1e-6)
    np.testing.assert_allclose(deep.view(nx, ny).cpu().numpy(), np.nansum(cube.astype(np.float64), axis=2), rtol=2e-6, atol=1e-24)


def test_bin_cube_matches_reference():
    from luci_b200 import engine
    g = load_golden('reductions_sn3')
    b, x0, x1, y0, y1 = [int(v) for v in g['bin_args']]
    out = engine.bin_cube(_t(g['cube']), b, x0, x1, y0, y1).cpu().numpy()
    np.testing.assert_allclose(out, g['ref_binned'], rtol=1e-6, atol=2e-24)   # FP32 sums of ~4e-18 terms
    odd = engine.bin_cube(_t(g['cube']), 3, 0, 23, 0, 14).cpu().numpy()     # ragged: 23/3 -> 7, 14/3 -> 4
    assert odd.shape == (7, 4, g['cube'].shape[2])


def test_sim_cube_matches_lucisim_restatement():
    from luci_b200 import engine
    from luci_b200.sim import SN3_LINES, sim_axis
    from oracle import luci_oracle as O
    axis, n, dx = sim_axis('SN3', 5000)
    ax32 = axis.astype(np.float32)
    for model in ('gaussian', 'sinc', 'sincgauss'):
        cfg = engine.FitConfig(model, SN3_LINES, [1] * 5, [1] * 5, ax32, 'SN3', dx, n, 0, nii_cons=False)
        amps = np.array([[1.0, 0.1, 0.3, 0.15, 0.12], [0.7, 0.2, 0.6, 0.1, 0.3]], np.float32)
        vel = np.array([[-120.0] * 5, [85.0] * 5], np.float32)
        broad = np.array([[40.0] * 5, [28.0] * 5], np.float32)
        out = engine.sim_cube(cfg, _t(amps), _t(vel), _t(broad), _t(np.full(2, 11.96, np.float32))).cpu().numpy()
        for i in range(2):
            # the device evaluates on the float32-rounded axis (what the cube's header axis is): same for the oracle
            spec = np.zeros(len(ax32))
            corr = 1 / np.cos(np.deg2rad(11.96))
            a64 = ax32.astype(np.float64)
            for k, line in enumerate(SN3_LINES):
                lam = 1e7 / O.LINE_DICT[line]
                pos = a64[np.argmin(np.abs(a64 - ((vel[i, k] / 3e5) * lam + lam)))]
                sig = broad[i, k] * pos / (3e5 * corr)
                w = corr * 1e7 / (2 * dx * n)
                spec += {'gaussian': lambda: O.gaussian_profile(a64, amps[i, k], pos, sig),
                         'sinc': lambda: O.sinc_profile(a64, amps[i, k], pos, w),
                         'sincgauss': lambda: O.sincgauss_profile(a64, amps[i, k], pos, sig, w)}[model]()
            np.testing.assert_allclose(out[i], spec, atol=2e-6)


def test_large_reduction_is_size_independent():
    """Checksum-of-checksums: deep image of a big random cube equals torch's own reduction; SNR is order invariant."""
    from luci_b200 import engine
    dev = torch.device('cuda', 0)
    g = torch.Generator(device=dev); g.manual_seed(1)
    cube = torch.rand((64 * 48, 841), generator=g, device=dev, dtype=torch.float32)
    deep = engine.deep_image(cube)
    torch.testing.assert_close(deep, cube.sum(dim=1), rtol=2e-6, atol=0)
    snr1 = engine.snr_map(cube, 1, 300, 360, 10, 60)
    perm = torch.randperm(cube.shape[0], device=dev)
    # row alignment (hence the float4 peel and the summation order) changes with the position in the cube
    torch.testing.assert_close(engine.snr_map(cube[perm].contiguous(), 1, 300, 360, 10, 60), snr1[perm], rtol=3e-6, atol=0)
    mx, mean = cube.max(dim=1).values, cube.mean(dim=1)
    std = cube[:, 10:60].double().std(dim=1, unbiased=False).float()
    ref = (mx - mean) / std.sqrt() / cube.abs().mean(dim=1).sqrt()
    torch.testing.assert_close(snr1, ref, rtol=2e-5, atol=0)


def test_row_median_equals_numpy_nanmedian():
    """lucifit_row_median (the level of the absorption template, LuciBase.py:354-356) against np.nanmedian, odd and even
    counts, NaN samples, all-NaN rows, both register-sort sizes."""
    from luci_b200 import engine
    dev = torch.device('cuda', 0)
    rng = np.random.default_rng(4)
    for nz, lo, hi in ((841, 210, 631), (841, 0, 841), (300, 7, 8), (2000, 500, 1500)):
        cube = rng.normal(size=(257, nz)).astype(np.float32)
        cube[rng.integers(0, 257, 300), rng.integers(0, nz, 300)] = np.nan
        cube[5, :] = np.nan
        got = engine.row_median(torch.from_numpy(cube).to(dev), lo, hi).cpu().numpy()
        with np.errstate(all='ignore'):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                want = np.nanmedian(cube[:, lo:hi].astype(np.float64), axis=1).astype(np.float32)
        np.testing.assert_array_equal(got, want)
