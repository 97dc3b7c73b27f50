"""The drop-in Python API (luci_b200.luci.Luci) on the GPU: same calls, returns and files as LuciBase.Luci."""
import os

import numpy as np
import pytest
import torch

from tests.common import load_golden

pytestmark = pytest.mark.gpu
SN3 = ['Halpha', 'NII6548', 'NII6583', 'SII6716', 'SII6731']


@pytest.fixture(scope='module')
def world(tmp_path_factory):
    from luci_b200.luci import Luci
    from luci_b200.sim import SyntheticCube
    dev = torch.device('cuda', 0)
    sc = SyntheticCube(12, 10, model='sincgauss', resolution=5000, seed=3, snr_range=(20.0, 100.0))
    cube = sc.synthesize(dev)
    out = str(tmp_path_factory.mktemp('luci'))
    hdr = {'CRPIX1': 6.0, 'CRPIX2': 5.0, 'CRVAL1': 10.0, 'CRVAL2': 20.0, 'CTYPE1': 'RA---TAN', 'CTYPE2': 'DEC--TAN',
           'PC1_1': 1.0, 'PC1_2': 0.0, 'PC2_1': 0.0, 'PC2_2': 1.0}
    luci = Luci.from_arrays(cube, sc.hdr_dict, out, 'OBJ', spectrum_axis=sc.axis32, header=hdr, device=dev)
    return luci, sc, out


def test_fit_cube_returns_and_files(world):
    from luci_b200.fitsio import read_fits
    luci, sc, out = world
    v0, b0 = sc.priors()
    vel, broad, flux, ampl = luci.fit_cube(SN3, 'sincgauss', [1] * 5, [1] * 5, 2, 9, 1, 8,
                                           prior_values=(v0[1:8, 2:9], b0[1:8, 2:9]), uncertainty_bool=True)
    assert vel.shape == broad.shape == flux.shape == ampl.shape == (7, 7, 5) and vel.dtype == np.float32
    truth = sc.vel_fit_truth[2:9, 1:8].T
    assert np.median(np.abs(vel[:, :, 0] - truth)) < 25.0     # LuciSim snaps lines to channels (+-24 km/s)
    base = os.path.join(out, 'Luci_outputs')
    d, h = read_fits(os.path.join(base, 'Velocity', 'OBJ_sincgauss_Halpha_velocity.fits'))
    np.testing.assert_array_equal(d, vel[:, :, 0])
    assert h['NAXIS1'] == 7 and h['NAXIS2'] == 7 and h['CRPIX1'] == 4.0 and h['CRPIX2'] == 4.0
    for rel in ('Amplitudes/OBJ_sincgauss_SII6731_Amplitude.fits', 'Fluxes/OBJ_sincgauss_NII6548_Flux_err.fits',
                'Broadening/OBJ_sincgauss_Halpha_broadening_err.fits', 'OBJ_sincgauss_Chi2.fits',
                'OBJ_sincgauss_continuum.fits', 'OBJ_sincgauss_continuum_error.fits', 'OBJ_deep.fits'):
        assert os.path.exists(os.path.join(base, rel)), rel
    m = luci.last_fit_maps
    assert (m['vels_errors'] > 0).all() and (m['flux_errors'] > 0).all() and (m['chi2'] > 0).all()
    np.testing.assert_allclose(m['corr'], 1 / np.cos(np.deg2rad(11.96)), rtol=1e-6)
    # tied amplitudes: [NII]6548 flux is a third of 6583 (area rule, LuciFit.py:584-595)
    np.testing.assert_allclose(flux[:, :, 1] * 3, flux[:, :, 2], rtol=1e-5)


def test_fit_region_masks_and_matches_fit_cube(world):
    luci, sc, out = world
    v0, b0 = sc.priors()
    mask = np.zeros((12, 10), bool)
    mask[3:7, 2:6] = True
    vel, broad, flux, chi2, m = luci.fit_region(SN3, 'sincgauss', [1] * 5, [1] * 5, mask, prior_values=(v0, b0))
    assert vel.shape == (10, 12, 5) and chi2.shape == (10, 12) and m is mask
    assert (vel[~mask.T] == 0).all() and (chi2[mask.T] > 0).all()
    velc, _, fluxc, _ = luci.fit_cube(SN3, 'sincgauss', [1] * 5, [1] * 5, 3, 7, 2, 6, prior_values=(v0[2:6, 3:7], b0[2:6, 3:7]))
    np.testing.assert_array_equal(vel[2:6, 3:7], velc)
    np.testing.assert_array_equal(flux[2:6, 3:7], fluxc)
    assert luci.last_fit_summary['n'] == 16
    assert os.path.exists(os.path.join(out, 'Luci_outputs', 'Velocity', 'OBJ_Halpha_velocity.fits'))   # no model suffix


def test_binned_fit_and_entire_cube(world):
    luci, sc, out = world
    v0, b0 = sc.priors()
    vb = 0.25 * (v0[0::2, 0::2] + v0[1::2, 0::2] + v0[0::2, 1::2] + v0[1::2, 1::2])
    bb = 0.25 * (b0[0::2, 0::2] + b0[1::2, 0::2] + b0[0::2, 1::2] + b0[1::2, 1::2])
    r = luci.fit_entire_cube(SN3, 'gaussian', [1] * 5, [1] * 5, binning=2, prior_values=(vb, bb))
    vel, broad, flux, ampl = r
    assert vel.shape == (5, 6, 5)
    assert luci.cube_binned.shape == (6, 5, luci.cube_final.shape[2])
    host = luci.cube_final.cpu().numpy()
    np.testing.assert_allclose(luci.cube_binned[1, 2].cpu().numpy(), host[2:4, 4:6].sum(axis=(0, 1)), rtol=1e-6)
    # binned amplitudes ~ 4x the unbinned ones (nansum, no division: LuciUtility.py:354)
    assert 3.0 < np.median(ampl[:, :, 0]) / 1e-17 < 5.0
    assert os.path.exists(os.path.join(out, 'Luci_outputs', 'Velocity', 'OBJ_2_gaussian_Halpha_velocity.fits'))


def test_deep_image_and_snr_map_files(world):
    from luci_b200.fitsio import read_fits
    luci, sc, out = world
    luci.create_deep_image()
    host = luci.cube_final.cpu().numpy().astype(np.float64)
    ref = host.sum(axis=2)
    ref[10:] = 0                                   # 12 mod 10 trailing x-rows stay zero in the reference
    np.testing.assert_allclose(luci.deep_image, ref.T, rtol=2e-6)
    d, h = read_fits(os.path.join(out, 'Luci_outputs', 'OBJ_deep.fits'))
    assert d.shape == (10, 12) and h['BITPIX'] == -64
    masks = luci.create_snr_map(0, 12, 0, 10, method=1)
    assert len(masks) == 4 and masks[0].shape == (10, 12)
    snr, _ = read_fits(os.path.join(out, 'Luci_outputs', 'SNR', 'OBJ_SNR.fits'))
    assert snr.shape == (10, 12) and (snr > 0).all()
    for v in (1, 3, 5, 10):
        mk = np.load(os.path.join(out, 'Luci_outputs', 'SNR', 'OBJ_SNR_%d_mask.npy' % v))
        np.testing.assert_array_equal(mk, snr >= v)
    luci.create_snr_map(2, 8, 1, 9, method=2)
    assert luci.snr_map.shape == (8, 6)


def test_fit_spectrum_region_dictionary(world):
    luci, sc, out = world
    mask = np.zeros((12, 10), bool)
    mask[4:8, 3:7] = True
    vtrue = float(sc.vel_fit_truth[4:8, 3:7].mean())
    axis, sky, d = luci.fit_spectrum_region(SN3, 'sincgauss', [1] * 5, [1] * 5, mask, prior_values=(vtrue + 8.0, 40.0),
                                            uncertainty_bool=True)
    assert axis.shape == sky.shape == (luci.cube_final.shape[2],)
    for key in ('fit_sol', 'fit_uncertainties', 'amplitudes', 'fluxes', 'flux_errors', 'chi2', 'velocities', 'sigmas',
                'vels_errors', 'sigmas_errors', 'axis_step', 'corr', 'continuum', 'continuum_error', 'fit_vector',
                'fit_axis'):
        assert key in d
    assert abs(d['velocities'][0] - vtrue) < 30.0
    assert d['fit_vector'].shape == sky.shape
    # the fit vector reproduces the integrated spectrum near H-alpha to within the noise
    i = int(np.argmax(sky))
    assert abs(d['fit_vector'][i] - sky[i]) < 0.2 * sky[i]


def test_unsupported_branches_raise(world):
    luci, sc, out = world
    with pytest.raises(NotImplementedError):
        luci.fit_cube(SN3, 'sinc', [1] * 5, [1] * 5, 0, 2, 0, 2, bayes_bool=True)
    with pytest.raises(Exception, match='freeze'):       # the reference's gaussian / sinc freeze paths raise IndexError
        luci.fit_cube(SN3, 'sinc', [1] * 5, [1] * 5, 0, 2, 0, 2, initial_values=[np.zeros((2, 2)), np.ones((2, 2))])
    with pytest.raises(NotImplementedError):
        luci.fit_cube(SN3, 'sincgauss', [1] * 5, [1] * 5, 0, 2, 0, 2, uncertainty_bool=True,
                      initial_values=[np.zeros((2, 2)), np.ones((2, 2))])
    with pytest.raises(Exception, match='vel_rel'):
        luci.fit_cube(SN3, 'sinc', [1] * 4, [1] * 5, 0, 2, 0, 2)


def test_double_component_fit_with_per_group_priors_and_uncertainties(tmp_path):
    """BASELINE config 4: 5 SN3 lines + a second Halpha component (vel_rel = sigma_rel = [1,1,1,1,1,2]) with
    uncertainties, through Luci.fit_cube with one prior per tie group; device result == CPU build of the same source."""
    from luci_b200.luci import Luci
    from luci_b200.sim import sim_axis
    from tests.test_hostemu_parity import _double_component_case
    c = _double_component_case(n=6, seed=3)
    dev = torch.device('cuda', 0)
    nz = c['cube'].shape[1]
    cube = torch.from_numpy(c['cube'].reshape(3, 2, nz)).to(dev)                  # [nx=3, ny=2, nz]
    hdr_dict = {'FILTER': 'SN3', 'STEP': c['delta_x'], 'STEPNB': c['n_steps'], 'ZPDINDEX': 0}
    luci = Luci.from_arrays(cube, hdr_dict, str(tmp_path), 'DBL', spectrum_axis=c['axis32'], device=dev)
    luci.interferometer_theta = np.full((3, 2), 11.96)
    rel = [1, 1, 1, 1, 1, 2]
    # maps are (ny, nx[, group]); spaxel s of the flat x-major list sits at x = s // 2, y = s % 2
    v0 = c['v0'].reshape(3, 2, 2).transpose(1, 0, 2).copy()
    b0 = c['b0'].reshape(3, 2, 2).transpose(1, 0, 2).copy()
    vel, broad, flux, ampl = luci.fit_cube(c['lines'], 'sincgauss', rel, rel, 0, 3, 0, 2, prior_values=(v0, b0),
                                           uncertainty_bool=True)
    assert vel.shape == (2, 3, 6)
    truth = c['truth'].reshape(3, 2, 4).transpose(1, 0, 2)
    np.testing.assert_allclose(vel[:, :, 0], truth[:, :, 0], atol=3.0)
    np.testing.assert_allclose(vel[:, :, 5], truth[:, :, 1], atol=6.0)
    m = luci.last_fit_maps
    assert (m['vels_errors'] > 0).all() and (m['vels_errors'] < 100).all() and (m['flux_errors'] > 0).all()
    assert (m['vels_errors'][:, :, 5] > m['vels_errors'][:, :, 0]).all()          # the weaker component is less certain
    with pytest.raises(Exception):
        luci.fit_cube(c['lines'], 'sincgauss', rel, rel, 0, 3, 0, 2, prior_values=(v0[:, :, :1], b0))
    # n_stoch restarts (LuciFit.py:662-687): never worse than the single start, same seed -> same maps
    chi1 = m['chi2'].copy()
    np.random.seed(7)
    luci.fit_cube(c['lines'], 'sincgauss', rel, rel, 0, 3, 0, 2, prior_values=(v0, b0), n_stoch=4)
    chi4 = luci.last_fit_maps['chi2'].copy()
    assert (chi4 <= chi1 * (1 + 1e-6)).all()
    np.random.seed(7)
    luci.fit_cube(c['lines'], 'sincgauss', rel, rel, 0, 3, 0, 2, prior_values=(v0, b0), n_stoch=4)
    np.testing.assert_array_equal(luci.last_fit_maps['chi2'], chi4)


def test_pca_background_equals_fit_of_presubtracted_cube(world, tmp_path):
    """bkgType='pca' (LuciBase.py:331-352): sky -= max(sky[band]) * (mean + coefficients . vectors) per spaxel, done
    on the device; same maps as fitting a cube from which that background was subtracted on the host in float64."""
    from luci_b200.luci import Luci
    luci, sc, out = world
    rng = np.random.default_rng(5)
    cube = luci.cube_final.cpu().numpy().astype(np.float64)                      # [nx, ny, nz]
    nx, ny, nz = cube.shape
    n_comp = 4
    z = np.linspace(0, 1, nz)
    vectors = np.stack([np.sin((k + 1) * np.pi * z) * 0.02 for k in range(n_comp)])
    mean = 0.01 + 0.005 * z
    coef = rng.normal(0, 1, (nx, ny, n_comp))
    nm = 1e7 / np.asarray(luci.spectrum_axis, dtype=np.float64)
    lo, hi = int(np.argmin(np.abs(nm - 675))), int(np.argmin(np.abs(nm - 670)))
    scale = np.nanmax(cube[:, :, lo:hi], axis=2)
    pre = cube - scale[:, :, None] * (mean[None, None, :] + coef @ vectors)
    v0, b0 = sc.priors()
    args = (SN3, 'sincgauss', [1] * 5, [1] * 5, 1, 11, 2, 9)
    pv = (v0[2:9, 1:11], b0[2:9, 1:11])
    vel, broad, flux, ampl = luci.fit_cube(*args, prior_values=pv, bkgType='pca', pca_coefficient_array=coef,
                                           pca_vectors=vectors, pca_mean=mean)
    luci2 = Luci.from_arrays(torch.from_numpy(pre.astype(np.float32)).to(luci.cube_final.device), sc.hdr_dict,
                             str(tmp_path), 'PRE', spectrum_axis=sc.axis32, device=luci.cube_final.device)
    vel2, broad2, flux2, ampl2 = luci2.fit_cube(*args, prior_values=pv)
    np.testing.assert_allclose(vel, vel2, atol=5e-2)
    np.testing.assert_allclose(flux, flux2, rtol=2e-3)
    np.testing.assert_allclose(broad, broad2, rtol=1e-3)
    with pytest.raises(Exception):
        luci.fit_cube(*args, prior_values=pv, bkgType='pca')                     # the PCA arrays are required


def test_freeze_mode_through_fit_cube(world):
    """initial_values = [velocity map, broadening map]: refit of the amplitudes with the kinematics of a previous fit
    held fixed (LuciBase.py:361-364, LuciFit.py:688-709)."""
    luci, sc, out = world
    v0, b0 = sc.priors()
    pv = (v0[1:8, 2:9], b0[1:8, 2:9])
    vel, broad, flux, ampl = luci.fit_cube(SN3, 'sincgauss', [1] * 5, [1] * 5, 2, 9, 1, 8, prior_values=pv)
    velf, broadf, fluxf, amplf = luci.fit_cube(SN3, 'sincgauss', [1] * 5, [1] * 5, 2, 9, 1, 8,
                                               initial_values=[vel[:, :, 0], broad[:, :, 0]])
    np.testing.assert_allclose(velf[:, :, 0], vel[:, :, 0], atol=1e-3)           # held
    np.testing.assert_allclose(broadf[:, :, 0], broad[:, :, 0], rtol=1e-5)
    np.testing.assert_allclose(amplf[:, :, 0], ampl[:, :, 0], rtol=2e-2)         # Halpha barely moves
    assert np.abs(fluxf[:, :, 1] * 3 / fluxf[:, :, 2] - 1).max() > 1e-3          # the [NII] tie is gone (no constraints)


def test_absorption_template_equals_fit_of_corrected_cube(world, tmp_path):
    """absorp (LuciBase.py:354-356): sky - absorp / median(absorp) * m + m, m = median of the central half."""
    from luci_b200.luci import Luci
    luci, sc, out = world
    cube = luci.cube_final.cpu().numpy().astype(np.float64)
    nx, ny, nz = cube.shape
    absorp = 1.0 + 0.05 * np.cos(np.linspace(0, 6, nz))
    cut = int(nz * 0.25)
    m = np.nanmedian(cube[:, :, cut:-cut], axis=2)
    pre = cube - (absorp / np.nanmedian(absorp))[None, None, :] * m[:, :, None] + m[:, :, None]
    v0, b0 = sc.priors()
    args = (SN3, 'sincgauss', [1] * 5, [1] * 5, 1, 11, 2, 9)
    pv = (v0[2:9, 1:11], b0[2:9, 1:11])
    vel, broad, flux, ampl = luci.fit_cube(*args, prior_values=pv, absorp=absorp)
    luci2 = Luci.from_arrays(torch.from_numpy(pre.astype(np.float32)).to(luci.cube_final.device), sc.hdr_dict,
                             str(tmp_path), 'ABS', spectrum_axis=sc.axis32, device=luci.cube_final.device)
    vel2, broad2, flux2, ampl2 = luci2.fit_cube(*args, prior_values=pv)
    np.testing.assert_allclose(vel, vel2, atol=5e-2)
    np.testing.assert_allclose(flux, flux2, rtol=2e-3)


def test_fit_wvt_bins_equal_fit_spectrum_region(world):
    """fit_wvt: one segmented reduction + one batch call == the reference's per-bin fit_spectrum_region loop
    (LuciBase.py:1383-1467), painted onto the bins' pixels; bins can come from the .npy files the reference writes."""
    luci, sc, out = world
    nx, ny = luci.cube_final.shape[:2]
    bin_map = np.full((nx, ny), -1, dtype=np.int64)
    b = 0
    for x0 in range(0, nx - 1, 3):
        for y0 in range(0, ny - 1, 2):
            bin_map[x0:x0 + 3, y0:y0 + 2] = b
            b += 1
    bin_map[0, 0] = -1                                                     # an unbinned pixel
    n_bins = b
    v0, b0 = sc.priors()
    vel, broad, flux, chi2, hdr = luci.fit_wvt(SN3, 'sincgauss', [1] * 5, [1] * 5, bin_map=bin_map,
                                               prior_values=(v0, b0), uncertainty_bool=True)
    assert vel.shape == (ny, nx, 5) and chi2.shape == (ny, nx)
    assert (vel[0, 0] == 0).all()                                           # unbinned pixel untouched
    for k in (0, n_bins // 2, n_bins - 1):
        mask = bin_map == k
        xs, ys = np.nonzero(mask)
        pv = (float(np.median(v0.T[mask])), float(np.median(b0.T[mask])))
        axis, sky, d = luci.fit_spectrum_region(SN3, 'sincgauss', [1] * 5, [1] * 5, mask, prior_values=pv,
                                                uncertainty_bool=True)
        for x, y in zip(xs, ys):
            np.testing.assert_allclose(vel[y, x], d['velocities'], atol=2e-3)
            np.testing.assert_allclose(flux[y, x], d['fluxes'], rtol=1e-4)
            np.testing.assert_allclose(chi2[y, x], d['chi2'], rtol=1e-4)
    assert os.path.exists(os.path.join(out, 'Luci_outputs', 'Velocity', 'OBJ_wvt_10_1_Halpha_velocity.fits'))
    # the same bins from the reference's on-disk layout
    folder = os.path.join(luci.output_dir, 'Numpy_Voronoi_Bins')
    os.makedirs(folder, exist_ok=True)
    for k in range(n_bins):
        np.save(os.path.join(folder, 'bool_bin_map_%i.npy' % k), bin_map == k)
    vel2 = luci.fit_wvt(SN3, 'sincgauss', [1] * 5, [1] * 5, prior_values=(v0, b0), uncertainty_bool=True)[0]
    np.testing.assert_array_equal(vel, vel2)
    # without prior_values (ML_bool=True): the priors of the bins are measured on the device, same optima
    vel3, _, flux3, _, _ = luci.fit_wvt(SN3, 'sincgauss', [1] * 5, [1] * 5, bin_map=bin_map)
    inside = bin_map.T >= 0
    assert (np.abs(vel3[inside] - vel[inside]).max(axis=1) < 0.05).mean() >= 0.95
    assert (np.abs(flux3[inside] / flux[inside] - 1).max(axis=1) < 2e-3).mean() >= 0.95


def test_fit_pixel_matches_fit_cube_and_binned_box(world):
    """fit_pixel (LuciBase.py:724-838): a single pixel equals the same spaxel of fit_cube; with binning the spectrum is
    the nansum of the [pixel - b, pixel + b) box."""
    luci, sc, out = world
    v0, b0 = sc.priors()
    x, y = 5, 4
    axis, sky, fd = luci.fit_pixel(SN3, 'sincgauss', [1] * 5, [1] * 5, x, y, prior_values=(v0[y, x], b0[y, x]),
                                   uncertainty_bool=True)
    host = luci.cube_final.cpu().numpy()
    np.testing.assert_array_equal(sky.astype(np.float32), host[x, y])
    assert axis.shape == sky.shape
    velc, broadc, fluxc, amplc = luci.fit_cube(SN3, 'sincgauss', [1] * 5, [1] * 5, x, x + 1, y, y + 1,
                                               prior_values=(v0[y:y + 1, x:x + 1], b0[y:y + 1, x:x + 1]),
                                               uncertainty_bool=True)
    np.testing.assert_array_equal(np.float32(fd['velocities']), velc[0, 0])
    np.testing.assert_array_equal(np.float32(fd['fluxes']), fluxc[0, 0])
    assert set(fd) >= {'fit_sol', 'fit_uncertainties', 'amplitudes', 'fluxes', 'flux_errors', 'chi2', 'velocities',
                       'sigmas', 'vels_errors', 'sigmas_errors', 'axis_step', 'corr', 'continuum', 'continuum_error',
                       'scale', 'vel_ml', 'broad_ml', 'fit_vector', 'fit_axis'}
    assert fd['fit_vector'].shape == axis.shape and fd['fit_sol'].shape == (16,)
    # binned box with a background: sky = nansum(box) - bkg * binning^2 (LuciBase.py:762-769)
    bkg = np.full(host.shape[2], 0.01e-17)
    _, sky2, fd2 = luci.fit_pixel(SN3, 'sincgauss', [1] * 5, [1] * 5, x, y, binning=2, bkg=bkg,
                                  prior_values=(v0[y, x], b0[y, x]))
    np.testing.assert_allclose(sky2, host[x - 2:x + 2, y - 2:y + 2].astype(np.float64).sum(axis=(0, 1)) - bkg * 4, rtol=1e-12)
    assert 8.0 < fd2['amplitudes'][0] / 1e-17 < 20.0 and fd2['chi2'] > 0


def test_extract_spectrum_and_region(world):
    """extract_spectrum / extract_spectrum_region (LuciBase.py:844-947)."""
    luci, sc, out = world
    host = luci.cube_final.cpu().numpy().astype(np.float64)
    axis, spec = luci.extract_spectrum(2, 6, 1, 5)
    np.testing.assert_allclose(spec, host[2:6, 1:5].sum(axis=(0, 1)), rtol=1e-12)
    np.testing.assert_array_equal(axis, luci.spectrum_axis)
    bkg = np.full(host.shape[2], 0.02e-17)
    _, specb = luci.extract_spectrum(2, 6, 1, 5, bkg=bkg)
    np.testing.assert_allclose(specb, host[2:6, 1:5].sum(axis=(0, 1)) - 16 * bkg, rtol=1e-12)
    # mean=True: the reference divides by a counter it increments once (LuciBase.py:887-889)
    _, specm = luci.extract_spectrum(2, 6, 1, 5, mean=True)
    np.testing.assert_allclose(specm, spec, rtol=1e-12)
    luci.strict_reference = False
    try:
        _, specm = luci.extract_spectrum(2, 6, 1, 5, mean=True)
        np.testing.assert_allclose(specm, spec / 16, rtol=1e-12)
    finally:
        luci.strict_reference = True
    _, spec2 = luci.extract_spectrum(0, 12, 0, 10, binning=2)
    np.testing.assert_allclose(spec2, host.sum(axis=(0, 1)), rtol=1e-5)
    mask = np.zeros((12, 10), bool)
    mask[1:4, 2:9] = True
    _, specr = luci.extract_spectrum_region(mask, mean=True)
    np.testing.assert_allclose(specr, host[1:4, 2:9].sum(axis=(0, 1)) / 21, rtol=1e-12)


def test_config1_sinc_strip_through_fit_region_map_for_map(tmp_path):
    """BASELINE config 1 as the reference ran it (oracle/make_golden.py::make_region_strip: a 100 x 8 LuciSim strip, five-line
    sinc, ``fit_region`` with a 70 % mask, every y-row through the reference's own row worker, maps assembled like
    LuciBase.py:706-718) through ``Luci.fit_region`` on the GPU, map for map:

      * the maps are the rows of the same 562 spaxels fitted through the C ABI directly, zeros outside the mask like the
        reference's;
      * those rows meet the fixture bar (every spaxel on the optimum of the reference's objective or a lower minimum);
      * wherever the reference's SLSQP answer sits on that optimum, the velocity / flux / chi2 MAPS agree with the
        reference's maps within the north_star tolerances."""
    from luci_b200 import engine
    from luci_b200.luci import Luci
    from tests.common import assert_fixture_parity, config_from_golden, golden_priors, split_rows
    g = load_golden('region_sinc_strip')
    dev = torch.device('cuda', 0)
    lines = [str(x) for x in g['lines']]
    hdr_dict = {'FILTER': 'SN3', 'STEP': int(g['delta_x']), 'STEPNB': int(g['n_steps']), 'ZPDINDEX': 0}
    luci = Luci.from_arrays(g['cube3d'], hdr_dict, str(tmp_path), 'STRIP', spectrum_axis=g['axis32'], device=dev)
    mask = np.asarray(g['mask'], bool)
    vel, broad, flux, chi2, m = luci.fit_region(lines, 'sinc', [1] * 5, [1] * 5, mask, prior_values=(g['v0map'], g['b0map']))
    assert vel.shape == (8, 100, 5) and chi2.shape == (8, 100)
    for ours, ref in ((vel, g['map_velocities']), (flux, g['map_fluxes']), (chi2, g['map_chi2'])):
        assert (ours[~mask.T] == 0).all() and (ref[~mask.T] == 0).all()
    # the same spaxels through the C ABI, in the fixture's (y, x) order
    cfg = config_from_golden(g)
    v0, b0 = golden_priors(g)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    res = engine.fit_batch(cfg, t(g['cube'].astype(np.float32)), t(g['theta']), t(v0), t(b0))
    rows, st = res['rows'].cpu().numpy(), res['status'].cpu().numpy()
    q = split_rows(rows, 5)
    y, x = g['yx'][:, 0], g['yx'][:, 1]
    np.testing.assert_allclose(vel[y, x], q['velocities'], rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(flux[y, x], q['fluxes'], rtol=1e-6)
    np.testing.assert_allclose(chi2[y, x], q['chi2'], rtol=1e-6)
    rep, opt, direct = assert_fixture_parity('region_sinc_strip', rows, st, g)
    assert opt['same'].mean() >= 0.99 and direct.mean() >= 0.8, (opt['same'].mean(), direct.mean())
    # map for map against the reference where the reference found its optimum
    yd, xd = y[direct], x[direct]
    assert np.abs(vel[yd, xd] - g['map_velocities'][yd, xd]).max() <= 0.5
    fr = g['map_fluxes'][yd, xd].astype(np.float64)
    den = np.maximum(np.abs(fr), 1e-3 * np.abs(fr).max(axis=1, keepdims=True))
    assert (np.abs(flux[yd, xd] - fr) / den).max() <= 1e-3
    cr = g['map_chi2'][yd, xd].astype(np.float64)
    assert ((chi2[yd, xd] - cr) / cr).max() <= 1e-4


def test_ds9_region_files_through_the_api(world, tmp_path):
    """``region='x.reg'`` (LuciBase.py:622-624, 917-919, 985-987 -> LuciConvenience.reg_to_mask): the ds9 file gives the
    same spectrum / maps as the boolean mask it describes, in pixel and in sky coordinates (the fixture's WCS is one
    degree per pixel around CRPIX (6, 5) = (10 deg, 20 deg))."""
    from luci_b200.regions import reg_to_mask
    luci, sc, out = world
    host = luci.cube_final.cpu().numpy().astype(np.float64)
    reg = tmp_path / 'r.reg'
    reg.write_text('# Region file format: DS9 version 4.1\nphysical\ncircle(6.2,5.1,2.6)\n')
    want = (np.arange(12)[:, None] - 5.2) ** 2 + (np.arange(10)[None, :] - 4.1) ** 2 <= 2.6 ** 2
    _, spec = luci.extract_spectrum_region(str(reg), mean=True)
    np.testing.assert_allclose(spec, host[want].sum(axis=0) / want.sum(), rtol=1e-12)
    v0, b0 = sc.priors()
    vel, broad, flux, chi2, m = luci.fit_region(SN3, 'sincgauss', [1] * 5, [1] * 5, str(reg), prior_values=(v0, b0))
    assert (m == want).all()
    velm, broadm, fluxm, chi2m, _ = luci.fit_region(SN3, 'sincgauss', [1] * 5, [1] * 5, want, prior_values=(v0, b0))
    np.testing.assert_array_equal(vel, velm)
    np.testing.assert_array_equal(chi2, chi2m)
    assert (chi2[want.T] > 0).all() and (chi2[~want.T] == 0).all()
    # the same circle in sky coordinates: centre = pixel (6.2, 5.1) of the TAN header, radius in degrees
    from luci_b200.regions import TanWCS
    lon, lat = TanWCS(luci.header).pix2world(6.2, 5.1)
    reg.write_text('fk5\ncircle(%.10f,%.10f,2.6)\n' % (lon, lat))
    sky = reg_to_mask(str(reg), luci.header, luci.cube_final.shape)
    assert (sky ^ want).sum() <= 1                          # one-degree pixels: the local scale is not exactly CDELT
    axis, sky_spec, d = luci.fit_spectrum_region(SN3, 'sincgauss', [1] * 5, [1] * 5, str(reg), prior_values=(float(v0.T[sky].mean()), float(b0.T[sky].mean())))
    np.testing.assert_allclose(sky_spec, host[sky].sum(axis=0), rtol=1e-12)
    assert d['chi2'] > 0


@pytest.mark.parametrize('filt,lines', [('SN2', ['Hbeta', 'OIII4959', 'OIII5007']), ('SN1', ['OII3726', 'OII3729'])])
def test_config3_sn1_sn2_gaussian_binning2(tmp_path, filt, lines):
    """BASELINE config 3: SN1 / SN2 cubes, gaussian model, binning=2 -- through fit_entire_cube with the CNN-like priors
    of the simulation and with the priors measured on the device (ML_bool=True, no prior_values)."""
    from luci_b200.luci import Luci
    from luci_b200.sim import SyntheticCube
    dev = torch.device('cuda', 0)
    L = len(lines)
    sc = SyntheticCube(12, 8, lines=lines, model='gaussian', filter_=filt, resolution=5000, seed=21,
                       broad_range=(60.0, 100.0), snr_range=(20.0, 100.0))
    cube = sc.synthesize(dev)
    # bins of 2 x 2 identical-velocity spaxels keep the binned truth simple: copy every even spaxel onto its bin
    cube = cube[0::2, 0::2].repeat_interleave(2, dim=0).repeat_interleave(2, dim=1).contiguous()
    luci = Luci.from_arrays(cube, sc.hdr_dict, str(tmp_path), 'C3', spectrum_axis=sc.axis32, device=dev)
    v0, b0 = sc.priors()
    vb, bb = v0[0::2, 0::2], b0[0::2, 0::2]
    vel, broad, flux, ampl = luci.fit_entire_cube(lines, 'gaussian', [1] * L, [1] * L, binning=2, prior_values=(vb, bb))
    assert vel.shape == (4, 6, L) and luci.last_fit_summary['n_converged'] == 24
    truth = sc.vel_fit_truth[0::2, 0::2].T
    chan_kms = 299792.0 * float(sc.axis32[1] - sc.axis32[0]) / float(sc.axis32[len(sc.axis32) // 2])
    assert np.median(np.abs(vel[:, :, 0] - truth)) < 0.6 * chan_kms          # LuciSim snaps lines to channels
    assert 3.0 < np.median(ampl[:, :, 0]) / 1e-17 < 5.0                      # nansum of 4 spaxels, no division
    vel_a, broad_a, flux_a, _ = luci.fit_entire_cube(lines, 'gaussian', [1] * L, [1] * L, binning=2)
    same = (np.abs(vel_a - vel).max(axis=2) < 0.1) & (np.abs(flux_a / flux - 1).max(axis=2) < 5e-3)
    assert same.mean() >= 0.9, (same.mean(), np.abs(vel_a - vel).max())
    assert os.path.exists(os.path.join(str(tmp_path), 'Luci_outputs', 'Velocity', 'C3_2_gaussian_%s_velocity.fits' % lines[0]))


def test_low_resolution_cube_with_a_three_sample_continuum_window(tmp_path):
    """R = 1000 SN2 (nz = 186, the resolution of the reference's own SN1/SN2 example cubes): the continuum window holds 3
    samples and the fit window 77; every scratch area must still be large enough and the fits finite."""
    from luci_b200.luci import Luci
    from luci_b200.sim import SyntheticCube
    dev = torch.device('cuda', 0)
    lines = ['Hbeta', 'OIII4959', 'OIII5007']
    sc = SyntheticCube(6, 5, lines=lines, model='gaussian', filter_='SN2', resolution=1000, seed=4,
                       broad_range=(250.0, 350.0), snr_range=(30.0, 100.0))
    assert sc.nz < 200
    luci = Luci.from_arrays(sc.synthesize(dev), sc.hdr_dict, str(tmp_path), 'LOWR', spectrum_axis=sc.axis32, device=dev)
    v0, b0 = sc.priors()
    for kw in (dict(prior_values=(v0, b0)), dict()):
        vel, broad, flux, ampl = luci.fit_cube(lines, 'gaussian', [1] * 3, [1] * 3, 0, 6, 0, 5, uncertainty_bool=True, **kw)
        assert np.isfinite(vel).all() and np.isfinite(flux).all() and (broad >= 0).all()
        assert luci.last_fit_summary['n_fitted'] == 30
    luci.create_deep_image()
    masks = luci.create_snr_map(0, 6, 0, 5, method=1)
    assert len(masks) == 4 and np.isfinite(luci.snr_map).all()


def _oracle_fit(sky, sc, theta, prior, **kw):
    """The oracle (restatement of LuciFit.Fit, pinned on the reference's goldens) on one float64 spectrum."""
    import warnings
    from oracle import luci_oracle as O
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return O.OracleFit(np.asarray(sky, dtype=np.float64), sc.axis32, 'sincgauss', SN3, [1] * 5, [1] * 5, theta=theta,
                           delta_x=sc.delta_x, n_steps=sc.n_steps, zpd_index=0, filter='SN3', nii_cons=True,
                           prior=prior, **kw).fit()


def _assert_fit_matches_oracle(d, o):
    """north_star bar: velocity 0.5 km/s, broadening / flux 1e-3, chi2 not worse by more than 1e-4."""
    np.testing.assert_allclose(d['velocities'], o['velocities'], atol=0.5)
    np.testing.assert_allclose(d['sigmas'], o['sigmas'], rtol=1e-3)
    np.testing.assert_allclose(d['fluxes'], o['fluxes'], rtol=1e-3)
    assert d['chi2'] <= o['chi2'] * (1 + 1e-4)


def test_fit_spectrum_region_matches_oracle(world):
    """fit_spectrum_region (LuciBase.py:949-1046) against the oracle on the same integrated spectrum: plain fit,
    uncertainties, freeze (initial_values) and n_stoch restarts."""
    luci, sc, out = world
    mask = np.zeros((12, 10), bool)
    mask[3:8, 2:7] = True
    mask[5, 4] = False
    host = luci.cube_final.cpu().numpy().astype(np.float64)
    sky_ref = host[mask].sum(axis=0)
    theta = float(luci.interferometer_theta[-1, -1])                       # LuciBase.py:1035: the last pixel of the loops
    prior = (float(sc.vel_fit_truth[3:8, 2:7].mean()) + 6.0, 42.0)
    axis, sky, d = luci.fit_spectrum_region(SN3, 'sincgauss', [1] * 5, [1] * 5, mask, prior_values=prior, uncertainty_bool=True)
    np.testing.assert_allclose(sky, sky_ref, rtol=1e-12)
    o = _oracle_fit(sky_ref, sc, theta, prior, uncertainty_bool=True)
    _assert_fit_matches_oracle(d, o)
    np.testing.assert_allclose(d['vels_errors'], o['vels_errors'], rtol=5e-2)
    np.testing.assert_allclose(d['flux_errors'], o['flux_errors'], rtol=5e-2)
    # mean=True and a background (:1024-1027)
    bkg = np.full(host.shape[2], 0.01e-17)
    _, sky_m, d_m = luci.fit_spectrum_region(SN3, 'sincgauss', [1] * 5, [1] * 5, mask, prior_values=prior, mean=True, bkg=bkg)
    n = int(mask.sum())
    np.testing.assert_allclose(sky_m, sky_ref / n - bkg * n, rtol=1e-9)
    # freeze: velocity and broadening held at initial_values, amplitudes + continuum fitted (LuciFit.py:688-709)
    iv = [prior[0], prior[1]]
    _, _, d_f = luci.fit_spectrum_region(SN3, 'sincgauss', [1] * 5, [1] * 5, mask, initial_values=iv)
    o_f = _oracle_fit(sky_ref, sc, theta, None, initial_values=iv)
    np.testing.assert_allclose(d_f['velocities'], o_f['velocities'], atol=1e-2)
    np.testing.assert_allclose(d_f['sigmas'], o_f['sigmas'], rtol=1e-4)
    np.testing.assert_allclose(d_f['fluxes'], o_f['fluxes'], rtol=2e-3)
    # restarts (LuciFit.py:662-687): never worse than the single start, same optimum here
    np.random.seed(5)
    _, _, d_s = luci.fit_spectrum_region(SN3, 'sincgauss', [1] * 5, [1] * 5, mask, prior_values=prior, n_stoch=3)
    assert d_s['chi2'] <= d['chi2'] * (1 + 1e-6)
    _assert_fit_matches_oracle(d_s, o)


def test_fit_wvt_bins_match_oracle_with_restarts_and_freeze(world):
    """fit_wvt: bins against the oracle on the summed spectrum of the bin (what the reference's per-bin
    fit_spectrum_region call fits, LuciBase.py:1434-1447); n_stoch restarts; freeze from initial-value maps."""
    luci, sc, out = world
    nx, ny = luci.cube_final.shape[:2]
    bin_map = np.full((nx, ny), -1, dtype=np.int64)
    bin_map[0:4, 0:5] = 0
    bin_map[4:9, 2:8] = 1
    bin_map[9:12, 5:10] = 2
    bin_map[6, 5] = -1
    host = luci.cube_final.cpu().numpy().astype(np.float64)
    theta = float(luci.interferometer_theta[-1, -1])
    v0, b0 = sc.priors()
    np.random.seed(3)
    vel, broad, flux, chi2, hdr = luci.fit_wvt(SN3, 'sincgauss', [1] * 5, [1] * 5, bin_map=bin_map, prior_values=(v0, b0), n_stoch=2)
    rows = luci.last_fit_maps['bin_rows']
    for k in range(3):
        m = bin_map == k
        prior = (float(np.median(v0.T[m])), float(np.median(b0.T[m])))
        o = _oracle_fit(host[m].sum(axis=0), sc, theta, prior)
        x, y = np.argwhere(m)[0]
        d = {'velocities': vel[y, x], 'sigmas': broad[y, x], 'fluxes': flux[y, x], 'chi2': chi2[y, x]}
        _assert_fit_matches_oracle(d, o)
    assert rows.shape == (3, 40)
    # freeze: [x, y]-indexed maps, each bin held at the values of its last pixel (LuciBase.py:1437-1442)
    vi = np.asarray(v0.T, dtype=np.float32).copy()
    bi = np.asarray(b0.T, dtype=np.float32).copy()
    velf, broadf, fluxf, _, _ = luci.fit_wvt(SN3, 'sincgauss', [1] * 5, [1] * 5, bin_map=bin_map, initial_values=[vi, bi])
    for k in range(3):
        m = bin_map == k
        xl, yl = np.argwhere(m)[-1]
        x, y = np.argwhere(m)[0]
        np.testing.assert_allclose(velf[y, x], vi[xl, yl], atol=1e-2)
        np.testing.assert_allclose(broadf[y, x], bi[xl, yl], rtol=1e-4)
        o_f = _oracle_fit(host[m].sum(axis=0), sc, theta, None, initial_values=[float(vi[xl, yl]), float(bi[xl, yl])])
        np.testing.assert_allclose(fluxf[y, x], o_f['fluxes'], rtol=2e-3)


def test_fit_pixel_with_nan_channels_matches_oracle(tmp_path):
    """NaN channels are dropped like the reference's good_sky_inds (LuciBase.py:801-803) instead of being refused."""
    from luci_b200.luci import Luci
    from luci_b200.sim import SyntheticCube
    dev = torch.device('cuda', 0)
    sc = SyntheticCube(3, 3, model='sincgauss', resolution=5000, seed=8, snr_range=(40.0, 100.0))
    cube = sc.synthesize(dev)
    lo, hi = 14750, 15400
    inside = np.flatnonzero((sc.axis32 > lo + 60) & (sc.axis32 < hi - 60))
    cube[1, 1, int(inside[7])] = float('nan')
    cube[1, 1, int(inside[40])] = float('nan')
    cube[1, 1, 5] = float('nan')
    luci = Luci.from_arrays(cube, sc.hdr_dict, str(tmp_path), 'NAN', spectrum_axis=sc.axis32, device=dev)
    v0, b0 = sc.priors()
    axis, sky, d = luci.fit_pixel(SN3, 'sincgauss', [1] * 5, [1] * 5, 1, 1, prior_values=(v0[1, 1], b0[1, 1]))
    assert np.isnan(sky).sum() == 3
    good = ~np.isnan(sky)
    import warnings
    from oracle import luci_oracle as O
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        o = O.OracleFit(sky[good], sc.axis32[good], 'sincgauss', SN3, [1] * 5, [1] * 5, theta=sc.theta, delta_x=sc.delta_x,
                        n_steps=sc.n_steps, zpd_index=0, filter='SN3', nii_cons=True,
                        prior=(float(v0[1, 1]), float(b0[1, 1]))).fit()
    _assert_fit_matches_oracle(d, o)


def test_segment_sum_against_numpy():
    """lucifit_segment_sum: ordered and unordered segment lists, skipped rows, nansum on / off, long spectra."""
    from luci_b200 import engine
    dev = torch.device('cuda', 0)
    rng = np.random.default_rng(2)
    for n_rows, nz, n_seg in ((500, 841, 7), (64, 1300, 1), (3000, 97, 40)):
        cube = rng.normal(size=(n_rows, nz)).astype(np.float32)
        cube[rng.integers(0, n_rows, 20), rng.integers(0, nz, 20)] = np.nan
        rows = rng.permutation(n_rows)[: n_rows * 3 // 4].astype(np.int64)
        seg = rng.integers(-1, n_seg, rows.size).astype(np.int32)
        order = np.argsort(seg, kind='stable')
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        for r, s in ((rows, seg), (rows[order], seg[order])):
            for nansum in (True, False):
                got = engine.segment_sum(t(cube), t(r), t(s), n_seg, nansum=nansum).cpu().numpy()
                want = np.zeros((n_seg, nz))
                c64 = cube.astype(np.float64)
                if nansum:
                    c64 = np.nan_to_num(c64, nan=0.0)
                np.add.at(want, s[s >= 0], c64[r[s >= 0]])
                np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-12, equal_nan=True)
        whole = engine.segment_sum(t(cube), nansum=True).cpu().numpy()
        np.testing.assert_allclose(whole[0], np.nansum(cube.astype(np.float64), axis=0), rtol=1e-11, atol=1e-12)


def test_staged_big_endian_output_equals_plain_path(world):
    """Large fits fetch their maps through pinned staging and write the FITS files from the byte-reversed copy the
    device made (lucifit_pack_maps): same arrays, byte-identical files."""
    import glob
    luci, sc, out = world
    v0, b0 = sc.priors()
    args = (SN3, 'sincgauss', [1] * 5, [1] * 5, 0, 12, 0, 10)
    base = os.path.join(out, 'Luci_outputs')
    plain = luci.fit_cube(*args, prior_values=(v0, b0), uncertainty_bool=True)
    files = sorted(glob.glob(os.path.join(base, '**', 'OBJ_sincgauss_*.fits'), recursive=True))
    assert len(files) == 38
    ref_bytes = {f: open(f, 'rb').read() for f in files}
    ref_maps = {k: v.copy() for k, v in luci.last_fit_maps.items()}
    for f in files:
        os.remove(f)
    luci.STAGE_MIN_VALUES = 1
    try:
        staged = luci.fit_cube(*args, prior_values=(v0, b0), uncertainty_bool=True)
        assert luci._last_be is not None
    finally:
        del luci.STAGE_MIN_VALUES
    for a, b in zip(plain, staged):
        np.testing.assert_array_equal(a, b)
    for k, v in ref_maps.items():
        np.testing.assert_array_equal(v, luci.last_fit_maps[k], err_msg=k)
    for f in files:
        assert open(f, 'rb').read() == ref_bytes[f], f
    # the returned arrays do not alias the staging buffers: a second fit leaves them alone
    keep = staged[0].copy()
    luci.STAGE_MIN_VALUES = 1
    try:
        luci.fit_cube(SN3, 'sincgauss', [1] * 5, [1] * 5, 0, 12, 0, 10, prior_values=(v0 + 3.0, b0))
    finally:
        del luci.STAGE_MIN_VALUES
    np.testing.assert_array_equal(staged[0], keep)
