"""HDF5 ingest (luci_b200/hdf5io.py): Luci.__init__ / read_in_cube and the LuciUtility readers (LuciBase.py:51-145,
LUCI/LuciUtility.py:88-140, 168-260, 306-316).  h5py is not installed here, so the files are stand-ins with the
mapping interface h5py exposes (``file.attrs``, ``file[name][...]``); the reference's own helper functions are called
on the very same stand-ins when /root/reference is present."""
import os
import sys
import types

import numpy as np
import pytest

from luci_b200 import hdf5io


class FakeH5(dict):
    def __init__(self, attrs, items):
        super().__init__(items)
        self.attrs = attrs


def old_style_file(dimx=13, dimy=11, dimz=24, quad_nb=9, seed=0, filter_name='SN3', stepnb=None):
    rng = np.random.default_rng(seed)
    cube = rng.normal(1e-17, 2e-17, (dimx, dimy, dimz))
    cube[1, 2, 3] = np.nan; cube[5, 5, 0] = np.inf; cube[dimx - 6, 1, dimz - 1] = -np.inf
    cube[2, 2, 2] = -5e-16; cube[dimx - 1, dimy - 1, dimz - 1] = 3e-9; cube[0, 0, 0] = -0.9e-16      # the last one is inside the box
    items = {}
    for q in range(quad_nb):
        x0, x1, y0, y1 = hdf5io.get_quadrant_dims(q, quad_nb, dimx, dimy)
        items['quad00%i' % q] = {'data': cube[x0:x1, y0:y1, :].copy()}
    rows = [(b'STEP', b'2943.0', b'', b'float'), (b'STEPNB', str(stepnb or dimz).encode(), b'', b'int'), (b'ZPDINDEX', b'3', b'', b'int'),
            (b'FILTER', filter_name.encode(), b'', b'str'), (b'CRVAL3', b'14500.0', b'', b'float'), (b'CDELT3', b'45.0', b'', b'float'),
            (b'CALIBNM', b'543.5', b'', b'float'), (b'NAXIS', b'3', b'', b'int'), (b'CRPIX1', b'6.0', b'', b'float'),
            (b'CRPIX2', b'5.5', b'', b'float'), (b'CRVAL1', b'10.5', b'', b'float'), (b'CRVAL2', b'-3.25', b'', b'float'),
            (b'CTYPE1', b'RA---TAN', b'', b'str'), (b'CTYPE2', b'DEC--TAN', b'', b'str'), (b'CD1_1', b'-8.9e-05', b'', b'float'),
            (b'CD1_2', b'0.0', b'', b'float'), (b'CD2_1', b'0.0', b'', b'float'), (b'CD2_2', b'8.9e-05', b'', b'float'),
            (b'OBJECT', b'M33_Field7', b'', b'str')]
    items['header'] = np.array(rows, dtype=object)
    items['calib_map'] = 543.5 / np.cos(np.deg2rad(rng.uniform(11.8, 19.6, (dimx, dimy))))
    attrs = {'quad_nb': quad_nb, 'dimx': dimx, 'dimy': dimy, 'dimz': dimz}
    return FakeH5(attrs, items), cube


def reference_clamp(cube):
    """The three mask passes of LuciBase.py:126-129."""
    out = cube.copy()
    out[(np.isfinite(out) == False)] = 1e-22                                              # noqa: E712
    out[(out < -1e-16)] = 1e-22
    out[(out > 1e-9)] = 1e-22
    return out


@pytest.mark.parametrize('dims', [(2048, 2064, 9), (13, 11, 9), (10, 10, 4), (7, 9, 1), (100, 33, 16)])
def test_quadrants_tile_the_cube_exactly(dims):
    dimx, dimy, quad_nb = dims
    cover = np.zeros((dimx, dimy), int)
    for q in range(quad_nb):
        x0, x1, y0, y1 = hdf5io.get_quadrant_dims(q, quad_nb, dimx, dimy)
        cover[x0:x1, y0:y1] += 1
    assert (cover == 1).all()
    with pytest.raises(Exception):
        hdf5io.get_quadrant_dims(quad_nb, quad_nb, dimx, dimy)


def test_header_dict_old_and_new_layout():
    f, _ = old_style_file()
    h = hdf5io.header_dict(f)
    assert h['STEP'] == 2943.0 and h['STEPNB'] == 24 and isinstance(h['STEPNB'], int) and h['ZPDINDEX'] == 3
    assert h['FILTER'] == 'SN3' and h['CRVAL3'] == 14500.0 and h['CTYPE3'] == 'WAVE-SIP' and h['CUNIT3'] == 'm'
    assert h['NAXIS1'] == 2048 and h['NAXIS2'] == 2064 and h['NAXIS'] == 3
    assert h['OBJECT'] == 'M33_Field7'
    w = hdf5io.wcs_header(h)
    assert w['CRPIX1'] == 6.0 and w['CTYPE2'] == 'DEC--TAN' and w['PC1_1'] == -8.9e-05 and w['CDELT1'] == 1.0
    assert not any(k.startswith('CD1') or k.startswith('CD2') for k in w)
    new = FakeH5({'step_nb': np.int64(842), 'zpd_index': np.int64(169), 'filter_name': 'SN2', 'axis_min': np.float64(19000.0),
                  'axis_step': np.float64(2.5), 'STEP': np.float64(2943.0), 'NAXIS': np.int64(3), 'nm_laser': np.float64(543.5),
                  'reduced_cube_path': '/some/where', 'flambda': np.ones(3), 'wcs_rotation': np.float64(1.0)}, {})
    h = hdf5io.header_dict(new)
    assert h['STEPNB'] == 842 and h['ZPDINDEX'] == 169 and h['FILTER'] == 'SN2' and h['CRVAL3'] == 19000.0 and h['CDELT3'] == 2.5
    assert h['reduced_cube_path'] == '' and h['flambda'].shape == (3,)


def test_interferometer_angles_and_transmission(tmp_path):
    from scipy import interpolate
    f, _ = old_style_file()
    h = hdf5io.header_dict(f)
    th = hdf5io.interferometer_angles(f, h)
    np.testing.assert_allclose(th, np.rad2deg(np.arccos(np.float32(543.5) / f['calib_map'])), rtol=0, atol=0)
    assert th.shape == (13, 11) and 11.7 < th.min() and th.max() < 19.7
    # transmission: interp1d(kind='slinear', fill_value='extrapolate') of the percent column / 100
    ref_dat = '/root/reference/Data/SN3_filter.dat'
    os.makedirs(tmp_path / 'Data')
    if os.path.exists(ref_dat):
        table = np.loadtxt(ref_dat)
    else:
        x = np.linspace(14400, 15700, 300)
        table = np.stack([x, 90 * np.exp(-((x - 15050) / 400.0) ** 8)], axis=1)
    np.savetxt(tmp_path / 'Data' / 'SN3_filter.dat', table)
    axis = np.linspace(14300.0, 15800.0, 777).astype(np.float32)        # beyond both ends: extrapolation
    got = hdf5io.read_in_transmission(str(tmp_path), h, axis)
    want = interpolate.interp1d(table[:, 0], [v / 100 for v in table[:, 1]], kind='slinear', fill_value='extrapolate')(axis)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15)


@pytest.mark.reference
def test_helpers_match_the_reference_functions():
    from oracle import ref_shim
    ref_shim.install()
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    import importlib
    LU = importlib.import_module('LUCI.LuciUtility')
    for dimx, dimy, quad_nb in ((2048, 2064, 9), (13, 11, 9), (100, 33, 16), (10, 10, 4)):
        for q in range(quad_nb):
            assert hdf5io.get_quadrant_dims(q, quad_nb, dimx, dimy) == LU.get_quadrant_dims(q, quad_nb, dimx, dimy)
    f, _ = old_style_file()
    h = hdf5io.header_dict(f)
    np.testing.assert_array_equal(hdf5io.interferometer_angles(f, h), LU.get_interferometer_angles(f, h))
    a, au = LU.spectrum_axis_func(h, 0.0012)
    from luci_b200.luci import spectrum_axis_func
    b, bu = spectrum_axis_func(h, 0.0012)
    np.testing.assert_array_equal(a, b); np.testing.assert_array_equal(au, bu)
    t_ref = LU.read_in_transmission(ref_shim.REFERENCE_ROOT, h, au)
    np.testing.assert_allclose(hdf5io.read_in_transmission(ref_shim.REFERENCE_ROOT, h, au), t_ref, rtol=1e-12, atol=1e-15)


def test_open_without_h5py_says_so():
    if 'h5py' in sys.modules or __import__('importlib').util.find_spec('h5py') is not None:
        pytest.skip('h5py is installed')
    with pytest.raises(ImportError, match='from_arrays'):
        hdf5io.open_hdf5('/nowhere/cube')


# ----------------------------------------------------------------------------------------------------- GPU tier
@pytest.mark.gpu
def test_gpu_read_cube_assembles_and_clamps_like_the_reference():
    import torch
    dev = torch.device('cuda', 0)
    for dims in ((13, 11, 24, 9), (16, 16, 7, 4), (9, 8, 31, 1)):
        f, cube = old_style_file(*dims)
        got = hdf5io.read_cube(f, dev).cpu().numpy()
        want = reference_clamp(cube).astype(np.float32)
        np.testing.assert_array_equal(got, want)
        assert (got == np.float32(1e-22)).sum() >= 5
    # a view that is not 16-byte aligned goes through the scalar head / tail of the kernel
    from luci_b200 import engine
    x = torch.full((1003,), float('nan'), device=dev)
    engine.sanitize_cube(x[1:1002])
    out = x.cpu().numpy()
    assert np.isnan(out[0]) and np.isnan(out[-1]) and (out[1:1002] == np.float32(1e-22)).all()
    # single-dataset layout: real part, unclamped
    g = FakeH5({}, {'data': (np.arange(24.0).reshape(2, 3, 4) + 1j)})
    np.testing.assert_array_equal(hdf5io.read_cube(g, dev).cpu().numpy(), np.arange(24.0, dtype=np.float32).reshape(2, 3, 4))


@pytest.mark.gpu
def test_gpu_luci_constructor_reads_hdf5_and_fits(tmp_path, monkeypatch):
    """Luci(Luci_path, cube_path, ...) end to end on a stand-in file: cube, header, theta, axis, transmission, one fit."""
    import torch
    from luci_b200.luci import Luci
    from luci_b200.sim import SN3_LINES, SyntheticCube
    dev = torch.device('cuda', 0)
    sc = SyntheticCube(6, 6, model='sincgauss', resolution=5000, seed=5, snr_range=(30.0, 100.0))
    host = sc.synthesize(dev).cpu().numpy().astype(np.float64)
    nz = host.shape[2]
    f, _ = old_style_file(6, 6, nz, 9, filter_name='SN3')
    for q in range(9):
        x0, x1, y0, y1 = hdf5io.get_quadrant_dims(q, 9, 6, 6)
        f['quad00%i' % q] = {'data': host[x0:x1, y0:y1, :].copy()}
    f['quad004']['data'][0, 0, 5] = np.inf
    step = (float(sc.axis32[-1]) - float(sc.axis32[0])) / (nz - 1)
    rows = [r for r in f['header'] if r[0] not in (b'CRVAL3', b'CDELT3', b'STEP', b'ZPDINDEX')]
    # LuciUtility.spectrum_axis_func uses linspace(start, start + n * step, n): spacing n / (n - 1) * CDELT3
    rows += [(b'CRVAL3', repr(float(sc.axis32[0])).encode(), b'', b'float'),
             (b'CDELT3', repr(step * (nz - 1) / nz).encode(), b'', b'float'),
             (b'STEP', repr(float(sc.hdr_dict['STEP'])).encode(), b'', b'float'), (b'ZPDINDEX', b'0', b'', b'int')]
    f['header'] = np.array(rows, dtype=object)
    f['calib_map'] = np.full((6, 6), 543.5 / np.cos(np.deg2rad(11.96)))
    fake = types.ModuleType('h5py')
    fake.File = lambda path, mode='r': f if path.endswith('cube.hdf5') else (_ for _ in ()).throw(OSError(path))
    monkeypatch.setitem(sys.modules, 'h5py', fake)
    os.makedirs(tmp_path / 'LUCI' / 'Data')
    x = np.linspace(14000, 16000, 200)
    np.savetxt(tmp_path / 'LUCI' / 'Data' / 'SN3_filter.dat', np.stack([x, np.full_like(x, 95.0)], axis=1))
    luci = Luci(str(tmp_path / 'LUCI'), str(tmp_path / 'cube'), str(tmp_path), 'OBJ', 0.0, 5000, device=dev)
    assert tuple(luci.cube_final.shape) == (6, 6, nz) and luci.cube_final.dtype == torch.float32
    assert luci.filter == 'SN3' and luci.step_nb == nz and luci.zpd_index == 0
    np.testing.assert_allclose(luci.interferometer_theta, 11.96, atol=1e-3)
    np.testing.assert_allclose(luci.spectrum_axis, sc.axis32, atol=0.01)          # cm^-1 (0.2 km/s)
    np.testing.assert_allclose(luci.transmission_interpolated, 0.95, rtol=1e-12)
    got = luci.cube_final.cpu().numpy()
    assert got[2, 2, 5] == np.float32(1e-22)                      # quadrant 4 of the 3 x 3 tiling starts at pixel (2, 2)
    vel, broad, flux, ampl = luci.fit_cube(SN3_LINES, 'sincgauss', [1] * 5, [1] * 5, 0, 6, 0, 6)
    assert vel.shape == (6, 6, 5) and np.isfinite(vel).all()
    truth = sc.vel_fit_truth.T
    assert np.median(np.abs(vel[:, :, 0] - truth)) < 25.0
