"""GPU-less checks of the drop-in boundary: the C-ABI library loads and exports every symbol declared in
include/lucifit.h, rejects bad configurations with a message (no CUDA call is made), and the host glue mirrors
the reference's argument checks, windows and file layout."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from luci_b200 import _capi, constants, fitsio, host_ops, sim
from luci_b200._capi import FitConfig
from oracle import luci_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SN3 = ['Halpha', 'NII6548', 'NII6583', 'SII6716', 'SII6731']


def _axis():
    axis, n, dx, w = O.sim_axis('SN3', 5000)
    return axis.astype(np.float32), n, dx


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'lucifit.h')).read()
    declared = sorted(set(re.findall(r'\b(lucifit_[a-z_]+)\s*\(', hdr)))
    assert set(declared) == set(_capi.EXPORTS), declared
    lib = _capi.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert b'sm_100a' in lib.lucifit_version()
    assert lib.lucifit_row_floats(5) == 40


def test_config_validation_happens_before_any_cuda_call():
    lib = _capi.lib()
    ax, n, dx = _axis()
    cfg = FitConfig('sincgauss', SN3, [1] * 5, [1] * 5, ax, 'SN3', dx, n, 0)
    need = C.c_size_t(0)
    assert lib.lucifit_workspace_bytes(cfg.byref(), 10, C.byref(need)) == 0 and need.value >= 8 * len(ax)
    cfg.c.model = 7
    assert lib.lucifit_workspace_bytes(cfg.byref(), 10, C.byref(need)) == -1
    assert b'model' in lib.lucifit_last_error()
    cfg.c.model = 2
    cfg.c.fit_hi = cfg.c.fit_lo + 3
    assert lib.lucifit_fit_batch(cfg.byref(), 1, None, None, None, None, None, None, None, None, None, None, None, 0, None) == -1
    assert b'fit window' in lib.lucifit_last_error()
    # un-instantiated tie pattern: 7 lines in 3 groups
    lines7 = SN3 + ['Halpha', 'NII6583']
    cfg7 = FitConfig('gaussian', lines7, [1, 1, 1, 2, 2, 3, 3], [1] * 7, ax, 'SN3', dx, n, 0, nii_cons=False)
    dummy = (C.c_float * 256)()
    p = C.c_void_p((C.addressof(dummy) + 255) & ~255)     # 256-byte aligned address inside the dummy buffer
    rc = lib.lucifit_fit_batch(cfg7.byref(), 1, p, None, p, None, None, None, p, None, None, p, p, 1 << 20, None)
    assert rc == -1 and b'not instantiated' in lib.lucifit_last_error()
    assert lib.lucifit_deep_image(None, 1, 10, 0, None, None) == -1
    assert lib.lucifit_snr_map(p, 1, 10, 3, 0, 1, 0, 1, p, None, None) == -1


def test_fitconfig_mirrors_reference_checks():
    ax, n, dx = _axis()
    with pytest.raises(Exception, match='fitting function'):
        FitConfig('lorentz', SN3, [1] * 5, [1] * 5, ax)
    with pytest.raises(Exception, match='line name'):
        FitConfig('sinc', ['Hgamma'], [1], [1], ax)
    with pytest.raises(Exception, match='vel_rel has 2 arguments'):
        FitConfig('sinc', SN3, [1, 1], [1] * 5, ax)
    with pytest.raises(Exception, match='sigma_rel has 1 arguments'):
        FitConfig('sinc', SN3, [1] * 5, [1], ax)
    cfg = FitConfig('gauss', SN3, [1] * 5, [1] * 5, ax, 'SN3', dx, n, 0)
    assert cfg.model_type == 'gaussian' and cfg.c.model == 0
    # windows are np.argmin(|axis - bound|) like LuciFit.py:273-274, 327-328, 476-477
    f = O.OracleFit(np.ones(len(ax)), ax, 'gaussian', SN3, [1] * 5, [1] * 5, theta=11.96, delta_x=dx, n_steps=n, zpd_index=0)
    assert (cfg.c.fit_lo, cfg.c.fit_hi) == (f.lo, f.hi)
    assert cfg.c.noise_hi - cfg.c.noise_lo == 11 and cfg.c.cont_hi - cfg.c.cont_lo == 41
    assert (cfg.c.nii6548_idx, cfg.c.nii6583_idx) == (1, 2)
    assert FitConfig('sinc', SN3, [1] * 5, [1] * 5, ax, nii_cons=False).c.nii6548_idx == -1


def test_constants_tables_match_oracle_tables():
    for flt in ('SN1', 'SN2', 'SN3'):
        fw, nw, cw = O.FILTER_WINDOWS[flt]
        assert constants.fit_window(flt) == fw and constants.noise_window(flt) == nw
        assert constants.continuum_window(flt) == cw
    assert constants.snr_windows('SN3') == O.SNR_WINDOWS['SN3']
    assert constants.snr_windows('SN2', ['OIII']) == O.SNR_WINDOWS['SN2_OIII']
    assert constants.LINE_DICT == O.LINE_DICT and constants.SIM_FILTERS == O.SIM_FILTERS


def test_sim_axis_matches_lucisim_restatement():
    for flt, R in (('SN3', 5000), ('SN2', 5000), ('SN1', 1000)):
        a, n, dx = sim.sim_axis(flt, R)
        ao, no, dxo, w = O.sim_axis(flt, R)
        np.testing.assert_array_equal(a, ao)
        assert (n, dx) == (no, dxo)
    a841, n841, _ = sim.sim_axis('SN3', n_steps=841)
    assert len(a841) == 841


def test_fits_roundtrip_and_save_layout(tmp_path):
    data = np.arange(12, dtype=np.float32).reshape(3, 4)
    hdr = {'CRPIX1': 10.5, 'CRVAL1': 308.71, 'CTYPE1': 'RA---TAN', 'WCSAXES': 2, 'EQUINOX': 2000.0}
    p = str(tmp_path / 'a.fits')
    fitsio.write_fits(p, data, hdr)
    assert os.path.getsize(p) % 2880 == 0
    back, h = fitsio.read_fits(p)
    np.testing.assert_array_equal(back, data)
    assert h['BITPIX'] == -32 and h['NAXIS1'] == 4 and h['NAXIS2'] == 3       # maps are (ny, nx): NAXIS1 = nx
    assert h['CRPIX1'] == 10.5 and h['CTYPE1'] == 'RA---TAN' and h['WCSAXES'] == 2
    L, ny, nx = 2, 3, 4
    m3 = np.zeros((ny, nx, L), np.float32)
    m2 = np.zeros((ny, nx), np.float32)
    out = str(tmp_path)
    fitsio.save_fits(out, 'OBJ', ['Halpha', 'Halpha'], m3, m3, m3, m3, m3, m3, m3, m2, m2, m2, hdr, 2, fit_function='sincgauss')
    for rel in ('Amplitudes/OBJ_2_sincgauss_Halpha_Amplitude.fits', 'Fluxes/OBJ_2_sincgauss_Halpha_Flux.fits',
                'Fluxes/OBJ_2_sincgauss_Halpha_Flux_err.fits', 'Velocity/OBJ_2_sincgauss_Halpha_velocity.fits',
                'Velocity/OBJ_2_sincgauss_Halpha_velocity_err.fits', 'Broadening/OBJ_2_sincgauss_Halpha_broadening.fits',
                'Broadening/OBJ_2_sincgauss_Halpha_broadening_err.fits', 'Velocity/OBJ_2_sincgauss_Halpha_2_velocity.fits',
                'OBJ_2_sincgauss_Chi2.fits', 'OBJ_2_sincgauss_continuum.fits', 'OBJ_2_sincgauss_continuum_error.fits'):
        assert os.path.exists(os.path.join(out, rel)), rel


def test_bin_mask_matches_oracle():
    rng = np.random.default_rng(3)
    mask = rng.uniform(size=(23, 14)) > 0.7
    np.testing.assert_array_equal(host_ops.bin_mask(mask, 2, 0, 23, 0, 14), O.bin_mask(mask, 2, 0, 23, 0, 14))
    np.testing.assert_array_equal(host_ops.bin_mask(mask, 3, 0, 23, 0, 14), O.bin_mask(mask, 3, 0, 23, 0, 14))


def test_compute_entry_points_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from luci_b200 import engine
    ax, n, dx = _axis()
    cfg = FitConfig('sinc', SN3, [1] * 5, [1] * 5, ax, 'SN3', dx, n, 0)
    with pytest.raises(TypeError, match='no CPU path'):
        engine.fit_batch(cfg, torch.zeros(2, len(ax)), torch.zeros(2))
    with pytest.raises(TypeError, match='no CPU path'):
        engine.deep_image(torch.zeros(2, len(ax)))


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under luci_b200/ (Python or C++/CUDA) may import, include or load it."""
    import re
    pkg = os.path.join(ROOT, 'luci_b200')
    bad = []
    for dirpath, _, files in os.walk(pkg):
        if '_obj' in dirpath or '__pycache__' in dirpath:
            continue
        for f in files:
            if not f.endswith(('.py', '.cu', '.cuh', '.hpp', '.h')):
                continue
            src = open(os.path.join(dirpath, f), errors='ignore').read()
            if re.search(r'^\s*(from|import)\s+oracle\b', src, re.M) or re.search(r'#include\s+".*oracle', src) \
                    or 'hostemu' in src and f.endswith('.py'):
                bad.append(os.path.join(dirpath, f))
    assert not bad, bad
