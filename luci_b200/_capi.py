"""ctypes binding of ``liblucifit.so`` (the C ABI declared in ``include/lucifit.h``).

The library is the only compute backend: if it is missing the import of :mod:`luci_b200` still works (so that
host-only helpers and CPU tests can run) but every compute entry point raises :class:`LucifitUnavailable`.
There is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from .constants import LINE_DICT, MODEL_CODE, continuum_window, fit_window, noise_window

MAX_LINES = 8
_HERE = os.path.dirname(os.path.abspath(__file__))
# LUCIFIT_LIB: developer override (kernel experiments build variant libraries next to the shipped one)
LIB_PATH = os.environ.get('LUCIFIT_LIB') or os.path.join(_HERE, 'liblucifit.so')


class LucifitUnavailable(RuntimeError):
    pass


class LucifitError(RuntimeError):
    pass


class lucifit_config(C.Structure):
    _fields_ = [
        ('model', C.c_int), ('n_lines', C.c_int),
        ('line_nm', C.c_double * MAX_LINES),
        ('vel_group', C.c_int * MAX_LINES), ('sigma_group', C.c_int * MAX_LINES),
        ('nii6548_idx', C.c_int), ('nii6583_idx', C.c_int),
        ('nz', C.c_int), ('axis', C.POINTER(C.c_double)),
        ('fit_lo', C.c_int), ('fit_hi', C.c_int), ('noise_lo', C.c_int), ('noise_hi', C.c_int),
        ('cont_lo', C.c_int), ('cont_hi', C.c_int),
        ('step_nm', C.c_double), ('n_steps', C.c_double), ('zpd_index', C.c_double),
        ('transmission', C.POINTER(C.c_double)),
        ('uncertainty', C.c_int), ('scale_mode', C.c_int), ('max_iter', C.c_int), ('ftol', C.c_double),
        ('prior_groups', C.c_int), ('freeze', C.c_int), ('no_prior_refine', C.c_int),
    ]


EXPORTS = ['lucifit_row_floats', 'lucifit_workspace_bytes', 'lucifit_fit_batch', 'lucifit_fit_stages', 'lucifit_deep_image',
           'lucifit_snr_map', 'lucifit_bin_cube', 'lucifit_sim_cube', 'lucifit_peak_probe', 'lucifit_estimate_prior',
           'lucifit_sanitize_cube', 'lucifit_row_median', 'lucifit_pack_maps', 'lucifit_segment_sum', 'lucifit_peer_export', 'lucifit_peer_open', 'lucifit_peer_close', 'lucifit_last_error',
           'lucifit_version']

_lib = None


def declare(lib):
    """Attach argtypes/restypes of every entry point of include/lucifit.h to a loaded CDLL."""
    vp, i64, sz = C.c_void_p, C.c_int64, C.c_size_t
    cfgp = C.POINTER(lucifit_config)
    lib.lucifit_row_floats.argtypes = [C.c_int]
    lib.lucifit_row_floats.restype = C.c_int
    lib.lucifit_workspace_bytes.argtypes = [cfgp, i64, C.POINTER(sz)]
    lib.lucifit_workspace_bytes.restype = C.c_int
    lib.lucifit_fit_batch.argtypes = [cfgp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.lucifit_fit_batch.restype = C.c_int
    lib.lucifit_fit_stages.argtypes = [cfgp, C.c_int, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.lucifit_fit_stages.restype = C.c_int
    lib.lucifit_deep_image.argtypes = [vp, i64, C.c_int, i64, vp, vp]
    lib.lucifit_deep_image.restype = C.c_int
    lib.lucifit_snr_map.argtypes = [vp, i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    lib.lucifit_snr_map.restype = C.c_int
    lib.lucifit_bin_cube.argtypes = [vp] + [C.c_int] * 8 + [vp, vp]
    lib.lucifit_bin_cube.restype = C.c_int
    lib.lucifit_sim_cube.argtypes = [cfgp, i64, vp, vp, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp, sz, vp]
    lib.lucifit_sim_cube.restype = C.c_int
    lib.lucifit_peak_probe.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp]
    lib.lucifit_peak_probe.restype = C.c_int
    lib.lucifit_estimate_prior.argtypes = [cfgp, i64, vp, vp, vp, vp, C.c_double, C.c_double, C.c_double, vp, vp, vp]
    lib.lucifit_estimate_prior.restype = C.c_int
    lib.lucifit_sanitize_cube.argtypes = [vp, i64, C.c_float, C.c_float, C.c_float, vp]
    lib.lucifit_sanitize_cube.restype = C.c_int
    lib.lucifit_row_median.argtypes = [vp, i64, C.c_int, C.c_int, C.c_int, vp, vp]
    lib.lucifit_row_median.restype = C.c_int
    lib.lucifit_pack_maps.argtypes = [vp, i64, C.c_int, vp, vp, vp]
    lib.lucifit_pack_maps.restype = C.c_int
    lib.lucifit_segment_sum.argtypes = [vp, C.c_int, vp, vp, i64, C.c_int, C.c_int, vp, vp]
    lib.lucifit_segment_sum.restype = C.c_int
    lib.lucifit_peer_export.argtypes = [vp, vp, C.POINTER(C.c_uint64)]
    lib.lucifit_peer_export.restype = C.c_int
    lib.lucifit_peer_open.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    lib.lucifit_peer_open.restype = C.c_int
    lib.lucifit_peer_close.argtypes = [vp, C.c_uint64]
    lib.lucifit_peer_close.restype = C.c_int
    lib.lucifit_last_error.argtypes = []
    lib.lucifit_last_error.restype = C.c_char_p
    lib.lucifit_version.argtypes = []
    lib.lucifit_version.restype = C.c_char_p
    return lib


def lib():
    """The loaded library; raises LucifitUnavailable (loudly) if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise LucifitUnavailable(
                'liblucifit.so is not built (%s). Run `python -c "import __graft_entry__ as g; g.build()"` '
                'or `make -C luci_b200/csrc`. There is no CPU fallback.' % LIB_PATH)
        _lib = declare(C.CDLL(LIB_PATH))
    return _lib


def check(rc):
    if rc != 0:
        raise LucifitError('lucifit error %d: %s' % (rc, lib().lucifit_last_error().decode()))


def argmin_abs(axis, value):
    return int(np.argmin(np.abs(np.asarray(axis) - value)))


class FitConfig:
    """Python owner of a ``lucifit_config`` (keeps the numpy arrays the struct points to alive).

    Arguments mirror ``LUCI.LuciFit.Fit.__init__`` (LuciFit.py:49-57).  ``axis`` is the float32 spectral axis of
    the cube (LuciUtility.py:154-157); it is multiplied by ``1 + obj_redshift`` in float32 like the reference
    (LuciFit.py:113: float32 array times python float).
    """

    def __init__(self, model_type, lines, vel_rel, sigma_rel, axis, filter='SN3', delta_x=2943, n_steps=842,
                 zpd_index=169, trans_filter=None, uncertainty_bool=False, nii_cons=True, spec_min=None,
                 spec_max=None, obj_redshift=0.0, ml_scale=True, max_iter=0, ftol=0.0, prior_groups=False, freeze=False,
                 refine_prior=True, extra_lines=None):
        if model_type not in MODEL_CODE:
            raise Exception('Please submit a fitting function name in the available list: \n {}'.format(
                ['gaussian', 'sinc', 'sincgauss', 'gauss']))
        line_dict = dict(LINE_DICT)
        if extra_lines:                       # e.g. the sky lines of Luci.skyline_calibration: {name: wavelength in nm}
            line_dict.update(extra_lines)
        if not set(lines).issubset(line_dict):
            raise Exception('Please submit a line name in the available list: \n {}'.format(line_dict.keys()))
        if len(vel_rel) != len(lines):
            raise Exception("The argument vel_rel has %i arguments, but it should have %i arguments" % (
                len(vel_rel), len(lines)))
        if len(sigma_rel) != len(lines):
            raise Exception("The argument sigma_rel has %i arguments, but it should have %i arguments" % (
                len(sigma_rel), len(lines)))
        if len(lines) > MAX_LINES:
            raise Exception('at most %d lines per fit are supported' % MAX_LINES)
        self.model_type = 'gaussian' if model_type == 'gauss' else model_type
        self.lines = list(lines)
        self.L = len(lines)
        corr = 1.0 + obj_redshift
        # the reference multiplies its float32 axis by a python float, which stays float32 (LuciFit.py:113): the same
        # rounding here, then widened -- window edges and snapped centres are argmins over exactly these values
        self.axis = np.ascontiguousarray((np.asarray(axis, dtype=np.float32) * np.float32(corr)).astype(np.float64)
                                         if obj_redshift != 0.0 else np.asarray(axis, dtype=np.float64))
        self.trans = None if trans_filter is None else np.ascontiguousarray(trans_filter, dtype=np.float64)
        if spec_min is None or spec_max is None:
            spec_min, spec_max = fit_window(filter, lines, corr)
        nmin, nmax = noise_window(filter, lines, corr)
        cmin, cmax = continuum_window(filter)
        c = lucifit_config()
        c.model = MODEL_CODE[model_type]
        c.n_lines = self.L
        for k, line in enumerate(lines):
            c.line_nm[k] = line_dict[line]
            c.vel_group[k] = int(vel_rel[k])
            c.sigma_group[k] = int(sigma_rel[k])
        c.nii6548_idx = c.nii6583_idx = -1
        if 'NII6548' in lines and 'NII6583' in lines and nii_cons is True:          # LuciFit.py:656
            c.nii6548_idx = self.lines.index('NII6548')                               # first occurrence (:580-581)
            c.nii6583_idx = self.lines.index('NII6583')
        c.nz = len(self.axis)
        c.axis = self.axis.ctypes.data_as(C.POINTER(C.c_double))
        c.fit_lo, c.fit_hi = argmin_abs(self.axis, spec_min), argmin_abs(self.axis, spec_max)
        c.noise_lo, c.noise_hi = argmin_abs(self.axis, nmin), argmin_abs(self.axis, nmax)
        c.cont_lo, c.cont_hi = argmin_abs(self.axis, cmin), argmin_abs(self.axis, cmax)
        c.step_nm, c.n_steps, c.zpd_index = float(delta_x), float(int(n_steps)), float(int(zpd_index))
        c.transmission = (self.trans.ctypes.data_as(C.POINTER(C.c_double)) if self.trans is not None
                          else C.POINTER(C.c_double)())
        c.uncertainty = int(bool(uncertainty_bool))
        c.scale_mode = int(bool(ml_scale))
        c.max_iter = int(max_iter)
        c.ftol = float(ftol)
        c.prior_groups = int(bool(prior_groups))
        c.freeze = int(bool(freeze))
        c.no_prior_refine = int(not refine_prior)
        # dense group order (first appearance), as the library numbers the tie groups
        self.n_vel_groups = len(dict.fromkeys(int(v) for v in vel_rel))
        self.n_sigma_groups = len(dict.fromkeys(int(v) for v in sigma_rel))
        self.c = c

    @property
    def row_floats(self):
        return 7 * self.L + 5

    def byref(self):
        return C.byref(self.c)
