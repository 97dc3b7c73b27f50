"""Torch-side plumbing around the C ABI: device buffers, streams, workspace, output unpacking.

PyTorch is used for device memory and streams only; every kernel is launched through ``liblucifit.so``.
"""
import ctypes as C

import numpy as np
import torch

from . import _capi
from ._capi import FitConfig, check

ROW_FIELDS = ('amplitudes', 'fluxes', 'flux_errors', 'velocities', 'vels_errors', 'sigmas', 'sigmas_errors')
ROW_SCALARS = ('chi2', 'corr', 'axis_step', 'continuum', 'continuum_error')

# counts launches of our kernels (bench.py reports it as `gpu_launches`)
launch_counter = {'n': 0}


class DevPtr:
    """A raw device address with a shape: an output buffer that is not a tensor of this process -- the destination
    rank's rows / status mapped through ``lucifit_peer_open`` (:class:`luci_b200.dist.PeerRows`)."""

    def __init__(self, ptr, shape):
        self.ptr, self.shape = int(ptr), tuple(int(v) for v in shape)

    def data_ptr(self):
        return self.ptr

    def numel(self):
        n = 1
        for v in self.shape:
            n *= v
        return n


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, name, dtype):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError('%s must be a CUDA tensor (there is no CPU path)' % name)
    if t.dtype != dtype:
        raise TypeError('%s must have dtype %s' % (name, dtype))
    if not t.is_contiguous():
        raise ValueError('%s must be contiguous' % name)
    return t


class Workspace:
    """Caller-owned device workspace, grown on demand and reused across calls (one per device)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        nbytes = max(int(nbytes), 256)
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        off = (-self.buf.data_ptr()) % 256
        return self.buf[off:off + nbytes]


_workspaces = {}


def workspace_for(device, nbytes):
    """One workspace per (device, stream): the fit keeps per-spaxel state in it between its stage kernels, so two
    streams fitting at the same time must not share it."""
    key = (str(device), int(torch.cuda.current_stream(device).cuda_stream))
    ws = _workspaces.setdefault(key, Workspace())
    return ws.get(nbytes, device)


SLICE_SPAXELS = 1 << 20            # LF_SLICE_SPAXELS of csrc/lucifit_host.hpp
STAGES = (('prep', 1), ('lm', 2), ('unc', 4), ('out', 8))


def fit_batch(cfg: FitConfig, cube, theta, vel0=None, broad0=None, index=None, bkg=None, want_sol=False,
              want_unc=False, time_stages=False, rows=None, status=None):
    """One ``lucifit_fit_batch`` call.  ``cube`` is ``[n_rows, nz]`` float32 on the GPU; ``index`` (int64, optional)
    selects the rows to fit; ``theta``/``vel0``/``broad0`` are per *fitted* spaxel.  Returns a dict of device
    tensors: ``rows [n, 7L+5]``, ``status [n]`` and optionally ``fit_sol`` / ``fit_unc`` ``[n, 3L+1]``.

    ``rows`` / ``status`` may be caller-owned tensors or :class:`DevPtr` addresses (peer memory of another rank).

    ``time_stages=True`` runs the same pipeline stage by stage through ``lucifit_fit_stages`` (slice by slice) with
    CUDA events around every kernel on the launching stream and adds ``stage_ms`` (dict, summed over the slices;
    this synchronises)."""
    lib = _capi.lib()
    require_cuda(cube, 'cube', torch.float32)
    dev = cube.device
    if cube.dim() != 2 or cube.shape[1] != cfg.c.nz:
        raise ValueError('cube must be [n_rows, nz=%d]' % cfg.c.nz)
    n = int(index.numel()) if index is not None else int(cube.shape[0])
    require_cuda(theta, 'theta', torch.float32)
    if theta.numel() != n:
        raise ValueError('theta must have one entry per fitted spaxel')
    for name, t, ng in (('vel0', vel0, cfg.n_vel_groups), ('broad0', broad0, cfg.n_sigma_groups)):
        if t is not None:
            require_cuda(t, name, torch.float32)
            want = n * (ng if cfg.c.prior_groups else 1)
            if t.numel() != want:
                raise ValueError('%s must have %s per fitted spaxel' % (name, 'one entry per tie group' if cfg.c.prior_groups
                                                                        else 'one entry'))
    if index is not None:
        require_cuda(index, 'index', torch.int64)
    if bkg is not None:
        require_cuda(bkg, 'bkg', torch.float32)
        if bkg.numel() != cfg.c.nz:
            raise ValueError('bkg must have nz entries')
    K, PF = cfg.row_floats, 3 * cfg.L + 1
    if rows is None:
        rows = torch.empty((n, K), dtype=torch.float32, device=dev)
    else:                                                  # caller-owned outputs (streaming pipelines reuse them)
        if not isinstance(rows, DevPtr):
            require_cuda(rows, 'rows', torch.float32)
        if tuple(rows.shape) != (n, K):
            raise ValueError('rows must be [n, 7L+5]')
    if status is None:
        status = torch.empty((n,), dtype=torch.int32, device=dev)
    else:
        if not isinstance(status, DevPtr):
            require_cuda(status, 'status', torch.int32)
        if status.numel() != n:
            raise ValueError('status must have n entries')
    if time_stages and (isinstance(rows, DevPtr) or isinstance(status, DevPtr)):
        raise ValueError('time_stages needs tensor outputs')
    sol = torch.empty((n, PF), dtype=torch.float32, device=dev) if want_sol else None
    unc = torch.empty((n, PF), dtype=torch.float32, device=dev) if want_unc else None
    need = C.c_size_t(0)
    check(lib.lucifit_workspace_bytes(cfg.byref(), n, C.byref(need)))
    ws = workspace_for(dev, need.value)
    n_stage = 4 if cfg.c.uncertainty else 3
    out = {'rows': rows, 'status': status}
    with torch.cuda.device(dev):
        if not time_stages:
            check(lib.lucifit_fit_batch(cfg.byref(), n, _ptr(cube), _ptr(index), _ptr(theta), _ptr(vel0), _ptr(broad0),
                                        _ptr(bkg), _ptr(rows), _ptr(sol), _ptr(unc), _ptr(status), _ptr(ws),
                                        C.c_size_t(ws.numel()), _stream()))
        else:
            marks = []
            if index is None:
                index = torch.arange(n, dtype=torch.int64, device=dev)     # slices address rows through the index
            for lo in range(0, n, SLICE_SPAXELS):
                hi = min(lo + SLICE_SPAXELS, n)
                sl = lambda t: None if t is None else t[lo:hi]
                slp = lambda t, ng: None if t is None else t.reshape(n, -1)[lo:hi]
                for name, bit in STAGES:
                    if bit == 4 and not cfg.c.uncertainty:
                        continue
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    check(lib.lucifit_fit_stages(cfg.byref(), bit, hi - lo, _ptr(cube), _ptr(sl(index)), _ptr(sl(theta)),
                                                 _ptr(slp(vel0, 0)), _ptr(slp(broad0, 0)), _ptr(bkg), _ptr(sl(rows)), _ptr(sl(sol)),
                                                 _ptr(sl(unc)), _ptr(sl(status)), _ptr(ws), C.c_size_t(ws.numel()),
                                                 _stream()))
                    e1.record()
                    marks.append((name, e0, e1))
            torch.cuda.current_stream().synchronize()
            ms = {}
            for name, e0, e1 in marks:
                ms[name] = ms.get(name, 0.0) + e0.elapsed_time(e1)
            out['stage_ms'] = ms
    launch_counter['n'] += n_stage * ((n + SLICE_SPAXELS - 1) // SLICE_SPAXELS)
    if want_sol:
        out['fit_sol'] = sol
    if want_unc:
        out['fit_unc'] = unc
    return out


def keep_best(best, alt, chi2_col):
    """Per spaxel, replace the rows / status of ``best`` by those of ``alt`` where ``alt`` converged with a lower chi2
    (random restarts, LuciFit.py:682-687; component search).  In place on ``best``; returns it."""
    better = ((alt['status'] & (1 << 16)) != 0) & (alt['rows'][:, chi2_col] < best['rows'][:, chi2_col])
    best['rows'] = torch.where(better[:, None], alt['rows'], best['rows'])
    best['status'] = torch.where(better, alt['status'], best['status'])
    for key in ('fit_sol', 'fit_unc'):
        if key in best and key in alt:
            best[key] = torch.where(better[:, None], alt[key], best[key])
    return best


class HostCubeStream:
    """Fits a cube that lives in pinned HOST memory: the spectra are streamed to the GPU in chunks on a copy stream
    while the previous chunk is being fitted, and the result rows are copied back on a third stream, so the whole
    cube never has to be resident (the reference keeps ``cube_final`` in host RAM, LuciBase.py:106-145).

    Buffers are allocated once and reused across calls.  ``run`` returns after the last row has landed in
    ``host_rows``."""

    def __init__(self, cfg: FitConfig, device, chunk=1 << 17, n_buffers=3):
        self.cfg, self.dev, self.chunk, self.nb = cfg, device, int(chunk), int(n_buffers)
        nz, K = cfg.c.nz, cfg.row_floats
        self.copy_s, self.comp_s, self.back_s = (torch.cuda.Stream(device=device) for _ in range(3))
        self.cube_b = [torch.empty((self.chunk, nz), dtype=torch.float32, device=device) for _ in range(self.nb)]
        self.par_b = [torch.empty((3, self.chunk), dtype=torch.float32, device=device) for _ in range(self.nb)]
        self.rows_b = [torch.empty((self.chunk, K), dtype=torch.float32, device=device) for _ in range(self.nb)]
        self.stat_b = [torch.empty((self.chunk,), dtype=torch.int32, device=device) for _ in range(self.nb)]
        self.free_ev = [None] * self.nb

    def run(self, host_cube, host_par, host_rows, host_status=None):
        """``host_cube [n, nz]``, ``host_par [3, n]`` (theta, prior velocity, prior broadening), ``host_rows [n, 7L+5]``
        and optional ``host_status [n]`` are pinned CPU tensors."""
        for name, t in (('host_cube', host_cube), ('host_par', host_par), ('host_rows', host_rows)):
            if t.is_cuda or not t.is_pinned():
                raise TypeError('%s must be a pinned CPU tensor' % name)
        n = host_cube.shape[0]
        for ci, lo in enumerate(range(0, n, self.chunk)):
            hi = min(lo + self.chunk, n)
            m, b = hi - lo, ci % self.nb
            with torch.cuda.stream(self.copy_s):
                if self.free_ev[b] is not None:
                    self.copy_s.wait_event(self.free_ev[b])
                self.cube_b[b][:m].copy_(host_cube[lo:hi], non_blocking=True)
                self.par_b[b][:, :m].copy_(host_par[:, lo:hi], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.copy_s)
            with torch.cuda.stream(self.comp_s):
                self.comp_s.wait_event(ready)
                fit_batch(self.cfg, self.cube_b[b][:m], self.par_b[b][0, :m], self.par_b[b][1, :m], self.par_b[b][2, :m],
                          rows=self.rows_b[b][:m], status=self.stat_b[b][:m])
                done = torch.cuda.Event()
                done.record(self.comp_s)
            with torch.cuda.stream(self.back_s):
                self.back_s.wait_event(done)
                host_rows[lo:hi].copy_(self.rows_b[b][:m], non_blocking=True)
                if host_status is not None:
                    host_status[lo:hi].copy_(self.stat_b[b][:m], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.back_s)
                self.free_ev[b] = ev
        self.back_s.synchronize()
        return host_rows


def unpack_rows(rows, L):
    """Split ``rows [n, 7L+5]`` into the 12 named quantities of ``Luci.fit_calc`` (LuciBase.py:406)."""
    out = {}
    for i, name in enumerate(ROW_FIELDS):
        out[name] = rows[:, i * L:(i + 1) * L]
    for i, name in enumerate(ROW_SCALARS):
        out[name] = rows[:, 7 * L + i]
    return out


PRIOR_V_RANGE = (-500.0, 500.0)    # km/s: the velocity range the reference's CNN was trained on
PRIOR_V_STEP = 12.0                # km/s (1/4 of an R = 5000 SN3 channel; the prep stage of the fit refines it on a 6 km/s grid)


def estimate_prior(cfg: FitConfig, cube, theta, index=None, bkg=None, v_range=PRIOR_V_RANGE, v_step=PRIOR_V_STEP):
    """``lucifit_estimate_prior``: (velocity, broadening) in km/s of every fitted spaxel, measured from the spectra
    (matched filter + line width) -- the stand-in for the reference's CNN prior (LuciFit.py:335-362).  Arguments as
    :func:`fit_batch`; returns two float32 device tensors ``[n]`` usable as its ``vel0`` / ``broad0``."""
    lib = _capi.lib()
    require_cuda(cube, 'cube', torch.float32)
    dev = cube.device
    if cube.dim() != 2 or cube.shape[1] != cfg.c.nz:
        raise ValueError('cube must be [n_rows, nz=%d]' % cfg.c.nz)
    n = int(index.numel()) if index is not None else int(cube.shape[0])
    require_cuda(theta, 'theta', torch.float32)
    if theta.numel() != n:
        raise ValueError('theta must have one entry per fitted spaxel')
    if index is not None:
        require_cuda(index, 'index', torch.int64)
    if bkg is not None:
        require_cuda(bkg, 'bkg', torch.float32)
    vel = torch.empty((n,), dtype=torch.float32, device=dev)
    broad = torch.empty((n,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.lucifit_estimate_prior(cfg.byref(), n, _ptr(cube), _ptr(index), _ptr(theta), _ptr(bkg),
                                         float(v_range[0]), float(v_range[1]), float(v_step), _ptr(vel), _ptr(broad),
                                         _stream()))
    launch_counter['n'] += 1 if n > 0 else 0
    return vel, broad


def deep_image(cube2d, n_zero_tail=0):
    """``cube2d [n_spaxel, nz]`` -> ``deep [n_spaxel]`` (nansum over z); the last ``n_zero_tail`` entries are 0."""
    lib = _capi.lib()
    require_cuda(cube2d, 'cube', torch.float32)
    n, nz = cube2d.shape
    deep = torch.empty((n,), dtype=torch.float32, device=cube2d.device)
    with torch.cuda.device(cube2d.device):
        check(lib.lucifit_deep_image(_ptr(cube2d), n, nz, int(n_zero_tail), _ptr(deep), _stream()))
    launch_counter['n'] += 1 if n > 0 else 0
    return deep


def snr_map(cube2d, method, flux_lo, flux_hi, noise_lo, noise_hi, with_deep=False):
    lib = _capi.lib()
    require_cuda(cube2d, 'cube', torch.float32)
    n, nz = cube2d.shape
    snr = torch.empty((n,), dtype=torch.float32, device=cube2d.device)
    deep = torch.empty((n,), dtype=torch.float32, device=cube2d.device) if with_deep else None
    with torch.cuda.device(cube2d.device):
        check(lib.lucifit_snr_map(_ptr(cube2d), n, nz, int(method), int(flux_lo), int(flux_hi), int(noise_lo),
                                  int(noise_hi), _ptr(snr), _ptr(deep), _stream()))
    launch_counter['n'] += 1 if n > 0 else 0
    return (snr, deep) if with_deep else snr


def bin_cube(cube3d, binning, x_min, x_max, y_min, y_max):
    """``cube3d [nx, ny, nz]`` -> ``[int((x_max-x_min)/b), int((y_max-y_min)/b), nz]`` b x b nansum."""
    lib = _capi.lib()
    require_cuda(cube3d, 'cube', torch.float32)
    nx, ny, nz = cube3d.shape
    b = int(binning)
    nxb, nyb = int((x_max - x_min) / b), int((y_max - y_min) / b)
    out = torch.empty((nxb, nyb, nz), dtype=torch.float32, device=cube3d.device)
    with torch.cuda.device(cube3d.device):
        check(lib.lucifit_bin_cube(_ptr(cube3d), nx, ny, nz, b, int(x_min), int(x_max), int(y_min), int(y_max),
                                   _ptr(out), _stream()))
    launch_counter['n'] += 1 if out.numel() > 0 else 0
    return out


def row_median(cube2d, lo, hi):
    """``lucifit_row_median``: float32 ``[n_rows]`` = ``np.nanmedian(cube2d[r, lo:hi])`` of every row."""
    lib = _capi.lib()
    require_cuda(cube2d, 'cube', torch.float32)
    n, nz = int(cube2d.shape[0]), int(cube2d.shape[1])
    out = torch.empty((n,), dtype=torch.float32, device=cube2d.device)
    with torch.cuda.device(cube2d.device):
        check(lib.lucifit_row_median(_ptr(cube2d), n, nz, int(lo), int(hi), _ptr(out), _stream()))
    launch_counter['n'] += 1 if n > 0 else 0
    return out


def pack_maps(rows, want_be=False):
    """``lucifit_pack_maps``: ``rows [n, K]`` -> field-major ``maps [K, n]`` (float32) and, with ``want_be``, the same
    values byte-reversed (``[K, n]`` int32 holding big-endian floats: the data units of the FITS maps)."""
    lib = _capi.lib()
    require_cuda(rows, 'rows', torch.float32)
    n, K = int(rows.shape[0]), int(rows.shape[1])
    maps = torch.empty((K, n), dtype=torch.float32, device=rows.device)
    be = torch.empty((K, n), dtype=torch.int32, device=rows.device) if want_be else None
    with torch.cuda.device(rows.device):
        check(lib.lucifit_pack_maps(_ptr(rows), n, K, _ptr(maps), _ptr(be), _stream()))
    launch_counter['n'] += 1 if n > 0 else 0
    return (maps, be) if want_be else maps


def segment_sum(cube2d, row_index=None, segment=None, n_segments=1, nansum=False):
    """``lucifit_segment_sum``: float64 ``[n_segments, nz]`` sums of the rows ``row_index`` (int64, default all rows) of
    ``cube2d [n_rows, nz]`` into the segments ``segment`` (int32, default all into segment 0; negative = skip)."""
    lib = _capi.lib()
    require_cuda(cube2d, 'cube', torch.float32)
    dev = cube2d.device
    nz = int(cube2d.shape[1])
    if row_index is not None:
        require_cuda(row_index, 'row_index', torch.int64)
    if segment is not None:
        require_cuda(segment, 'segment', torch.int32)
    n_sel = int(row_index.numel()) if row_index is not None else (int(segment.numel()) if segment is not None else int(cube2d.shape[0]))
    if segment is not None and segment.numel() != n_sel:
        raise ValueError('segment must have one entry per selected row')
    sums = torch.empty((int(n_segments), nz), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib.lucifit_segment_sum(_ptr(cube2d), nz, _ptr(row_index), _ptr(segment), n_sel, int(n_segments),
                                      int(bool(nansum)), _ptr(sums), _stream()))
    launch_counter['n'] += 1 if n_sel > 0 else 0
    return sums


def sanitize_cube(cube, lo=-1e-16, hi=1e-9, fill=1e-22):
    """In-place clamp of ``Luci.read_in_cube`` (LuciBase.py:123-129) on a contiguous float32 device tensor."""
    lib = _capi.lib()
    require_cuda(cube, 'cube', torch.float32)
    with torch.cuda.device(cube.device):
        check(lib.lucifit_sanitize_cube(_ptr(cube), cube.numel(), float(lo), float(hi), float(fill), _stream()))
    launch_counter['n'] += 1 if cube.numel() > 0 else 0
    return cube


def sim_cube(cfg: FitConfig, amp, vel, broad, theta, noise_sigma=None, noise=None, continuum=0.0, flux_scale=1.0,
             out=None):
    """LuciSim-style synthetic spectra written on the device (LuciSim.py:159-203).  ``amp [n, L]``, ``vel``,
    ``broad``, ``theta``, ``noise_sigma`` ``[n]``; ``noise [n, nz]`` unit-variance draws (may alias ``out``)."""
    lib = _capi.lib()
    require_cuda(amp, 'amp', torch.float32)
    n = amp.shape[0]
    dev = amp.device
    nz = cfg.c.nz
    if out is None:
        out = torch.empty((n, nz), dtype=torch.float32, device=dev)
    require_cuda(out, 'out', torch.float32)
    for name, t in (('vel', vel), ('broad', broad), ('theta', theta)):
        require_cuda(t, name, torch.float32)
    ws = workspace_for(dev, 8 * nz + 256)
    with torch.cuda.device(dev):
        check(lib.lucifit_sim_cube(cfg.byref(), n, _ptr(amp), _ptr(vel), _ptr(broad), _ptr(theta), _ptr(noise_sigma),
                                   _ptr(noise), C.c_float(continuum), C.c_float(flux_scale), _ptr(out), _ptr(ws),
                                   C.c_size_t(ws.numel()), _stream()))
    launch_counter['n'] += 1 if n > 0 else 0
    return out


def status_summary(status):
    """Summary of the per-spaxel status words: mean evaluations E, iterations, flag counts.  Reduced on the device the
    words live on (one small copy to the host)."""
    flags = _capi_flags()
    st = status.detach().to(torch.int64).reshape(-1)
    n = int(st.numel())
    if n == 0:
        out = {'n': 0, 'n_fitted': 0, 'mean_evals': 0.0, 'mean_iters': 0.0, 'max_evals': 0}
        out.update({'n_' + name.lower(): 0 for name in flags})
        return out
    bad = flags['NAN_INPUT'] | flags['BAD_NORM'] | flags['MASKED']
    fitted = (st & bad) == 0
    ev, it = (st >> 8) & 0xff, st & 0xff
    parts = [fitted.sum(), (ev * fitted).sum(), (it * fitted).sum(), ev.max()]
    parts += [((st & bit) != 0).sum() for bit in flags.values()]
    vals = torch.stack([p.to(torch.int64) for p in parts]).cpu().tolist()
    nf = vals[0]
    out = {'n': n, 'n_fitted': int(nf), 'mean_evals': vals[1] / nf if nf else 0.0,
           'mean_iters': vals[2] / nf if nf else 0.0, 'max_evals': int(vals[3])}
    for k, name in enumerate(flags):
        out['n_' + name.lower()] = int(vals[4 + k])
    return out


def _capi_flags():
    return {'CONVERGED': 1 << 16, 'MAXITER': 1 << 17, 'BOUND': 1 << 18, 'SINGULAR': 1 << 19, 'NAN_INPUT': 1 << 20,
            'MASKED': 1 << 21, 'BAD_NORM': 1 << 22, 'NAN_SAMPLES': 1 << 24}


def peak_probe(device, mode, iters=4096, blocks_per_sm=8):
    """Measured FP32-FMA (mode 0, TFLOP/s) or MUFU.EX2 (mode 1, Tops/s) throughput of this GPU."""
    lib = _capi.lib()
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    blocks = sms * blocks_per_sm
    out = torch.empty(blocks * 256, dtype=torch.float32, device=device)
    best = 0.0
    with torch.cuda.device(device):
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(lib.lucifit_peak_probe(mode, iters, blocks, _ptr(out), _stream()))
            e1.record()
            e1.synchronize()
            ops = blocks * 256 * iters * 8 * (2 if mode == 0 else 1)
            if rep > 0:
                best = max(best, ops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best
