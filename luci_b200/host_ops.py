"""Small host-side helpers of the path that are not worth a kernel (O(nx*ny) booleans, one spectrum)."""
import numpy as np
import torch

from . import engine
from ._capi import argmin_abs


def bin_mask(mask, binning, x_min, x_max, y_min, y_max):
    """``LUCI/LuciUtility.py:358-388``: a bin is kept when any of its pixels is set (then divided by b^2, so the
    value is truthy exactly like the reference's)."""
    b = int(binning)
    nxb, nyb = int((x_max - x_min) / b), int((y_max - y_min) / b)
    m = np.asarray(mask)[x_min:x_min + nxb * b, y_min:y_min + nyb * b].astype(bool)
    out = m.reshape(nxb, b, nyb, b).any(axis=(1, 3)).astype(np.float64)
    return out / (b ** 2)


def plot_vector(cfg, fit_sol, theta_deg, device):
    """``Model.plot(axis, fit_sol) + continuum`` (LuciFit.py:747-752): every centre snapped to the nearest axis
    sample.  Evaluated on the device by the synthetic-spectrum kernel (same snapping, same profiles)."""
    L = cfg.L
    axis = cfg.axis
    corr = 1.0 / abs(np.cos(np.deg2rad(theta_deg)))
    amp = np.zeros((1, L), np.float32)
    vel = np.zeros((1, L), np.float32)
    broad = np.zeros((1, L), np.float32)
    for k in range(L):
        A, p, s = fit_sol[3 * k:3 * k + 3]
        lam_cm = 1e7 / cfg.c.line_nm[k]
        psn = axis[argmin_abs(axis, p)]
        amp[0, k] = A
        vel[0, k] = (p / lam_cm - 1.0) * 3e5
        broad[0, k] = abs(s) * 3e5 * corr / psn
    t = lambda a: torch.from_numpy(a).to(device)
    out = engine.sim_cube(cfg, t(amp), t(vel), t(broad), t(np.array([theta_deg], np.float32)),
                          continuum=float(fit_sol[-1]), flux_scale=1.0)
    return out[0].cpu().numpy().astype(np.float64)
