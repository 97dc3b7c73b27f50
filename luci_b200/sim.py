"""Synthetic SITELLE cubes in the manner of ``LUCI.LuciSim.Spectrum`` (``LUCI/LuciSim.py:12-203``).

``sim_axis`` restates the interferometer axis construction; ``SyntheticCube`` draws per-spaxel truth maps
(velocity field, broadening, amplitudes, SNR) on the host with a seeded generator and lets the device kernel
``lucifit_sim_cube`` synthesise the spectra directly in HBM (noise: seeded ``torch.randn`` on the device).  The
fitted conventions follow the reference: a simulated velocity ``v`` is recovered as ``-v * 299792/3e5`` and a
simulated broadening ``b`` as ``b cos(theta) 299792/3e5``.
"""
import numpy as np
import torch

from . import engine
from ._capi import FitConfig
from .constants import SIM_FILTERS


def steps_from_resolution(filter_, resolution):
    """``LuciSim.py:90-96``"""
    delta_x, order = SIM_FILTERS[filter_]
    temp = (order + 0.5) / (2 * delta_x)
    return int(np.ceil(1.20671 * resolution / (2 * temp * delta_x)))


def sim_axis(filter_, resolution=None, n_steps=None, theta=11.96):
    """``LuciSim.py:169-175`` -> (axis float64 [n], n_steps, delta_x).  Give ``n_steps`` directly to pin nz."""
    delta_x, order = SIM_FILTERS[filter_]
    n = int(n_steps) if n_steps is not None else steps_from_resolution(filter_, resolution)
    corr = 1 / np.cos(np.deg2rad(theta))
    x_min = order / (2 * delta_x) * corr * 1e7
    x_max = (order + 1) / (2 * delta_x) * corr * 1e7
    step_ = x_max - x_min
    axis = np.array([x_min + j * step_ / n for j in range(n)])
    return axis, n, delta_x


SN3_LINES = ['Halpha', 'NII6548', 'NII6583', 'SII6716', 'SII6731']


class SyntheticCube:
    """Truth maps + device synthesis for an ``nx x ny`` cube (BASELINE.json configs)."""

    def __init__(self, nx, ny, lines=SN3_LINES, model='sincgauss', filter_='SN3', resolution=5000, n_steps=None,
                 seed=12345, vel_range=(-150.0, 150.0), broad_range=(25.0, 60.0), snr_range=(10.0, 100.0),
                 continuum=0.1, flux_scale=1e-17, theta=11.96, x0=0, nx_total=None, second_component=None):
        self.nx, self.ny, self.lines, self.model, self.filter = nx, ny, list(lines), model, filter_
        self.axis64, self.n_steps, self.delta_x = sim_axis(filter_, resolution, n_steps, theta=11.96)
        self.axis32 = self.axis64.astype(np.float32)
        self.nz = len(self.axis32)
        self.continuum, self.flux_scale, self.theta = continuum, flux_scale, theta
        self.hdr_dict = {'FILTER': filter_, 'STEP': self.delta_x, 'STEPNB': self.n_steps, 'ZPDINDEX': 0}
        nx_total = nx_total if nx_total is not None else nx
        L = len(lines)
        # smooth rotation field over the FULL frame so shards of one cube agree with the single-GPU cube
        xs = (np.arange(x0, x0 + nx)[:, None] - nx_total / 2.0) / max(nx_total, 1)
        ys = (np.arange(ny)[None, :] - ny / 2.0) / max(ny, 1)
        vmid, vamp = 0.5 * (vel_range[0] + vel_range[1]), 0.5 * (vel_range[1] - vel_range[0])
        self.vel_sim = (vmid + vamp * np.sin(2.2 * xs + 1.3 * ys) * np.cos(1.1 * ys)).astype(np.float32)
        rng = np.random.default_rng([seed, x0])
        self.broad_sim = rng.uniform(broad_range[0], broad_range[1], (nx, ny)).astype(np.float32)
        self.snr = np.exp(rng.uniform(np.log(snr_range[0]), np.log(snr_range[1]), (nx, ny))).astype(np.float32)
        amp = np.zeros((nx, ny, L), np.float32)
        base = {'Halpha': 1.0, 'NII6583': 0.3, 'NII6548': 0.1, 'SII6716': 0.15, 'SII6731': 0.12, 'Hbeta': 1.0,
                'OIII5007': 0.9, 'OIII4959': 0.3, 'OII3726': 1.0, 'OII3729': 0.8}
        for k, line in enumerate(lines):
            a0 = base.get(line, 0.5)
            amp[:, :, k] = a0 if a0 == 1.0 else a0 * rng.uniform(0.5, 1.5, (nx, ny))
        if 'NII6548' in lines and 'NII6583' in lines:
            amp[:, :, lines.index('NII6548')] = amp[:, :, lines.index('NII6583')] / 3.0
        self.second_component = second_component
        vel_l = np.repeat(self.vel_sim[:, :, None], L, axis=2)
        broad_l = np.repeat(self.broad_sim[:, :, None], L, axis=2)
        if second_component is not None:      # (line index of the extra component, amplitude, delta v [km/s, fitted sign])
            k2, a2, dv = second_component
            amp[:, :, k2] = a2
            vel_l[:, :, k2] = self.vel_sim - dv * 3e5 / 299792.0
        self.amp, self.vel_l, self.broad_l = amp, vel_l.astype(np.float32), broad_l.astype(np.float32)
        self.prior_rng = np.random.default_rng([seed, x0, 1])

    # truth in the FITTED convention
    @property
    def vel_fit_truth(self):
        return (-self.vel_sim * (299792.0 / 3e5)).astype(np.float32)

    @property
    def broad_fit_truth(self):
        return (self.broad_sim * np.cos(np.deg2rad(self.theta)) * (299792.0 / 3e5)).astype(np.float32)

    def priors(self, vel_sigma=15.0, broad_frac=0.15):
        """(vel0, broad0) maps shaped (ny, nx) like the fit outputs: truth + N(0, 15 km/s), truth*(1+N(0,0.15))."""
        v = self.vel_fit_truth + self.prior_rng.normal(0, vel_sigma, (self.nx, self.ny)).astype(np.float32)
        b = self.broad_fit_truth * (1 + self.prior_rng.normal(0, broad_frac, (self.nx, self.ny))).astype(np.float32)
        return np.ascontiguousarray(v.T), np.ascontiguousarray(np.abs(b).T)

    def config_for_sim(self):
        return FitConfig(self.model, self.lines, [1] * len(self.lines), [1] * len(self.lines), self.axis32,
                         filter=self.filter, delta_x=self.delta_x, n_steps=self.n_steps, zpd_index=0, nii_cons=False)

    def synthesize(self, device, seed=999):
        """Returns the cube ``[nx, ny, nz]`` float32 on ``device``.  LuciSim adds ``max(spectrum) N(0, 1/snr)``
        (LuciSim.py:202); the line peak is ~max(amplitudes) so noise_sigma = max_k(amp)/snr."""
        n = self.nx * self.ny
        L = len(self.lines)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        cube = torch.empty((n, self.nz), dtype=torch.float32, device=device)
        cube.normal_(generator=g)
        nsig = (self.amp.max(axis=2) / self.snr).reshape(n).astype(np.float32)
        theta = np.full(n, self.theta, np.float32)
        engine.sim_cube(self.config_for_sim(), t(self.amp.reshape(n, L)), t(self.vel_l.reshape(n, L)),
                        t(self.broad_l.reshape(n, L)), t(theta), noise_sigma=t(nsig), noise=cube,
                        continuum=self.continuum, flux_scale=self.flux_scale, out=cube)
        return cube.view(self.nx, self.ny, self.nz)
