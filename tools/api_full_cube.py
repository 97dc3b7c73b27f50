#!/usr/bin/env python
"""Wall time of the drop-in call itself -- ``Luci.fit_entire_cube`` on the full 2048 x 2064 x 841 synthetic SN3 cube
(BASELINE config 2) -- split into the device fit, the host glue around it and the FITS output the reference also writes.

    python tools/api_full_cube.py [--nx 2048] [--ny 2064]
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from luci_b200 import engine                         # noqa: E402
from luci_b200 import luci as luci_mod               # noqa: E402
from luci_b200.luci import Luci                      # noqa: E402
from luci_b200.sim import SN3_LINES, SyntheticCube   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nx', type=int, default=2048)
    ap.add_argument('--ny', type=int, default=2064)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    sc = SyntheticCube(a.nx, a.ny, model='sincgauss', n_steps=841)
    cube = sc.synthesize(dev)
    out = tempfile.mkdtemp(prefix='luci_api_')
    luci = Luci.from_arrays(cube, sc.hdr_dict, out, 'FULL', spectrum_axis=sc.axis32, device=dev)
    v0, b0 = sc.priors()
    t_fit = {'s': 0.0}
    real_fit = engine.fit_batch

    def timed_fit(*args, **kw):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = real_fit(*args, **kw)
        torch.cuda.synchronize(); t_fit['s'] += time.perf_counter() - t0
        return r

    t_io = {'s': 0.0}
    real_save = Luci._save_maps

    def timed_save(self, *args, **kw):
        t0 = time.perf_counter()
        r = real_save(self, *args, **kw)
        t_io['s'] += time.perf_counter() - t0
        return r

    engine.fit_batch = timed_fit
    Luci._save_maps = timed_save
    luci.create_deep_image()
    for label, kw in (('given priors (first call: allocates the pinned staging)', dict(prior_values=(v0, b0))),
                      ('given priors', dict(prior_values=(v0, b0))), ('device-measured priors', dict())):
        t_fit['s'] = t_io['s'] = 0.0
        torch.cuda.synchronize(); t0 = time.perf_counter()
        vel, broad, flux, ampl = luci.fit_entire_cube(SN3_LINES, 'sincgauss', [1] * 5, [1] * 5, **kw)
        torch.cuda.synchronize(); total = time.perf_counter() - t0
        n = a.nx * a.ny
        print('%s: total %.2f s (%.2f M spaxels/s) = device fit %.2f s + FITS output %.2f s + host glue %.2f s; converged %d / %d'
              % (label, total, n / total / 1e6, t_fit['s'], t_io['s'], total - t_fit['s'] - t_io['s'],
                 luci.last_fit_summary['n_converged'], n))
    assert vel.shape == (a.ny, a.nx, 5)


if __name__ == '__main__':
    main()
