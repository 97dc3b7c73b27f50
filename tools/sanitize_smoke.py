#!/usr/bin/env python
"""Small run of every kernel (all three models, ties, uncertainties, index lists, reductions, binning, streaming) meant to
be executed under `compute-sanitizer --tool memcheck` on the GPU box:

    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from luci_b200 import engine                       # noqa: E402
from luci_b200.sim import SN3_LINES, SyntheticCube  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for model in ('gaussian', 'sinc', 'sincgauss'):
        sc = SyntheticCube(7, 5, model=model, resolution=5000, seed=2)
        cube = sc.synthesize(dev).view(35, -1)
        v0, b0 = sc.priors()
        v0f, b0f = t(v0.T.reshape(-1).copy()), t(b0.T.reshape(-1).copy())
        theta = torch.full((35,), sc.theta, dtype=torch.float32, device=dev)
        for rel, unc in (([1] * 5, True), ([1, 2, 2, 3, 3], False)):
            cfg = engine.FitConfig(model, SN3_LINES, rel, rel, sc.axis32, 'SN3', sc.delta_x, sc.n_steps, 0,
                                   uncertainty_bool=unc)
            r = engine.fit_batch(cfg, cube, theta, v0f, b0f, want_sol=True, want_unc=True)
            idx = torch.arange(34, -1, -2, device=dev)
            engine.fit_batch(cfg, cube, theta[idx], v0f[idx], b0f[idx], index=idx, bkg=cube[0] * 0.01)
            engine.fit_batch(cfg, cube, theta, v0f, b0f, time_stages=True)
            pv, pb = engine.estimate_prior(cfg, cube, theta[idx], index=idx, bkg=cube[0] * 0.01)
            engine.fit_batch(cfg, cube, theta[idx], pv, pb, index=idx)
            engine.fit_batch(cfg, cube, theta, None, None)          # no prior: the prep stage refines around 0
            bad = ~torch.isfinite(r['rows'])
            if bad.any():
                print('non-finite rows:', model, rel, unc, bad.nonzero()[:8].tolist(), r['status'][bad.any(dim=1)][:8].tolist())
        engine.deep_image(cube)
        engine.deep_image(cube[1:])                                   # unaligned view: plain-load kernel
        for method in (1, 2):
            engine.snr_map(cube, method, 300, 360, 10, 60, with_deep=True)
        engine.bin_cube(sc.synthesize(dev), 2, 0, 7, 0, 5)
        junk = sc.synthesize(dev).view(-1)
        junk[5] = float('nan'); junk[-1] = float('inf')
        engine.sanitize_cube(junk); engine.sanitize_cube(junk[3:-2])   # aligned and unaligned views
    sc = SyntheticCube(64, 8, model='sincgauss', resolution=5000, seed=4)
    cube = sc.synthesize(dev).view(512, -1)
    engine.deep_image(cube); engine.snr_map(cube, 1, 300, 360, 10, 60)   # several TMA tiles per CTA
    torch.cuda.synchronize()
    print('sanitize smoke ok')


if __name__ == '__main__':
    main()
