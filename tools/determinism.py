"""Debug helper (GPU box): the same batch fitted several times (and in two halves) must give bit-identical rows."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from luci_b200 import engine                                 # noqa: E402
from luci_b200._capi import FitConfig                        # noqa: E402
from luci_b200.sim import SN3_LINES, SyntheticCube           # noqa: E402

dev = torch.device('cuda', 0)
model = sys.argv[1] if len(sys.argv) > 1 else 'sincgauss'
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 32
sc = SyntheticCube(nx, 2064, n_steps=841, model=model)
cube = sc.synthesize(dev).view(-1, 841)
v0m, b0m = sc.priors()
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
v0, b0 = t(v0m.T.reshape(-1).copy()), t(b0m.T.reshape(-1).copy())
theta = torch.full((cube.shape[0],), float(sc.theta), dtype=torch.float32, device=dev)
cfg = FitConfig(model, SN3_LINES, [1] * 5, [1] * 5, sc.axis32, 'SN3', sc.delta_x, sc.n_steps, 0, nii_cons=True)
ref = engine.fit_batch(cfg, cube, theta, v0, b0)
r0, s0 = ref['rows'].clone(), ref['status'].clone()
for rep in range(3):
    r = engine.fit_batch(cfg, cube, theta, v0, b0)
    same = (r['rows'] == r0) | (torch.isnan(r['rows']) & torch.isnan(r0))
    bad = (~same).any(dim=1)
    print('repeat', rep, 'rows differing', int(bad.sum()), 'status differing', int((r['status'] != s0).sum()))
    if bad.any():
        i = int(torch.nonzero(bad)[0])
        print('   first', i, 'status %x vs %x' % (int(r['status'][i]), int(s0[i])), r['rows'][i, 15:18].tolist(), r0[i, 15:18].tolist())
h = cube.shape[0] // 2
ra = engine.fit_batch(cfg, cube[:h], theta[:h], v0[:h], b0[:h])['rows']
rb = engine.fit_batch(cfg, cube[h:], theta[h:], v0[h:], b0[h:])['rows']
both = torch.cat([ra, rb])
same = (both == r0) | (torch.isnan(both) & torch.isnan(r0))
print('two halves: rows differing', int((~same).any(dim=1).sum()))
vt = torch.from_numpy(sc.vel_fit_truth.reshape(-1)).to(dev)
print('|v - truth| > 60 km/s:', int(((r0[:, 15] - vt).abs() > 60).sum()), 'of', r0.shape[0], ' broadening at 0:', int((r0[:, 25] <= 0).sum()), ' max evals', int(((s0 >> 8) & 0xff).max()), ' singular', int(((s0 >> 19) & 1).sum()))
