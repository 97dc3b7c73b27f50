import sys
import numpy as np, torch
sys.path.insert(0, '.')
from luci_b200 import engine
from luci_b200.sim import SN3_LINES, SyntheticCube
dev = torch.device('cuda', 0)
sc = SyntheticCube(37, 29, model='sincgauss', resolution=5000, seed=11)
cube = sc.synthesize(dev).view(37 * 29, -1)
n, nz = cube.shape
v0, b0 = sc.priors()
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
v0f, b0f = t(v0.T.reshape(-1).copy()), t(b0.T.reshape(-1).copy())
theta = torch.full((n,), sc.theta, dtype=torch.float32, device=dev)
cfg = engine.FitConfig('sincgauss', SN3_LINES, [1] * 5, [1] * 5, sc.axis32, 'SN3', sc.delta_x, sc.n_steps, 0)
ref = engine.fit_batch(cfg, cube, theta, v0f, b0f)
r = ref['rows']
print('nan rows', int(torch.isnan(r[:, :5]).any(dim=1).sum()), 'row347', r[347, :5].tolist(), 'status %x' % int(ref['status'][347]))
neq = 0
for rep in range(5):
    ref2 = engine.fit_batch(cfg, cube, theta, v0f, b0f)
    neq += int(((ref['rows'] != ref2['rows']) & ~(torch.isnan(ref['rows']) & torch.isnan(ref2['rows']))).any(dim=1).sum())
print('differing rows in 5 repeats', neq)
