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
neq = 0
for rep in range(10):
    ref2 = engine.fit_batch(cfg, cube, theta, v0f, b0f)
    d = (ref['rows'] != ref2['rows']).any(dim=1)
    neq += int(d.sum())
    if d.any() and rep < 3:
        i = int(torch.nonzero(d)[0]); print('  differs at', torch.nonzero(d).flatten().tolist(), 'status %x %x' % (int(ref['status'][i]), int(ref2['status'][i])))
print('resident repeats: differing rows in 10 repeats:', neq)
for chunk in ():
    rows = []
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        rows.append(engine.fit_batch(cfg, cube[lo:hi], theta[lo:hi], v0f[lo:hi], b0f[lo:hi])['rows'])
    rows = torch.cat(rows)
    bad = (rows != ref['rows']).any(dim=1)
    print('chunk', chunk, 'rows differing', int(bad.sum()), torch.nonzero(bad).flatten()[:8].tolist())
    if bad.any():
        i = int(torch.nonzero(bad)[0]); print('   ', rows[i, 15:18].tolist(), ref['rows'][i, 15:18].tolist(), rows[i, 35].item(), ref['rows'][i, 35].item())
