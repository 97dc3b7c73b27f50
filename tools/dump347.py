import sys
import numpy as np, torch
sys.path.insert(0, '.')
from luci_b200 import engine
from luci_b200.sim import SN3_LINES, SyntheticCube
dev = torch.device('cuda', 0)
sc = SyntheticCube(37, 29, model='sincgauss', resolution=5000, seed=11)
cube = sc.synthesize(dev).view(37 * 29, -1)
v0, b0 = sc.priors()
v0f, b0f = v0.T.reshape(-1).copy(), b0.T.reshape(-1).copy()
np.savez('gpurun_out/spx347.npz', cube=cube[340:356].cpu().numpy(), v0=v0f[340:356], b0=b0f[340:356], axis32=sc.axis32, theta=sc.theta,
         delta_x=sc.delta_x, n_steps=sc.n_steps, snr=sc.snr.reshape(-1)[340:356])
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
theta = torch.full((cube.shape[0],), sc.theta, dtype=torch.float32, device=dev)
cfg = engine.FitConfig('sincgauss', SN3_LINES, [1] * 5, [1] * 5, sc.axis32, 'SN3', sc.delta_x, sc.n_steps, 0)
for rep in range(4):
    r = engine.fit_batch(cfg, cube, theta, t(v0f), t(b0f))
    print(rep, 'status %x' % int(r['status'][347]), r['rows'][347, [0, 4, 15, 25, 35]].tolist())
