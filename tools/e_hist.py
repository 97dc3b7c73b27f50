"""Developer tool (GPU box): histogram of the LM evaluations per spaxel on the benchmark cube, by SNR decade third."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from luci_b200 import engine                                 # noqa: E402
from luci_b200._capi import FitConfig                        # noqa: E402
from luci_b200.sim import SN3_LINES, SyntheticCube           # noqa: E402

dev = torch.device('cuda', 0)
sc = SyntheticCube(128, 2064, n_steps=841)
cube = sc.synthesize(dev).view(-1, 841)
v0m, b0m = sc.priors()
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
v0, b0 = t(v0m.T.reshape(-1).copy()), t(b0m.T.reshape(-1).copy())
theta = torch.full((cube.shape[0],), float(sc.theta), dtype=torch.float32, device=dev)
cfg = FitConfig('sincgauss', SN3_LINES, [1] * 5, [1] * 5, sc.axis32, 'SN3', sc.delta_x, sc.n_steps, 0, nii_cons=True)
st = engine.fit_batch(cfg, cube, theta, v0, b0)['status'].cpu().numpy()
ev, it = (st >> 8) & 0xff, st & 0xff
snr = sc.snr.reshape(-1)
print('mean E %.3f  mean it %.3f  singular %.4f  bound %.4f  converged %.4f' % (ev.mean(), it.mean(), ((st >> 19) & 1).mean(), ((st >> 18) & 1).mean(), ((st >> 16) & 1).mean()))
print('evals - iters histogram', np.bincount(np.clip(ev.astype(int) - it.astype(int), 0, 9)) / len(ev))
print('E histogram', np.bincount(ev)[:16] / len(ev))
for lo, hi in ((10, 21.5), (21.5, 46.4), (46.4, 100)):
    m = (snr >= lo) & (snr < hi)
    print('SNR %5.1f-%5.1f  E %.3f  hist %s' % (lo, hi, ev[m].mean(), np.round(np.bincount(ev[m], minlength=12)[:12] / m.sum(), 3)))
