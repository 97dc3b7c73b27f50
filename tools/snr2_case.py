"""Developer tool (GPU box): SNR method 2 on a 128 x 2064 x 841 synthetic slab (timing / ncu target)."""
import sys

import torch

sys.path.insert(0, '.')
from luci_b200 import engine                                 # noqa: E402
from luci_b200.sim import SyntheticCube                      # noqa: E402

dev = torch.device('cuda', 0)
sc = SyntheticCube(128, 2064, n_steps=841)
cube = sc.synthesize(dev).view(-1, 841)
for _ in range(2):
    engine.snr_map(cube, 2, 500, 560, 60, 110)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    engine.snr_map(cube, 2, 500, 560, 60, 110)
e1.record(); e1.synchronize()
print('snr2 ms per 264192 spaxels', e0.elapsed_time(e1) / 5)
