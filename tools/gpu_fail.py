"""Debug helper (GPU box): spaxels of a golden fixture whose velocity / flux leave the parity bar, with their status words."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from tests.common import config_from_golden, golden_priors, load_golden, parity_report  # noqa: E402


def main(name):
    from luci_b200 import engine
    g = load_golden(name)
    cfg = config_from_golden(g, uncertainty=False)
    v0, b0 = golden_priors(g)
    dev = torch.device('cuda', 0)
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    res = engine.fit_batch(cfg, t(g['cube'].astype(np.float32)), t(g['theta']), t(v0), t(b0), want_sol=True)
    torch.cuda.synchronize()
    st = res['status'].cpu().numpy()
    rows = res['rows'].cpu().numpy()
    rep = parity_report(rows, g, g['model'], check_broadening=g['model'] != 'sinc')
    bad = np.flatnonzero((rep['dv'] > 0.05) | (rep['df'] > 1e-3) | (rep['dchi'] > 1e-4))
    print(name, 'n', len(st), 'bad', len(bad), 'mean E', ((st >> 8) & 0xff).mean())
    for i in bad[:20]:
        print('  [%d] it %d ev %d flags %x dv %.4f db %.2e df %.2e dchi %.2e  v0 %.2f b0 %.2f  ref v %.2f b %.2f  ours v %.2f b %.2f'
              % (i, st[i] & 0xff, (st[i] >> 8) & 0xff, st[i] >> 16, rep['dv'][i], rep['db'][i], rep['df'][i], rep['dchi'][i],
                 np.ravel(v0[i])[0], np.ravel(b0[i])[0], g['ref_velocities'][i][0], g['ref_sigmas'][i][0],
                 rep['q']['velocities'][i][0], rep['q']['sigmas'][i][0]))


if __name__ == '__main__':
    for nm in sys.argv[1:]:
        main(nm)
