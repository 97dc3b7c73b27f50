"""Debug helper: one golden fixture through the C ABI on the GPU; prints status words and the parity report."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from tests.common import config_from_golden, golden_priors, load_golden, optimum_report, parity_report  # noqa: E402


def main(name):
    from luci_b200 import engine
    g = load_golden(name)
    cfg = config_from_golden(g)
    v0, b0 = golden_priors(g)
    dev = torch.device('cuda', 0)
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    res = engine.fit_batch(cfg, t(g['cube'].astype(np.float32)), t(g['theta']), t(v0), t(b0), want_sol=True, want_unc=True)
    torch.cuda.synchronize()
    st = res['status'].cpu().numpy()
    rows = res['rows'].cpu().numpy()
    rep = parity_report(rows, g, g['model'], check_broadening=g['model'] != 'sinc')
    opt = optimum_report(rows, g, check_broadening=g['model'] != 'sinc')
    np.set_printoptions(linewidth=200, precision=4)
    print('iters', st & 0xff, 'evals', (st >> 8) & 0xff, 'flags', [hex(s >> 16) for s in st])
    print('spike', rep['ref_spike'].astype(int))
    print('same', opt['same'].astype(int), 'lower', opt['lower'].astype(int))
    print('dv', opt['dv'], 'df', opt['df'], 'excess', opt['excess'])


if __name__ == '__main__':
    for nm in sys.argv[1:]:
        print('==', nm)
        main(nm)
