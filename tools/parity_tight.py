"""Developer tool (build container or any CPU box): who is closer to the true optimum of the reference's objective?

For every spaxel of a golden fixture the reference's problem is solved to machine precision (oracle/tight.py, started
from the reference's own solution) and both the reference's SLSQP answer and ours (the device source compiled for the
CPU, tests/hostemu, or rows passed in) are compared with that optimum under the north_star tolerances.

    python tools/parity_tight.py parity_sincgauss_cfg2 [--jobs 8] [--save]

--save stores the tight optimum next to the fixture as tests/golden/<name>_tight.npz (used by the parity tests).
"""
import multiprocessing as mp
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import tight as T                                       # noqa: E402
from tests.common import GOLDEN, load_golden                        # noqa: E402

_G = {}


def _solve(i):
    g = _G['g']
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        prob = T.problem_from_golden(g, i)
        scale = float(g['ref_scale'][i])
        ref = np.array(g['ref_fit_sol'][i], float).copy()
        ref[0:-1:3] /= scale
        ref[-1] /= scale
        full, obj, res = prob.solve(ref)
        d = T.derived(prob, full, scale)
        return (i, full, obj, prob.objective_full(ref), d['velocities'], d['sigmas'], d['fluxes'], d['amplitudes'],
                float(np.linalg.norm(res.grad, np.inf)))


def tight_solutions(name, jobs=8):
    g = load_golden(name)
    _G['g'] = g
    n = g['cube'].shape[0]
    with mp.get_context('fork').Pool(jobs) as pool:
        out = pool.map(_solve, range(n), chunksize=4)
    out.sort(key=lambda t: t[0])
    return g, dict(tight_sol=np.stack([o[1] for o in out]), tight_obj=np.array([o[2] for o in out]),
                   ref_obj=np.array([o[3] for o in out]), tight_velocities=np.stack([o[4] for o in out]),
                   tight_sigmas=np.stack([o[5] for o in out]), tight_fluxes=np.stack([o[6] for o in out]),
                   tight_amplitudes=np.stack([o[7] for o in out]), tight_grad=np.array([o[8] for o in out]))


def report(tag, vel, sig, flux, t):
    dv = np.abs(vel - t['tight_velocities']).max(axis=1)
    db = (np.abs(sig - t['tight_sigmas']) / np.abs(t['tight_sigmas'])).max(axis=1)
    df = (np.abs(flux - t['tight_fluxes']) / np.abs(t['tight_fluxes'])).max(axis=1)
    n = len(dv)
    print('%-10s vs tight optimum: dv<=0.5: %d/%d (max %.4f)  db<=1e-3: %d/%d (max %.2e)  dflux<=1e-3: %d/%d (max %.2e)'
          % (tag, (dv <= 0.5).sum(), n, dv.max(), (db <= 1e-3).sum(), n, db.max(), (df <= 1e-3).sum(), n, df.max()))
    return dv, db, df


if __name__ == '__main__':
    argv = sys.argv[1:]
    jobs = 8
    if '--jobs' in argv:
        jobs = int(argv[argv.index('--jobs') + 1])
    name = argv[0]
    g, t = tight_solutions(name, jobs)
    print('objective: reference above tight by (relative) median %.2e max %.2e; |grad|_inf at tight max %.2e'
          % (np.median(t['ref_obj'] / t['tight_obj'] - 1), (t['ref_obj'] / t['tight_obj'] - 1).max(), t['tight_grad'].max()))
    report('reference', g['ref_velocities'], g['ref_sigmas'], g['ref_fluxes'], t)
    if '--save' in argv:
        path = os.path.join(GOLDEN, name + '_tight.npz')
        np.savez_compressed(path, **t)
        print('saved', path)
