"""Developer tool: evaluation counts and parity of the LM stage (CPU build of the device source, tests/hostemu) on the
golden fixtures, for experiments with the optimiser.   python tools/lm_stats.py [-DFLAG ...] [--small] [--big]"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from luci_b200._capi import lucifit_config                                  # noqa: E402
from tests.common import (FIT_CASES, config_from_golden, golden_priors, load_golden, load_tight, optimum_report,    # noqa: E402
                          parity_report, run_hostemu)


def build(flags):
    so = os.path.join(tempfile.mkdtemp(prefix='lmstats_'), 'libhostemu.so')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', *flags, '-o', so,
                           os.path.join(ROOT, 'tests', 'hostemu', 'hostemu.cpp')])
    lib = C.CDLL(so)
    lib.hostemu_fit_batch.argtypes = [C.POINTER(lucifit_config), C.c_int64] + [C.c_void_p] * 10
    lib.hostemu_fit_batch.restype = C.c_int
    return lib


def main():
    flags = [a for a in sys.argv[1:] if a.startswith('-D')]
    lib = build(flags)
    names = [a for a in sys.argv[1:] if not a.startswith('-')]
    if names:
        sys.argv += ['--none']
    if '--small' in sys.argv or ('--big' not in sys.argv and '--none' not in sys.argv):
        names += [n for n in FIT_CASES if n != 'fit_sincgauss_double'] + ['region_sinc_strip']
    if '--big' in sys.argv or ('--small' not in sys.argv and '--none' not in sys.argv):
        names += ['parity_sincgauss_cfg2', 'parity_sincgauss_heldout']
    for name in names:
        g = load_golden(name)
        cfg = config_from_golden(g, uncertainty=False)
        v0, b0 = golden_priors(g)
        rows, sol, unc, st = run_hostemu(lib, cfg, g['cube'], g['theta'], v0, b0, want_unc=False)
        ev, it = (st >> 8) & 0xff, st & 0xff
        conv = (st >> 16) & 1
        cb = g['model'] != 'sinc'
        big = name.startswith('parity_')
        opt = optimum_report(rows, g, tight=load_tight(name) if big else None, check_broadening=cb)
        rep = parity_report(rows, g, g['model'], check_broadening=cb)
        keep = ~rep['ref_spike']
        print('%-26s n %4d  E %.3f  max it %3d  not conv %d  ok %d/%d  same: df %.2e dv %.4f db %.2e  excess %.1e'
              % (name, len(st), ev.mean(), it.max(), (1 - conv)[keep].sum(), opt['ok'][keep].sum(), keep.sum(),
                 np.nanmax(opt['df'][opt['same']]), opt['dv'][opt['same']].max(), opt['db'][opt['same']].max(),
                 opt['excess'][opt['same']].max()), flush=True)
        if '--fails' in sys.argv:
            for i in np.flatnonzero(~opt['ok'] & keep)[:12]:
                print('    [%d] it %d ev %d flags %x  dv %.4f db %.2e df %.2e excess %.2e' % (i, it[i], ev[i], st[i] >> 16, opt['dv'][i], opt['db'][i], opt['df'][i], opt['excess'][i]))


if __name__ == '__main__':
    main()
