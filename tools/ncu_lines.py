#!/usr/bin/env python
"""Developer tool: executed warp instructions per source line of one kernel of an Nsight Compute report.
    python tools/ncu_lines.py report.ncu-rep lf_out_kernel [--spaxels N] [--top 60]"""
import argparse
import csv
import io
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report'); ap.add_argument('kernel')
    ap.add_argument('--spaxels', type=float, default=264192.0)
    ap.add_argument('--top', type=int, default=60)
    a = ap.parse_args()
    out = subprocess.run(['ncu', '-i', a.report, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '-k',
                          'regex:' + a.kernel, '-c', '1'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    cur, h2, rows = None, None, []
    for r in csv.reader(io.StringIO(out)):
        if len(r) == 2 and r[0] == 'File Path':
            cur, h2 = r[1], None
            continue
        if len(r) == 2:
            continue
        if r and r[0] == 'Line No':
            h2 = r
            continue
        if h2 is None or len(r) < 8:
            continue
        try:
            rows.append((os.path.basename(cur), int(r[0]), int(r[h2.index('Instructions Executed')]) / a.spaxels,
                         int(r[h2.index('# Samples')])))
        except ValueError:
            continue
    src = {}
    for f in set(r[0] for r in rows):
        p = os.path.join(ROOT, 'luci_b200', 'csrc', f)
        if os.path.exists(p):
            src[f] = open(p).read().split('\n')
    tot, smp = sum(r[2] for r in rows), sum(r[3] for r in rows)
    rows.sort(key=lambda r: -r[2])
    print('total %.0f instructions per spaxel, %d samples' % (tot, smp))
    for r in rows[:a.top]:
        s = src.get(r[0], [])
        print('%-16s %5d %8.1f %5.1f%%  smp %4.1f%%  %s' % (r[0].replace('lucifit_', ''), r[1], r[2], 100 * r[2] / tot, 100.0 * r[3] / max(smp, 1),
                                                  (s[r[1] - 1] if r[1] - 1 < len(s) else '').strip()[:105]))


if __name__ == '__main__':
    main()
