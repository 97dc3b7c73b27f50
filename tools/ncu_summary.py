#!/usr/bin/env python
"""Summarise an Nsight Compute report (read here, without a GPU) into the text files kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_lm_ncu.txt [--spaxels N]

Writes the launch configuration, pipe utilisation, stall reasons per issued instruction, DRAM traffic and -- when
the report was captured with --import-source on and the kernels were built with -lineinfo -- the executed
instructions per source function (per spaxel when --spaxels is given).
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [r'^gpu__time_duration\.sum$', r'^launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_\w+)$',
        r'^sm__warps_active\.avg\.pct_of_peak_sustained_active$', r'^smsp__issue_active\.avg\.pct_of_peak_sustained_active$',
        r'^sm__inst_executed_pipe_(fma|alu|xu|lsu|fp64)\.avg\.pct_of_peak_sustained_active$',
        r'^smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio$', r'^smsp__inst_executed\.sum$',
        r'^smsp__thread_inst_executed_per_inst_executed\.ratio$', r'^sass__inst_executed_local_(loads|stores)$',
        r'^dram__bytes_(read|write)\.sum$', r'^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$',
        r'^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$']


def ncu(args):
    return subprocess.run(['ncu'] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def functions_of(path):
    out = []
    try:
        for i, l in enumerate(open(path), 1):
            m = re.match(r'^(?:static )?(?:LF_CORE|LF_HD|__device__ __forceinline__)\s+[\w:<>\*& ]+?\s+(\w+)\(', l)
            if m:
                out.append((i, m.group(1)))
            m = re.match(r'^template <int N.*struct (LfReduceScatter)', l)
            if m:
                out.append((i, m.group(1)))
            m = re.match(r'^lf_(\w+_kernel)\(', l)
            if m:
                out.append((i - 2, 'lf_' + m.group(1)))
    except OSError:
        pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('out')
    ap.add_argument('--spaxels', type=float, default=0.0)
    ap.add_argument('--note', default='')
    ap.add_argument('--kernel', default='', help='regex: summarise only one launch of this kernel')
    ap.add_argument('--skip', type=int, default=0, help='launches of --kernel to skip')
    a = ap.parse_args()
    filt = ['-k', 'regex:' + a.kernel, '-s', str(a.skip), '-c', '1'] if a.kernel else []
    raw = list(csv.reader(io.StringIO(ncu(['-i', a.report, '--page', 'raw', '--csv'] + filt))))
    hdr, units = raw[0], raw[1]
    lines = []
    for r in raw[2:]:
        lines.append('kernel: %s   grid %s block %s' % (r[hdr.index('Kernel Name')], r[hdr.index('Grid Size')], r[hdr.index('Block Size')]))
        for i, h in enumerate(hdr):
            if any(re.search(k, h) for k in KEYS):
                lines.append('  %-88s %-16s %s' % (h, units[i], r[i]))
        lines.append('')
    src = ncu(['-i', a.report, '--page', 'source', '--csv', '--print-source', 'cuda,sass'] + filt)
    fmap = {}
    agg = collections.defaultdict(lambda: [0, 0])
    cur, h2, tot, kern = None, None, 0, None
    for r in csv.reader(io.StringIO(src)):
        if len(r) == 2 and r[0] == 'File Path':
            cur = r[1]; h2 = None
            if cur not in fmap:
                fmap[cur] = functions_of(cur if os.path.exists(cur) else os.path.join(ROOT, 'luci_b200', 'csrc', os.path.basename(cur)))
            continue
        if len(r) == 2:
            if r[0] == 'Function Name' and kern is None:
                kern = r[1]
            continue
        if r and r[0] == 'Line No':
            h2 = r; continue
        if h2 is None or len(r) < 8:
            continue
        try:
            ln, ie, smp = int(r[0]), int(r[h2.index('Instructions Executed')]), int(r[h2.index('# Samples')])
        except ValueError:
            continue
        name = os.path.basename(cur)
        best = None
        for l0, fn in fmap.get(cur, []):
            if l0 <= ln + 1:
                best = fn
        if best:
            name = '%s:%s' % (os.path.basename(cur).replace('lucifit_', ''), best)
        agg[name][0] += ie; agg[name][1] += smp; tot += ie
    if tot:
        ts = max(1, sum(v[1] for v in agg.values()))
        lines.append('executed warp instructions by source function (first kernel of the report%s):' % (
            ', per spaxel, %d spaxels' % a.spaxels if a.spaxels else ''))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            if v[0] * 1000 < tot:
                continue
            lines.append('  %-44s %12.0f  %5.1f %%   stall samples %5.1f %%' % (k, v[0] / a.spaxels if a.spaxels else v[0],
                                                                               100.0 * v[0] / tot, 100.0 * v[1] / ts))
        lines.append('  %-44s %12.0f' % ('total', tot / a.spaxels if a.spaxels else tot))
    with open(a.out, 'w') as f:
        f.write('source report: %s\n' % a.report)
        if a.note:
            f.write(a.note + '\n')
        f.write('\n'.join(lines) + '\n')
    print(open(a.out).read())


if __name__ == '__main__':
    main()
