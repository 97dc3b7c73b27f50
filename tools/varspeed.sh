#!/bin/bash
# Developer helper (GPU box): repeated speed runs of variant libraries, interleaved.   tools/varspeed.sh a.so b.so ...
for rep in 1 2 3; do
for v in "$@"; do
  LUCIFIT_LIB=$PWD/luci_b200/csrc/_variants/$v python bench.py --nx 512 --no-cpu-baseline --no-e2e --no-reductions --steps 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v', round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],2), {k:round(v,2) for k,v in r['stage_ms'].items()}, 'E', round(r['mean_evals_E'],3), d['clocks']['sm_mhz'])"
done
done
