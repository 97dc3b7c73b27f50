#!/bin/bash
# Developer helper (GPU box): speed of variant libraries with uncertainty_bool=True, interleaved.
for rep in 1 2; do
for v in "$@"; do
  LUCIFIT_LIB=$PWD/luci_b200/csrc/_variants/$v python bench.py --nx 256 --uncertainty 1 --no-cpu-baseline --no-e2e --no-reductions --no-api --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v', round(d['value']/1e6,3), 'M/s', {k:round(v,2) for k,v in r['stage_ms'].items()}, 'E', round(r['mean_evals_E'],3))"
done
done
