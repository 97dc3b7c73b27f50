#!/bin/bash
# Developer helper (GPU box): speed + parity of single-instance variant libraries (csrc/dev_variant.sh, model 2, L 5).
#   tools/varcheck.sh a.so b.so ...
for v in "$@"; do
  echo "== $v"
  LUCIFIT_LIB=$PWD/luci_b200/csrc/_variants/$v python bench.py --nx 512 --no-cpu-baseline --no-e2e --no-reductions --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']/1e6,3), 'M/s', {k:round(v,2) for k,v in r['stage_ms'].items()}, 'E', round(r['mean_evals_E'],3))"
  LUCIFIT_LIB=$PWD/luci_b200/csrc/_variants/$v python -m pytest tests/test_gpu_fit.py -q -x -k "parity_sincgauss or fit_sincgauss_sn3 or fit_sincgauss_trans or fit_sincgauss_nan or hostemu_build" 2>&1 | tail -3
done
