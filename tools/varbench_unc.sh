for v in "$@"; do
  echo "== $v"
  LUCIFIT_LIB=$PWD/_variants/$v python bench.py --nx 512 --uncertainty 1 --no-cpu-baseline --no-e2e --no-reductions --steps 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']/1e6,3), 'M/s', {k:round(v,2) for k,v in r['stage_ms'].items()})"
done
