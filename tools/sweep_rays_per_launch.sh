for r in 65536 131072 262144 524288; do
  python bench.py --no-extras --no-cpu-baseline --rays-per-launch $r 2>/dev/null | tail -1 > gpurun_out/sweep_$r.json
  python -c "
import json,sys
d=json.loads(open('gpurun_out/sweep_$r.json').read())
print($r, d['ms_per_step'], d['value'])"
done
