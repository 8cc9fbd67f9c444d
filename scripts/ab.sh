#!/bin/bash
# A/B timing of library variants on ONE box: scripts/ab.sh "<variant> <variant> ..." [repeats]
# ("base" = the production library).  Prints ms/step and per-kernel ms, alternating variants.
variants=${1:-base}; reps=${2:-2}
for r in $(seq $reps); do for v in $variants; do
  if [ "$v" = base ]; then export JXF_LIB_VARIANT=; else export JXF_LIB_VARIANT=$v; fi
  python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernel_ms']
print('$v', round(d['ms_per_step'],2), 'ms', {a:b for a,b in k.items() if b}, d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))"
done; done
