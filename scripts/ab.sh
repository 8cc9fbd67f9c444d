#!/bin/bash
# A/B timing of library variants on ONE box: scripts/ab.sh "<spec> <spec> ..." [repeats]
# spec = variant[,ENV=VAL[,ENV=VAL...]]   ("base" = the production library).
# Prints ms/step and per-kernel ms, alternating the specs.
specs=${1:-base}; reps=${2:-2}
for r in $(seq $reps); do for spec in $specs; do
  IFS=',' read -ra parts <<< "$spec"
  v=${parts[0]}
  envs=()
  for e in "${parts[@]:1}"; do envs+=("$e"); done
  if [ "$v" = base ]; then lv=; else lv=$v; fi
  env JXF_LIB_VARIANT=$lv "${envs[@]}" python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernel_ms']
print('$spec', round(d['ms_per_step'],2), 'ms', {a:b for a,b in k.items() if b}, d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))"
done; done
