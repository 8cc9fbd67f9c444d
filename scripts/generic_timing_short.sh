#!/bin/bash
# subset of scripts/generic_timing.sh (the stencils whose code path changed last): MCUPS at 256^3
for spec in "--stencil TENO6-A" "--stencil TENO6" "--stencil TENO5-A" "--stencil WENO6-CU" "--stencil TENO5"; do
  python bench.py --config tgv256 $spec --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernel_ms']
print('$spec |', round(d['value'],1), 'MCUPS', round(d['ms_per_step'],3), 'ms', {a:b for a,b in k.items() if b}, d['clocks']['sm_mhz'])"
done
