#!/bin/bash
# First GPU call after a CPU-only stretch: everything that was built and host-simulated
# without a GPU, in the order that localises a failure fastest.  Run under gpurun from the repo root:
#   gpurun --timeout 1500 -- 'bash scripts/gpu_verify_cpu_built.sh > gpurun_out/verify.log 2>&1'
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" || echo "SMOKE FAILED"
# 1. the tuned path must be what it was (same SASS as the last measured build): parity file first
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
# 2. the generic instantiations, Riemann variants, RK2_LS4, flux splitting, host-applied boundaries (no -x: list every failure)
python -m pytest tests/test_gpu_stencils.py -q -m gpu 2>&1 | tail -40
# 3. the shipped examples at their original sizes (sanity: finite, positive)
python scripts/run_fixture_cases_fullsize.py 200 2>&1 | tail -20
# 4. the headline number is unchanged
python bench.py --steps 10 --warmup 3 2>&1 | tail -1
