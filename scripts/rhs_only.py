"""compute_rhs (three EPI=0 sweeps) on a 512^3 TGV block -- for ncu / timing of the pure sweeps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
im = InputManager(*bench.tgv_case(n, (1, 1, 1), 10))
buf = InitializationManager(im).initialization()
sim = SimulationManager(im)
rt = sim.runtime
rhs = rt.solver.new_rhs()
for _ in range(3):
    rt.solver.compute_rhs(rt.primitives, rhs)
torch.cuda.synchronize()
rt.solver.profile_enable(True)
for _ in range(5):
    rt.solver.compute_rhs(rt.primitives, rhs)
print({k: round(v[0] / max(v[1], 1), 3) for k, v in rt.solver.profile_read().items() if v[1]})
