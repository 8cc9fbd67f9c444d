"""Time the three dissipative sweeps separately on a TGV block (default 512^3)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
im = InputManager(*bench.tgv_case(n, (1, 1, 1), 10, viscous=True))
buf = InitializationManager(im).initialization()
sim = SimulationManager(im)
rt = sim.runtime
rhs = rt.solver.new_rhs()
out = {}
for ax in range(3):
    for _ in range(2):
        rt.solver.dissipative_sweep(ax, rt.primitives, rhs, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        rt.solver.dissipative_sweep(ax, rt.primitives, rhs, True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out["xyz"[ax]] = (round(ms, 3), round(n ** 3 * 106.7 / (ms * 1e-3) / 1e9), "GB/s")
print(out)
