"""GPU diagnostic (not a test): per-axis / per-field abs error of jxf_sweep vs the fixture."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import port
from tests import helpers as H
from tests.test_gpu_parity import make_solver, dev, host

name = sys.argv[1] if len(sys.argv) > 1 else "tgv16_sym_char_hllc_rk3"
g, case, num = H.load_golden(name)
s = H.setup_from_json(case, num)
sol = make_solver(s)
p0 = dev(g["prims0_halo"])
for a in s.active:
    rhs = sol.new_rhs()
    sol.sweep(a, p0, rhs, accumulate=False)
    got, ref = host(rhs), g[f"rhs_axis{a}"]
    err = np.abs(got - ref)
    print("axis", a, "abs err per field", err.reshape(5, -1).max(1), "mag", np.abs(ref).reshape(5, -1).max(1))
    i = np.unravel_index(np.argmax(err), err.shape)
    print("   worst", i, got[i], ref[i])
    # error pattern along each axis for the worst field
    v = i[0]
    print("   max err by x index", err[v].max(axis=(1, 2)))
    print("   max err by z index", err[v].max(axis=(0, 1)))
