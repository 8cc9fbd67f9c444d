"""GPU diagnostic (not a test): per-component error of the device face flux vs the oracle and vs
the host simulation, on the TGV fixture windows."""
import ctypes as C
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import port
from tests import helpers as H, hostsim
from jaxfluids_b200 import _lib

lib = _lib.load()
name = sys.argv[1] if len(sys.argv) > 1 else "tgv16_sym_char_hllc_rk3"
g, case, num = H.load_golden(name)
s = H.setup_from_json(case, num)
prims = g["prims0_halo"]
hl = hostsim.load(True)
for a in s.active:
    w = np.stack(port._window(prims, a, s), axis=-1)
    w = np.ascontiguousarray(np.moveaxis(w, 0, -2).reshape(-1, 5, 6))
    n = w.shape[0]
    ref = np.moveaxis(port.face_flux(prims, a, s), 0, -1).reshape(-1, 5)
    hs = np.empty((n, 5))
    hl.face_flux_host(a, 1 if s.recon == "CHAR-PRIMITIVE" else 0, 0 if s.riemann == "HLLC" else 1, w.ctypes.data, n, s.gamma, hs.ctypes.data)
    wd = torch.as_tensor(w).cuda()
    out = torch.empty((n, 5), dtype=torch.float64, device="cuda")
    rc = lib.jxf_debug_face_flux(a, 1 if s.recon == "CHAR-PRIMITIVE" else 0, 0 if s.riemann == "HLLC" else 1,
                                 C.c_void_p(wd.data_ptr()), n, s.gamma, C.c_void_p(out.data_ptr()), None)
    assert rc == 0
    torch.cuda.synchronize()
    gd = out.cpu().numpy()
    print("axis", a, "flux mag", np.abs(ref).max(0))
    print("   gpu  - ref :", np.abs(gd - ref).max(0))
    print("   host - ref :", np.abs(hs - ref).max(0))
    i = np.unravel_index(np.argmax(np.abs(gd - ref)), gd.shape)
    print("   worst face", i, "gpu", gd[i[0]], "ref", ref[i[0]])
    print("   window", w[i[0]])
