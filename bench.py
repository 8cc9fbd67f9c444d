#!/usr/bin/env python
"""bench.py -- cell-updates/s of the fused WENO5-Z / HLLC / SSP-RK3 step (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3                      # BASELINE config 4 at N = 1 (TGV 512^3 per GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W          # ... weak-scaled over N GPUs
    python bench.py --config tgv256 | riemann1024 | sod1000              # BASELINE configs 3 / 2 / 1 (one GPU)
    torchrun ... bench.py --gpus N --config hit1024                      # BASELINE config 5: 1024^3 STRONG scaling
    python bench.py --impl reference ...                                 # the CPU arm (oracle port on the host cores)

Default workload: examples/examples_3D/01_tgv case (SYMMETRY walls, gamma 5/3, CHAR-PRIMITIVE WENO5-Z + HLLC/EINFELDT +
RK3, CFL 0.5, nh 5) at 512^3 cells PER GPU, weak-scaled with the case file's block decomposition
(1,1,1)/(2,1,1)/(2,2,1)/(2,2,2); the domain grows with the split so dx is fixed.  `--config hit1024`: periodic
[0, 2 pi]^3, gamma 1.4, synthetic solenoidal initial condition (jaxfluids_b200/turbulence.py), GLOBAL grid fixed at
1024^3 and split over the ranks (strong scaling).  A "step" is one full RK3 step (3 RHS evaluations + stage updates +
halo fills / exchanges + dt / min reductions).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

TWO_PI = 6.283185307179586
SPLITS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
# SURVEY 8(d) / BASELINE.md section 2: algorithmic bytes and fp64 operations per cell and RK3 step (reference's own
# operation count, CHAR-PRIMITIVE + HLLC + EINFELDT); 2-D / 1-D have 2 / 1 sweeps per stage
ALG_BYTES_PER_CELL_STEP = 440.0
ALG_FLOPS_PER_RHS_AXIS = 3190.0 / 3.0       # one axis' share of the 3.19e3 operations of a 3-D RHS evaluation
ALG_FLOPS_EPILOGUE = 43.0                   # stage update + primitive recovery (130 per step / 3 stages)
# compulsory bytes per cell of ONE launch of each kernel kind (DESIGN.md "kernels"):
KERNEL_BYTES = {"sweep_x": 80.0, "sweep_y": 120.0, "sweep_z": 120.0,
                # prims in + rhs in + U in + U^n in (2 of 3 stages) + U out + prims out
                "sweep_x_epilogue": 226.7, "sweep_y_epilogue": 226.7, "sweep_z_epilogue": 226.7,
                # dissipative sweep: prims in (40) + 4 rhs rows in/out (64) (+8 for the mass row of the first)
                "dissipative": 106.7}
# DRAM traffic per launch at 512^3: dram__bytes_read.sum + dram__bytes_write.sum of ONE `ncu --set full` capture
# (a CONSTANT from the named file, not measured in this run: ncu cannot run inside the bench)
NCU_TRAFFIC_512 = {"file": "profiles/ncu_full_r02a_summary.txt",
                   "sweep_x": None, "sweep_y": None, "sweep_z_epilogue": None, "dissipative": None}
try:
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as _fh:
        NCU_TRAFFIC_512.update(json.load(_fh))
except Exception:
    pass

NUM_SETUP = {
    "conservatives": {"halo_cells": 5, "time_integration": {"integrator": "RK3", "CFL": 0.5},
                      "convective_fluxes": {"convective_solver": "GODUNOV", "godunov": {
                          "riemann_solver": "HLLC", "signal_speed": "EINFELDT",
                          "reconstruction_stencil": "WENO5-Z", "reconstruction_variable": "CHAR-PRIMITIVE"}}},
    "active_physics": {"is_convective_flux": True, "is_viscous_flux": False, "is_heat_flux": False,
                       "is_volume_force": False},
    "precision": {"is_double_precision_compute": True, "is_double_precision_output": True},
    "output": {"logging": {"level": "NONE"}},
}


def numerical_setup(viscous=False, stencil=None, recon=None, riemann=None, solver=None):
    num = json.loads(json.dumps(NUM_SETUP))
    cf = num["conservatives"]["convective_fluxes"]
    if stencil:
        cf["godunov"]["reconstruction_stencil"] = stencil
    if recon:
        cf["godunov"]["reconstruction_variable"] = recon
    if riemann:
        cf["godunov"]["riemann_solver"] = riemann
    if solver == "FLUX-SPLITTING":
        cf["convective_solver"] = "FLUX-SPLITTING"
        cf["flux_splitting"] = {"flux_splitting": "ROE", "reconstruction_stencil": stencil or "WENO5-Z"}
    if viscous:
        # the TGV at Re = 1600, Pr = 0.71: viscous + heat flux with the CENTRAL4 stencils
        num["active_physics"].update(is_viscous_flux=True, is_heat_flux=True)
        num["conservatives"]["dissipative_fluxes"] = {"reconstruction_stencil": "CENTRAL4",
                                                      "derivative_stencil_center": "CENTRAL4",
                                                      "derivative_stencil_face": "CENTRAL4"}
    return num


def _domain(cells, ranges, split):
    dom = {ax: {"cells": int(c), "range": [float(r[0]), float(r[1])]} for ax, c, r in zip("xyz", cells, ranges)}
    dom["decomposition"] = {"split_x": split[0], "split_y": split[1], "split_z": split[2]}
    return dom


def tgv_case(cells_per_gpu: int, split, end_step: int, viscous: bool = False, bc: str = "SYMMETRY", **num_kw):
    """examples/examples_3D/01_tgv/tgv.json at `cells_per_gpu`^3 per block; the domain grows with the split."""
    case = {
        "general": {"case_name": "tgv", "end_step": int(end_step), "save_path": "./results"},
        "domain": _domain([cells_per_gpu * s for s in split], [(0.0, TWO_PI * s) for s in split], split),
        "boundary_conditions": {f: {"type": bc} for f in ("east", "west", "north", "south", "top", "bottom")},
        "initial_condition": {
            "rho": 1.0,
            "u": "lambda x, y, z:  1.0 * jnp.sin(x / 1.0) * jnp.cos(y / 1.0) * jnp.cos(z / 1.0)",
            "v": "lambda x, y, z: -1.0 * jnp.cos(x / 1.0) * jnp.sin(y / 1.0) * jnp.cos(z / 1.0)",
            "w": 0.0,
            "p": "lambda x, y, z: 1.0 * 1.0**2 * (1 / 1.4 / 0.1**2 + 1/16.0 * ((jnp.cos(2 * x / 1.0) + "
                 "jnp.cos(2 * y / 1.0)) * (jnp.cos(2 * z / 1.0) + 2)))"},
        "material_properties": {"equation_of_state": {"model": "IdealGas", "specific_heat_ratio": 1.6666666666666667,
                                                      "specific_gas_constant": 1.0}},
    }
    if viscous:
        case["material_properties"]["transport"] = {
            "dynamic_viscosity": {"model": "CUSTOM", "value": 1.0 / 1600.0}, "bulk_viscosity": 0.0,
            "thermal_conductivity": {"model": "PRANDTL", "prandtl_number": 0.71}}
    return case, numerical_setup(viscous=viscous, **num_kw)


def hit_case(cells_global: int, split, end_step: int, **num_kw):
    """BASELINE config 5: periodic [0, 2 pi]^3, gamma 1.4; the initial condition is injected (user_prime_init)."""
    case = {
        "general": {"case_name": "hit", "end_step": int(end_step), "save_path": "./results"},
        "domain": _domain([cells_global] * 3, [(0.0, TWO_PI)] * 3, split),
        "boundary_conditions": {f: {"type": "PERIODIC"} for f in ("east", "west", "north", "south", "top", "bottom")},
        "initial_condition": {"rho": 1.0, "u": 0.0, "v": 0.0, "w": 0.0, "p": 1.0},
        "material_properties": {"equation_of_state": {"model": "IdealGas", "specific_heat_ratio": 1.4,
                                                      "specific_gas_constant": 1.0}},
    }
    return case, numerical_setup(**num_kw)


def sod_case(n: int, end_step: int, **num_kw):
    """examples/examples_1D/02_sod_shock_tube/sod.json with x.cells = n."""
    case = {"general": {"case_name": "sod", "end_step": int(end_step), "save_path": "./results"},
            "domain": _domain([n, 1, 1], [(0.0, 1.0)] * 3, (1, 1, 1)),
            "boundary_conditions": {"east": {"type": "ZEROGRADIENT"}, "west": {"type": "ZEROGRADIENT"},
                                    "north": {"type": "INACTIVE"}, "south": {"type": "INACTIVE"},
                                    "top": {"type": "INACTIVE"}, "bottom": {"type": "INACTIVE"}},
            "initial_condition": {"rho": "lambda x: 1.0*(x <= 0.5) + 0.125*(x > 0.5)", "u": 0.0, "v": 0.0, "w": 0.0,
                                  "p": "lambda x: 1.0*(x <= 0.5) + 0.1*(x > 0.5)"},
            "material_properties": {"equation_of_state": {"model": "IdealGas", "specific_heat_ratio": 1.4,
                                                          "specific_gas_constant": 1.0}}}
    return case, numerical_setup(**num_kw)


def riemann2d_case(n: int, end_step: int, **num_kw):
    """examples/examples_2D/07_riemann_problem/riemann2D.json (Lax-Liu configuration 3) at n^2."""
    def q(a, b, c, d):
        return (f"lambda x, y: ((x >= 0.5) & (y >= 0.5)) * {a} + ((x < 0.5) & (y >= 0.5)) * {b} + "
                f"((x < 0.5) & (y < 0.5)) * {c} + ((x >= 0.5) & (y < 0.5)) * {d}")
    case, num = sod_case(n, end_step, **num_kw)
    case["general"]["case_name"] = "riemann2D"
    case["domain"]["y"]["cells"] = n
    for f in ("north", "south"):
        case["boundary_conditions"][f] = {"type": "ZEROGRADIENT"}
    case["initial_condition"] = {"rho": q(1.5, 0.5323, 0.138, 0.5323), "u": q(0.0, 1.206, 1.206, 0.0),
                                 "v": q(0.0, 0.0, 1.206, 1.206), "w": 0.0, "p": q(1.5, 0.3, 0.029, 0.3)}
    return case, num


# name -> (workload, cells, scaling, default steps / warmup, BASELINE.json configs[] index)
CONFIGS = {
    "tgv512": ("tgv", 512, "weak", None, 3),
    "tgv256": ("tgv", 256, "weak", None, 2),
    "riemann1024": ("riemann2d", 1024, "weak", (200, 20), 1),
    "sod1000": ("sod", 1000, "weak", (2000, 200), 0),
    "hit1024": ("hit", 1024, "strong", None, 4),
}


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle-reason sampling DURING the timed region (B200_PROFILING.md clocks line).
    NVML in a thread (nvidia_ml_py; ~1 ms per sample, works on 8-GPU boxes where one `nvidia-smi` query takes
    hundreds of ms), falling back to an `nvidia-smi -lms` child process."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []
        self.samples, self.thread, self._stop, self.nvml = [], None, threading.Event(), None
        self.t_mark = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.gpu < len(ids) and ids[self.gpu].isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            R = pynvml
            bits = {"hw_slowdown": getattr(R, "nvmlClocksEventReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(R, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(R, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(R, "nvmlClocksEventReasonSwPowerCap", 0x4)}

            def reasons_of(h):
                try:
                    return pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    return pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)

            def pump():
                while not self._stop.is_set():
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        mask = reasons_of(h)
                        self.samples.append((time.perf_counter(), mhz, pw, {k for k, b in bits.items() if mask & b}))
                    except Exception:
                        pass
                    time.sleep(0.01)
            self.thread = threading.Thread(target=pump, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark(self):
        """Start of the timed region: only samples taken after this count (NVML path)."""
        self.t_mark = time.perf_counter()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=1)
            sel = [x for x in self.samples if self.t_mark is None or x[0] >= self.t_mark]
            if not sel:
                sel = self.samples
            if not sel:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
            reasons = set()
            for x in sel:
                reasons |= x[3]
            return {"sm_mhz": statistics.median(x[1] for x in sel), "sm_max_mhz": self.max_mhz,
                    "power_w_max": max(x[2] for x in sel), "samples": len(sel), "reasons": sorted(reasons),
                    "source": "nvml, 10 ms period, timed region only"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
            except ValueError:
                continue
            for n, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if len(sm) > 2:                    # the first samples may precede the first kernel
            sm, mx, pw = sm[1:], mx[1:], pw[1:]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons), "source": "nvidia-smi -lms 50"}


# ---------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores, bounded sample of the same workload
# ---------------------------------------------------------------------------
def _oracle_tgv(n, bc="SYMMETRY"):
    from oracle import port
    s = port.Setup(cells=(n, n, n), domain=((0.0, TWO_PI),) * 3, bc={f: bc for f in port.FACES},
                   gamma=1.6666666666666667)
    x, y, z = np.meshgrid(*s.cell_centers(), indexing="ij", sparse=True)
    pr = np.empty((5, n, n, n))
    pr[0] = 1.0
    pr[1] = np.sin(x) * np.cos(y) * np.cos(z)
    pr[2] = -np.cos(x) * np.sin(y) * np.cos(z)
    pr[3] = 0.0
    pr[4] = 1 / 1.4 / 0.1 ** 2 + 1 / 16.0 * ((np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2))
    with np.errstate(all="ignore"):
        prims, cons = port.initialize(pr, s)
    return s, prims, cons


def _oracle_hit(n):
    from oracle import port
    from jaxfluids_b200 import turbulence
    s = port.Setup(cells=(n, n, n), domain=((0.0, TWO_PI),) * 3, bc={f: "PERIODIC" for f in port.FACES}, gamma=1.4)
    with np.errstate(all="ignore"):
        prims, cons = port.initialize(turbulence.synthetic_solenoidal_ic(n), s)
    return s, prims, cons


def cpu_reference_run(steps: int, warmup: int, budget_s: float, threads: int | None = None, workload: str = "tgv"):
    """The NumPy oracle port, slab-threaded over the host cores, on the largest grid of the same case whose
    (steps + warmup) fit `budget_s`.  Returns (cpu_baseline object, ms per step, grid n)."""
    from oracle import port, port_mt
    threads = threads or os.cpu_count() or 1
    make = _oracle_hit if workload == "hit" else _oracle_tgv
    # size the sample: one probe step at 32^3, then the largest n whose (steps+warmup) fit the budget
    s, prims, cons = make(32)
    st = port_mt.ThreadedStepper(s, threads)
    dt = port.time_step_size(prims, s)
    t0 = time.time()
    st.step(prims, cons, dt)
    per_cell = (time.time() - t0) / 32 ** 3
    n = 32
    for cand in (48, 64, 96, 128):
        if per_cell * cand ** 3 * (steps + warmup) <= budget_s:
            n = cand
    s, prims, cons = make(n)
    st = port_mt.ThreadedStepper(s, threads)
    dt = port.time_step_size(prims, s)
    for _ in range(warmup):
        prims, cons, dt = st.step(prims, cons, dt)
    t0 = time.time()
    for _ in range(steps):
        prims, cons, dt = st.step(prims, cons, dt)
    el = time.time() - t0
    mcups = n ** 3 * steps / el / 1e6
    what = "synthetic isotropic turbulence, PERIODIC" if workload == "hit" else "TGV, SYMMETRY"
    return {"value": mcups, "unit": "MCUPS", "cores": st.threads, "kind": "port",
            "sample": f"{what} {n}^3 (same case / numerics as the GPU arm on a smaller grid), {steps} RK3 steps after "
                      f"{warmup} warm-up, NumPy oracle port slab-threaded over {st.threads} host threads; bit-identical "
                      f"to the reference sources run on the NumPy jax stand-in (not XLA)"}, el / steps * 1e3, n


# ---------------------------------------------------------------------------
# multi-rank parity leg (before the timed region): the decomposed GPU run against the single-block oracle
# ---------------------------------------------------------------------------
def parity_leg(split, rank, world):
    """TGV, PERIODIC, 32^3 cells per GPU with THIS run's decomposition, 3 steps through the same runtime the timed
    region uses (overlapped exchange, 3-layer slabs, fused halo images); rank 0 gathers the blocks and compares with
    oracle.port on the global grid.  -> {"err": rel Linf of the primitives, "dt_err", "nranks", ...} on rank 0."""
    import torch
    import torch.distributed as dist
    from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager
    n, steps = 32, 3
    case, num = tgv_case(n, split, end_step=steps, bc="PERIODIC")
    im = InputManager(case, num)
    buf = InitializationManager(im).initialization()
    sim = SimulationManager(im)
    rt = sim.runtime
    tcv = buf.time_control_variables
    rt.set_time_control(tcv.physical_simulation_time, tcv.physical_timestep_size)
    for _ in range(steps):
        rt.step()
    t, dt, _, _, _ = rt.read_step_scalars(complete_halos=True)
    nh = 5
    mine = rt.primitives[:, nh:-nh, nh:-nh, nh:-nh].cpu().numpy()
    di = im.domain_information
    item = (di.block_slices(rank), mine, dt, t)
    gathered = [item]
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, item)
    out = None
    if rank == 0:
        from oracle import port, port_mt
        cells = tuple(di.global_number_of_cells)
        s = port.Setup(cells=cells, domain=tuple((0.0, TWO_PI * sp) for sp in split),
                       bc={f: "PERIODIC" for f in port.FACES}, gamma=1.6666666666666667)
        x, y, z = np.meshgrid(*s.cell_centers(), indexing="ij", sparse=True)
        pr = np.empty((5,) + cells)
        pr[0] = 1.0
        pr[1] = np.sin(x) * np.cos(y) * np.cos(z)
        pr[2] = -np.cos(x) * np.sin(y) * np.cos(z)
        pr[3] = 0.0
        pr[4] = 1 / 1.4 / 0.1 ** 2 + 1 / 16.0 * ((np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2))
        with np.errstate(all="ignore"):
            p, c = port.initialize(pr, s)
        st = port_mt.ThreadedStepper(s, os.cpu_count() or 1)
        dto = port.time_step_size(p, s)
        to = 0.0
        for _ in range(steps):
            to += dto
            p, c, dto = st.step(p, c, dto)
        ref = p[(slice(None),) + s.interior]
        glob = np.empty_like(ref)
        for sl, arr, _, _ in gathered:
            glob[(slice(None),) + sl] = arr
        scale = np.array([max(np.max(np.abs(ref[v])), 1e-3 * np.max(np.abs(ref))) for v in range(5)])
        err = max(float(np.max(np.abs(glob[v] - ref[v])) / scale[v]) for v in range(5))
        out = {"err": err, "dt_err": max(abs(g[2] - dto) / dto for g in gathered), "t_err": abs(gathered[0][3] - to) / to,
               "nranks": world, "decomposition": list(split), "tol": 1e-12,
               "what": f"TGV PERIODIC {n}^3 per GPU (global {'x'.join(map(str, cells))}), {steps} RK3 steps through the "
                       f"bench's runtime (overlap={bool(rt.overlap)}, stage exchange layers={rt.stage_layers}) vs "
                       f"oracle.port on the global grid; err = max rel Linf over the 5 primitive fields"}
        assert err <= 1e-12 and out["dt_err"] <= 1e-12, f"multi-GPU parity failed: {out}"
    del sim, rt, buf
    torch.cuda.empty_cache()
    return out


def measure_e2e(args, rt, sim, buffers, tcv, time_now, dt_now, cells_global, world, barrier):
    """End to end through the public API with HOST buffers."""
    import torch
    import torch.distributed as dist
    k_e2e = args.e2e_steps or max(2, min(args.steps, 5))
    # the step's host input is what the reference's user hands over: the INTERIOR cells of the block's primitives
    # (material_fields_initializer.py:148-210); halos are filled on the device (outer rules + inter-block exchange)
    sl = (slice(None),) + tuple(rt.cfg.interior)
    host_state = torch.empty(tuple(rt.primitives[sl].shape), dtype=torch.float64, pin_memory=True)
    host_state.copy_(rt.primitives[sl])
    jb = buffers._replace(time_control_variables=tcv._replace(physical_simulation_time=time_now,
                                                              physical_timestep_size=dt_now))

    def one_step(jb_):
        rt.solver.cons_from_prims(rt.primitives, rt.conservatives)
        rt.halo_update(rt.primitives, rt.conservatives)
        mf = jb_.simulation_buffers.material_fields._replace(primitives=rt.primitives, conservatives=rt.conservatives)
        jb_ = jb_._replace(simulation_buffers=jb_.simulation_buffers._replace(material_fields=mf))
        return sim.do_integration_step(jb_)[0]          # public API; reads (t, dt, min rho, min p) back = D2H

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item())

    # (1) serial: upload, then step, then read back -- nothing overlaps
    def serial():
        nonlocal jb
        for _ in range(k_e2e):
            # H2D: the step's input state (the block's halo'd primitive buffer) from pinned host memory;
            # the conservatives are rebuilt from it on the device
            stagebuf[0].copy_(host_state, non_blocking=True)
            rt.primitives[sl].copy_(stagebuf[0], non_blocking=True)
            jb = one_step(jb)
    stagebuf = [torch.empty(tuple(host_state.shape), dtype=torch.float64, device=rt.primitives.device) for _ in range(2)]
    ms_serial = timed(serial)

    # (2) pipelined: the upload of step k+1's host input (copy stream, device staging buffer) overlaps step k;
    # every step still uploads its own input from pinned host memory and reads its result back inside the
    # timed region, and the first upload is not overlapped with anything.  `state_back`: the step's primitive
    # state also returns to pinned host memory (copy-back stream, overlapping the next step).
    k_pipe = max(k_e2e, args.steps)
    copy_stream = torch.cuda.Stream()
    back_stream = torch.cuda.Stream()
    up = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[i % 2])
            stagebuf[i % 2].copy_(host_state, non_blocking=True)
            up[i % 2].record(copy_stream)

    def pipelined(state_back, host_out, outbuf=None):
        nonlocal jb
        cur = torch.cuda.current_stream()
        for ev in free:
            ev.record(cur)
        upload(0)
        back_done = None
        for k in range(k_pipe):
            if k + 1 < k_pipe:
                upload(k + 1)
            cur.wait_event(up[k % 2])
            rt.primitives[sl].copy_(stagebuf[k % 2], non_blocking=True)
            free[k % 2].record(cur)
            jb = one_step(jb)
            if state_back:
                # result -> device staging buffer (fast), then D2H on the copy-back stream while the next step and the
                # next upload run; the staging buffer is reused once its copy-back has finished
                if back_done is not None:
                    cur.wait_event(back_done)
                outbuf.copy_(rt.primitives, non_blocking=True)
                done = torch.cuda.Event()
                done.record(cur)
                with torch.cuda.stream(back_stream):
                    back_stream.wait_event(done)
                    host_out.copy_(outbuf, non_blocking=True)
                    back_done = torch.cuda.Event()
                    back_done.record(back_stream)
        if back_done is not None:
            cur.wait_event(back_done)
    ms_pipe = timed(lambda: pipelined(False, None))
    nbytes = int(host_state.numel() * 8)
    e2e = {"value": cells_global * k_pipe / (ms_pipe * 1e-3) / 1e6, "unit": "MCUPS",
           "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 40,
           "steps": k_pipe, "ms_per_step": ms_pipe / k_pipe,
           "what": "per step: H2D of the block's primitive state (interior cells, as the reference's user supplies them) "
                   "from pinned host memory (copy stream, device staging buffer; the upload of step k+1 overlaps step k, "
                   "the first upload overlaps nothing), device copy into the halo'd state, prim->cons and the halo "
                   "update on the device, "
                   "SimulationManager.do_integration_step, D2H of the step's SCALARS ONLY (t, dt, max speed, min rho, "
                   "min p: 40 B) -- the state stays on the device; PCIe-bound (the upload)",
           "serial": {"value": cells_global * k_e2e / (ms_serial * 1e-3) / 1e6, "steps": k_e2e,
                      "ms_per_step": ms_serial / k_e2e, "what": "same without any overlap (upload, then step)"}}
    if world > 1:      # one more pinned buffer per rank: kept to single-GPU runs (8 ranks would pin 2 x 46 GB of host memory)
        e2e["state_back"] = {"value": None, "skipped": "measured at N = 1 only"}
        del stagebuf
        return e2e
    try:
        host_out = torch.empty(tuple(rt.primitives.shape), dtype=torch.float64, pin_memory=True)
        outbuf = torch.empty_like(rt.primitives)
        ms_back = timed(lambda: pipelined(True, host_out, outbuf))
        e2e["state_back"] = {"value": cells_global * k_pipe / (ms_back * 1e-3) / 1e6, "unit": "MCUPS", "steps": k_pipe,
                             "ms_per_step": ms_back / k_pipe, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": int(host_out.numel() * 8) + 40,
                             "what": "the pipelined leg with the step's halo'd primitive state ALSO copied back to pinned "
                                     "host memory every step (device staging copy, then D2H on a copy-back stream that "
                                     "overlaps the next step and the next upload; PCIe is full duplex)"}
        del host_out, outbuf
    except Exception as exc:
        e2e["state_back"] = {"value": None, "error": f"{type(exc).__name__}: {exc}"}
    del stagebuf
    return e2e


def time_small_config(rt, steps, warmup, flush_bytes=512 << 20):
    """Small working sets (Sod-1000, Riemann-1024^2 fit in L2): every step is timed on its own with CUDA events and
    an L2 flush (a 512 MB fill) between steps, outside the timed spans."""
    import torch
    flush = torch.empty(flush_bytes // 8, dtype=torch.float64, device="cuda")
    # launch-bound steps (Sod-1000: 4 launches around ~10 us of work): replay the step from a captured CUDA graph
    # (bit-identical to the launched step, tests/test_gpu_parity.py); JXF_BENCH_NO_GRAPH=1 launches kernel by kernel
    if os.environ.get("JXF_BENCH_NO_GRAPH", "0") != "1" and not rt.parallel.is_parallel:
        rt.use_cuda_graph(True)
    for _ in range(warmup):
        rt.step()
    torch.cuda.synchronize()
    total = 0.0
    evs = []
    for _ in range(steps):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rt.step()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    for a, b in evs:
        total += a.elapsed_time(b)
    return total


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS),
                    help="a BASELINE.json config (default tgv512 = configs[3], the one the metric is quoted on)")
    ap.add_argument("--cells", type=int, default=None,
                    help="cells per axis: per GPU (weak scaling) or of the global grid (strong scaling)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"])
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-rank parity leg before the timed region")
    ap.add_argument("--viscous", action="store_true",
                    help="widened workload (not the BASELINE line): TGV Re=1600 with viscous + heat flux")
    ap.add_argument("--stencil", default=None, help="widened workload: another reconstruction_stencil name")
    ap.add_argument("--recon", default=None, help="widened workload: another reconstruction_variable")
    ap.add_argument("--riemann", default=None, help="widened workload: another riemann_solver")
    ap.add_argument("--solver", default=None, help="widened workload: convective_solver (FLUX-SPLITTING)")
    args = ap.parse_args()
    cfg_name = args.config or ("tgv512" if args.cells in (None, 512) else None)
    if args.config is None and os.environ.get("JXF_BENCH_CELLS"):
        args.cells = args.cells or int(os.environ["JXF_BENCH_CELLS"])
        cfg_name = "tgv512" if args.cells == 512 else None
    workload, cells, scaling, kw, _ = CONFIGS[cfg_name or "tgv512"]
    cells = args.cells or cells
    scaling = args.scaling or scaling
    args.steps = args.steps or (kw[0] if kw else 10)
    args.warmup = args.warmup if args.warmup is not None else (kw[1] if kw else 3)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = args.gpus
    split = SPLITS.get(n_gpus)
    if split is None:
        raise SystemExit(f"--gpus must be one of {sorted(SPLITS)}")
    if workload in ("sod", "riemann2d") and n_gpus != 1:
        raise SystemExit(f"--config {cfg_name} is a one-GPU config")
    num_kw = dict(stencil=args.stencil, recon=args.recon, riemann=args.riemann, solver=args.solver)
    widened = args.viscous or any(num_kw.values())
    numerics = (f"{args.stencil or 'WENO5-Z'} {args.recon or 'CHAR-PRIMITIVE'} + "
                f"{(args.solver + ' ROE') if args.solver else (args.riemann or 'HLLC') + '/EINFELDT'} + SSP-RK3, CFL 0.5, "
                f"halo_cells 5, fp64")
    if workload == "tgv":
        per_gpu = cells ** 3 if scaling == "weak" else cells ** 3 // n_gpus
        glob = per_gpu * n_gpus
        what = (f"3D Taylor-Green vortex {cells}^3 cells per GPU (examples_3D/01_tgv case, SYMMETRY), " if scaling == "weak"
                else f"3D Taylor-Green vortex, global {cells}^3 cells split over the GPUs (SYMMETRY), ")
    elif workload == "hit":
        if scaling == "weak":
            raise SystemExit("the hit workload is the strong-scaling config (global grid fixed)")
        per_gpu, glob = cells ** 3 // n_gpus, cells ** 3
        what = (f"3D compressible isotropic turbulence, global {cells}^3 cells split over the GPUs (periodic [0,2pi]^3, "
                f"gamma 1.4, synthetic solenoidal IC: random-phase modes E(k)~k^4 exp(-2k^2/k0^2), k0=4, Ma_t=0.4, seed 0), ")
    elif workload == "riemann2d":
        per_gpu = glob = cells ** 2
        what = f"2D Riemann problem (Lax-Liu configuration 3) {cells}^2 cells (examples_2D/07_riemann_problem, ZEROGRADIENT), "
    else:
        per_gpu = glob = cells
        what = f"1D Sod shock tube {cells} cells (examples_1D/02_sod_shock_tube, ZEROGRADIENT), "
    config = {"workload": what + numerics, "baseline_config": cfg_name, "cells_per_gpu": per_gpu, "global_cells": glob,
              "decomposition": list(split), "dx_fixed": scaling == "weak"}
    if args.viscous:
        config["workload"] += " + viscous and heat flux (CENTRAL4), Re 1600, Pr 0.71"
    if widened:
        config["workload"] += " [widened workload, not a BASELINE line]"
    metric = {"tgv": "cell-updates/sec (MCUPS) per RK3 step, 3D TGV fp64",
              "hit": "cell-updates/sec (MCUPS) per RK3 step, 3D isotropic turbulence fp64",
              "riemann2d": "cell-updates/sec (MCUPS) per RK3 step, 2D Riemann problem fp64",
              "sod": "cell-updates/sec (MCUPS) per RK3 step, 1D Sod shock tube fp64"}[workload]

    if args.impl == "reference":
        if rank != 0:
            return
        budget = float(os.environ.get("JXF_REF_BUDGET_S", "150"))
        cb, ms, n = cpu_reference_run(args.steps, args.warmup, budget, workload="hit" if workload == "hit" else "tgv")
        # this arm's line names ITS OWN grid: the CPU port cannot run the GPU arm's grid in minutes
        ref_config = dict(config)
        ref_config.update(workload=f"{cb['sample']}", sample_of=config["workload"], cells_per_gpu=n ** 3, global_cells=n ** 3,
                          decomposition=[1, 1, 1], same_config=False,
                          note="bounded CPU sample: same case file, numerics and metric as the GPU arm, smaller grid; "
                               "MCUPS is per cell, so the two arms are comparable as throughputs, not as equal work")
        line = {"impl": "reference", "metric": metric, "value": cb["value"],
                "unit": "MCUPS", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": ref_config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "MCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    if world > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        opts = None
        if os.environ.get("JXF_COMM_PRIORITY", "1") != "0":
            try:                      # NCCL kernels on a high-priority stream: they overlap a sweep that fills the SMs
                opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            except Exception:
                opts = None
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local), pg_options=opts)
        if rank == 0:
            entry.build()
        dist.barrier()
    else:
        entry.build()
    assert world == n_gpus, f"--gpus {n_gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {n_gpus}"
    from jaxfluids_b200 import InputManager, InitializationManager, SimulationManager
    dev = torch.cuda.current_device()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity of the decomposed run (multi-rank; also cheap at N = 1), before anything is timed -------------
    parity = None
    if workload in ("tgv", "hit") and not args.no_parity and not widened:
        parity = parity_leg(split, rank, world)
        barrier()

    if workload == "tgv":
        n_blk = cells if scaling == "weak" else None
        if scaling == "weak":
            case, num = tgv_case(cells, split, end_step=10 ** 9, viscous=args.viscous, **num_kw)
        else:
            case, num = tgv_case(cells, (1, 1, 1), end_step=10 ** 9, viscous=args.viscous, **num_kw)
            case["domain"]["decomposition"] = {"split_x": split[0], "split_y": split[1], "split_z": split[2]}
    elif workload == "hit":
        case, num = hit_case(cells, split, end_step=10 ** 9, **num_kw)
    elif workload == "riemann2d":
        case, num = riemann2d_case(cells, 10 ** 9, **num_kw)
    else:
        case, num = sod_case(cells, 10 ** 9, **num_kw)
    im = InputManager(case, num)
    init = InitializationManager(im)
    if workload == "hit":
        from jaxfluids_b200 import turbulence

        def block_ic(block_slices, out):
            """this rank's block of the synthetic field, evaluated on the GPU straight into the state buffer"""
            turbulence.synthetic_solenoidal_ic(cells, gamma=1.4, k0=4.0, ma_t=0.4, seed=0, block=block_slices,
                                               device=f"cuda:{dev}", out=out)
        buffers = init.initialization(user_prime_init=block_ic)
    else:
        buffers = init.initialization()
    sim = SimulationManager(im)
    rt = sim.runtime
    cells_global = int(np.prod(im.domain_information.global_number_of_cells))
    field_gb = rt.primitives.numel() * 8 / 1e9
    config["memory_plan"] = getattr(rt, "memory_plan", "pingpong")
    config["stage_exchange_layers"] = rt.stage_layers if rt.neighbors else None

    # ---- device-resident timed region ------------------------------------------------
    tcv = buffers.time_control_variables
    rt.set_time_control(tcv.physical_simulation_time, tcv.physical_timestep_size)
    sampler = ClockSampler(dev)
    sampler.start()                       # runs through warm-up + timed region (same workload throughout)
    # working set (five field-sized buffers) within a few L2 sizes: time step by step with a flush in between
    small = workload in ("sod", "riemann2d") or 5 * field_gb < 0.5
    if small:
        config["l2"] = (f"working set {5 * field_gb * 1e3:.0f} MB: every step timed separately (CUDA events) with a "
                        f"512 MB L2 flush between steps, outside the timed spans; the step is replayed from a CUDA graph")
        rt.solver.profile_enable(False)
        sampler.mark()
        ms_total = time_small_config(rt, args.steps, args.warmup)
        # per-kernel times from a second, unflushed, kernel-by-kernel pass (shares only)
        rt.use_cuda_graph(False)
        rt.solver.profile_read(reset=True)
        rt.solver.profile_enable(True)
        for _ in range(min(args.steps, 50)):
            rt.step()
        torch.cuda.synchronize()
        prof = rt.solver.profile_read(reset=True)
        rt.solver.profile_enable(False)
        clocks = sampler.stop()
    else:
        config["l2"] = f"inputs larger than L2 ({field_gb:.2f} GB per field buffer), no flush needed"
        for _ in range(args.warmup):
            rt.step()
        barrier()
        rt.solver.profile_read(reset=True)
        rt.solver.profile_enable(True)
        rt.comm_profile(bool(rt.neighbors))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler.mark()
        ev0.record()
        for _ in range(args.steps):
            rt.step()
        ev1.record()
        barrier()
        clocks = sampler.stop()
        ms_total = ev0.elapsed_time(ev1)
        prof = rt.solver.profile_read(reset=True)
        rt.solver.profile_enable(False)
    comm_ms = rt.comm_profile_read() if rt.neighbors else {}
    rt.comm_profile(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    mcups = cells_global * args.steps / (ms_total * 1e-3) / 1e6
    time_now, dt_now, _, min_rho, min_p = rt.read_step_scalars()
    assert np.isfinite([time_now, dt_now, min_rho, min_p]).all() and min_rho > 0 and min_p > 0, "solution blew up"

    # ---- per-rank exchange times (max over ranks), per RK stage -------------------------------------------
    comm = None
    if rt.neighbors:
        keys = ("pack", "nccl", "unpack", "wait", "signal")
        v = torch.tensor([comm_ms.get(k, 0.0) for k in keys], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        nst = args.steps * rt.stages
        comm = {f"{k}_ms_per_stage": float(x) / nst for k, x in zip(keys, v.tolist())}
        peer = rt.peer is not None
        layers = rt.cfg.nh if peer else rt.stage_layers
        slab_mb = sum(rt.send[f].numel() for f in rt.neighbors) * 8 / 1e6 * layers / rt.cfg.nh
        comm.update(mode="peer" if peer else "nccl", faces=len(rt.neighbors), layers=layers, sent_mb_per_stage=slab_mb,
                    overlap=bool(rt.overlap),
                    what=("peer mode: the stage epilogue stores the shared faces' halo images straight into the "
                          "neighbours' buffers (CUDA IPC over NVLink, inside the sweep kernel's time); signal = the flag "
                          "kernel after each stage, wait = the flag spin before the first kernel that reads halos; "
                          "CUDA-event spans / stages, max over ranks") if peer else
                         ("CUDA-event spans, summed over the timed region / stages, max over ranks: pack / nccl / unpack "
                          "on the communication stream (nccl = the grouped send/recv batch incl. waiting for the peer), "
                          "wait = time the COMPUTE stream stalled on the exchange event"))

    # ---- FP64 pipe peak (measured here, same clocks) ------------------------------------
    import ctypes as C
    from jaxfluids_b200 import _lib
    lib = _lib.load()
    scratch = torch.empty(148 * 8 * 256, dtype=torch.float64, device="cuda")
    nf = C.c_int64()
    lib.jxf_fp64_probe(C.c_void_p(scratch.data_ptr()), 2000, C.byref(nf), None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.jxf_fp64_probe(C.c_void_p(scratch.data_ptr()), 40000, C.byref(nf), None)
    e1.record()
    torch.cuda.synchronize()
    fp64_tflops = 2.0 * nf.value / (e0.elapsed_time(e1) * 1e-3) / 1e12

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    sweeps = {k: v for k, v in prof.items() if k.startswith("sweep") and v[1] > 0}
    dom = max(sweeps, key=lambda k: sweeps[k][0])
    # time of the kernel kind PER RK STAGE: one launch in the default plan; the sum of a stage's launches when the sweep
    # is issued in pieces (slabs of the in-place plan, interior + strips of the overlapped exchange)
    n_stage_launches = max(1, (args.steps if not small else min(args.steps, 50)) * rt.stages)
    dom_ms = sweeps[dom][0] / n_stage_launches
    cells_local = int(np.prod(im.domain_information.device_number_of_cells))
    n_axes = len(rt.solver.active)
    flops_launch = ALG_FLOPS_PER_RHS_AXIS + (ALG_FLOPS_EPILOGUE if dom.endswith("epilogue") else 0.0)
    ach_tf = cells_local * flops_launch / (dom_ms * 1e-3) / 1e12
    ach_gbs = cells_local * KERNEL_BYTES[dom] / (dom_ms * 1e-3) / 1e9
    kernel_ms = {k: (round(v[0] / n_stage_launches, 4) if v[1] else None) for k, v in prof.items()}
    launches_per_stage = {k: round(v[2] / n_stage_launches, 2) for k, v in prof.items() if v[2]}
    tot_prof = sum(x[0] for x in prof.values())
    share = {k: round(v[0] / (ms_total if (world == 1 and not small) else max(tot_prof, 1e-30)), 4) for k, v in prof.items()}
    alg_flops_step = 3.0 * (n_axes * ALG_FLOPS_PER_RHS_AXIS + ALG_FLOPS_EPILOGUE)     # 9.7e3 in 3-D
    t_hbm = ALG_BYTES_PER_CELL_STEP / (hbm_peak * 1e9)
    t_fp64 = alg_flops_step / (fp64_tflops * 1e12)
    traffic = NCU_TRAFFIC_512.get(dom) if (cells_local == 512 ** 3 and not widened) else None
    roofline = {
        "bound": "fp64", "kernel": dom, "achieved": ach_tf, "peak": fp64_tflops, "unit": "TFLOP/s",
        "frac": ach_tf / fp64_tflops,
        "peak_source": "builder-measured in this run, same clocks: DFMA probe jxf_fp64_probe (8 independent chains per "
                       "thread, 148 x 8 CTAs x 256 threads, 2 flop per DFMA); MEASURED_PEAKS.json has no FP64 entry; "
                       "nominal 37.2 TFLOP/s at 1965 MHz",
        "alg_flops_per_cell_launch": flops_launch, "launch_ms": dom_ms,
        "flops_definition": "the reference's own elementwise operation count (SURVEY 8(a): 3190 per 3-D RHS evaluation = "
                            "1063 per axis, + 43 per stage update; div and sqrt count 1 each), NOT executed instructions",
        "traffic": traffic,
        "traffic_source": (f"{NCU_TRAFFIC_512['file']} (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full "
                           f"launch at 512^3 -- for the z + epilogue kernel the STAGE-0 instantiation, which reads no U^n: 200 B/cell "
                           f"algorithmic against 240 in stages 1 and 2; a constant from that capture, not measured in this run)"
                           if traffic else None),
        "hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                "bytes_per_cell_launch": KERNEL_BYTES[dom], "peak_source": peak_src},
        "kernel_ms": kernel_ms, "kernel_ms_is": "per RK stage (sum over the launches of the kind in one stage)",
        "launches_per_stage": launches_per_stage, "kernel_share_of_step": share,
        "step_roofline": {
            "alg_bytes_per_cell_step": ALG_BYTES_PER_CELL_STEP, "alg_flops_per_cell_step": alg_flops_step,
            "fp64_peak_tflops_measured": fp64_tflops, "hbm_bound_mcups": 1e-6 / t_hbm, "fp64_bound_mcups": 1e-6 / t_fp64,
            "roof_mcups_per_gpu": 1e-6 / max(t_hbm, t_fp64),
            "frac_of_roof": (mcups / n_gpus) / (1e-6 / max(t_hbm, t_fp64)),
            "hbm_frac_step": (mcups / n_gpus) * 1e6 * ALG_BYTES_PER_CELL_STEP / (hbm_peak * 1e9)},
    }
    if args.viscous and prof.get("dissipative", (0, 0, 0))[1]:
        d_ms = prof["dissipative"][0] / prof["dissipative"][1]
        d_gbs = cells_local * KERNEL_BYTES["dissipative"] / (d_ms * 1e-3) / 1e9
        roofline["dissipative_sweep"] = {"bound": "hbm", "launch_ms": d_ms, "bytes_per_cell_launch": KERNEL_BYTES["dissipative"],
                                         "achieved": d_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": d_gbs / hbm_peak}
    launches = int(sum(v[2] for v in prof.values()))

    # ---- end-to-end through the public API with HOST buffers ---------------------------------
    e2e = None
    if not args.no_e2e and workload == "hit":
        e2e = {"value": None, "unit": "MCUPS", "skipped": "strong-scaling config: the per-step host upload of a 44 GB "
               "global state is not what this config measures; see the default (tgv512) line for e2e"}
    elif not args.no_e2e:
        try:
            e2e = measure_e2e(args, rt, sim, buffers, tcv, time_now, dt_now, cells_global, world, barrier)
        except Exception as exc:                      # never lose the device-resident line to the host-buffer leg
            if world > 1:
                raise                                 # ranks must stay in step: fail loudly instead of hanging
            e2e = {"value": None, "unit": "MCUPS", "error": f"{type(exc).__name__}: {exc}"}

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------
    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cpu_baseline, _, _ = cpu_reference_run(steps=3, warmup=1, budget_s=float(os.environ.get("JXF_CPU_BUDGET_S", "100")),
                                               workload="hit" if workload == "hit" else "tgv")

    if rank == 0:
        line = {"metric": metric, "value": mcups, "unit": "MCUPS",
                "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": cpu_baseline, "parity": parity, "comm": comm,
                "state": {"time": time_now, "dt": dt_now, "min_density": min_rho, "min_pressure": min_p}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
