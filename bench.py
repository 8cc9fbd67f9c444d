#!/usr/bin/env python
"""bench.py -- cell-updates/s of the fused WENO5-Z / HLLC / SSP-RK3 step on the 3-D Taylor-Green
vortex (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...       # the CPU arm (oracle port on the host cores)

Workload: examples/examples_3D/01_tgv case (SYMMETRY walls, gamma 5/3, CHAR-PRIMITIVE WENO5-Z +
HLLC/EINFELDT + RK3, CFL 0.5, nh 5) at 512^3 cells PER GPU, weak-scaled with the case file's
block decomposition (1,1,1)/(2,1,1)/(2,2,1)/(2,2,2); the domain grows with the split so dx is
fixed.  A "step" is one full RK3 step (3 RHS evaluations + stage updates + halo fills + dt/min
reductions).  Rank 0 prints ONE JSON line.
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
ALG_BYTES_PER_CELL_STEP = 440.0      # 3*(40 + 80) + 2*40
ALG_FLOPS_PER_CELL_STEP = 9.7e3      # CHAR-PRIMITIVE + HLLC + EINFELDT, 3-D
# compulsory bytes per cell of ONE launch of each kernel kind (DESIGN.md "kernels"):
KERNEL_BYTES = {"sweep_x": 80.0, "sweep_y": 120.0, "sweep_z": 120.0,
                # prims in + rhs in + U in + U^n in (2 of 3 stages) + U out + prims out
                "sweep_x_epilogue": 226.7, "sweep_y_epilogue": 226.7, "sweep_z_epilogue": 226.7,
                # dissipative sweep: prims in (40) + 4 rhs rows in/out (64) (+8 for the mass row of the first)
                "dissipative": 106.7}


def tgv_case(cells_per_gpu: int, split, end_step: int, viscous: bool = False):
    sx, sy, sz = split
    dom = {}
    for ax, s in zip("xyz", split):
        dom[ax] = {"cells": cells_per_gpu * s, "range": [0.0, TWO_PI * s]}
    dom["decomposition"] = {"split_x": sx, "split_y": sy, "split_z": sz}
    case = {
        "general": {"case_name": "tgv", "end_step": int(end_step), "save_path": "./results"},
        "domain": dom,
        "boundary_conditions": {f: {"type": "SYMMETRY"} for f in ("east", "west", "north", "south", "top", "bottom")},
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
    num = {
        "conservatives": {"halo_cells": 5, "time_integration": {"integrator": "RK3", "CFL": 0.5},
                          "convective_fluxes": {"convective_solver": "GODUNOV", "godunov": {
                              "riemann_solver": "HLLC", "signal_speed": "EINFELDT",
                              "reconstruction_stencil": "WENO5-Z", "reconstruction_variable": "CHAR-PRIMITIVE"}}},
        "active_physics": {"is_convective_flux": True, "is_viscous_flux": False, "is_heat_flux": False,
                           "is_volume_force": False},
        "precision": {"is_double_precision_compute": True, "is_double_precision_output": True},
        "output": {"logging": {"level": "NONE"}},
    }
    if viscous:
        # the TGV at Re = 1600, Pr = 0.71: viscous + heat flux with the CENTRAL4 stencils
        num["active_physics"].update(is_viscous_flux=True, is_heat_flux=True)
        num["conservatives"]["dissipative_fluxes"] = {"reconstruction_stencil": "CENTRAL4",
                                                      "derivative_stencil_center": "CENTRAL4",
                                                      "derivative_stencil_face": "CENTRAL4"}
        case["material_properties"]["transport"] = {
            "dynamic_viscosity": {"model": "CUSTOM", "value": 1.0 / 1600.0}, "bulk_viscosity": 0.0,
            "thermal_conductivity": {"model": "PRANDTL", "prandtl_number": 0.71}}
    return case, num


# DRAM traffic per launch at 512^3 (dram__bytes_read.sum + dram__bytes_write.sum of ONE `ncu --set full` capture,
# profiles/r01b_ncu_full_sweeps.txt): equals the compulsory bytes above to within 3 % -- no wasted re-reads
# (the z+epilogue capture is stage 0, which does not read U^n: 210.7 B/cell against 186.7 compulsory + halo images)
NCU_TRAFFIC_512 = {"sweep_x": 5.668472e9 + 5.328012e9, "sweep_y": 11.035736e9 + 5.328945e9,
                   "sweep_z_epilogue": 16.883604e9 + 11.400052e9, "dissipative": 10.118630e9 + 4.270035e9}


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
def cpu_reference_run(steps: int, warmup: int, budget_s: float, threads: int | None = None):
    from oracle import port, port_mt
    threads = threads or os.cpu_count() or 1

    def make(n):
        s = port.Setup(cells=(n, n, n), domain=((0.0, TWO_PI),) * 3, bc={f: "SYMMETRY" for f in port.FACES},
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
    return {"value": mcups, "unit": "MCUPS", "cores": st.threads, "kind": "port",
            "sample": f"TGV {n}^3 (same case/numerics as the GPU arm), {steps} RK3 steps after {warmup} warm-up, "
                      f"NumPy oracle port slab-threaded over {st.threads} host threads; bit-identical to the "
                      f"reference sources run on the NumPy jax stand-in"}, el / steps * 1e3


def measure_e2e(args, rt, sim, buffers, tcv, time_now, dt_now, cells_global, world, barrier):
    """End to end through the public API with HOST buffers."""
    import torch
    import torch.distributed as dist
    k_e2e = args.e2e_steps or max(2, min(args.steps, 5))
    host_state = torch.empty(tuple(rt.primitives.shape), dtype=torch.float64, pin_memory=True)
    host_state.copy_(rt.primitives)
    jb = buffers._replace(time_control_variables=tcv._replace(physical_simulation_time=time_now,
                                                              physical_timestep_size=dt_now))
    def one_step(jb_):
        rt.solver.cons_from_prims(rt.primitives, rt.conservatives)
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
            rt.primitives.copy_(host_state, non_blocking=True)
            jb = one_step(jb)
    ms_serial = timed(serial)

    # (2) pipelined: the upload of step k+1's host input (copy stream, device staging buffer) overlaps step k;
    # every step still uploads its own input from pinned host memory and reads its result back inside the
    # timed region, and the first upload is not overlapped with anything
    k_pipe = max(k_e2e, args.steps)
    stagebuf = [torch.empty_like(rt.primitives) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    up = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[i % 2])
            stagebuf[i % 2].copy_(host_state, non_blocking=True)
            up[i % 2].record(copy_stream)

    def pipelined():
        nonlocal jb
        cur = torch.cuda.current_stream()
        for ev in free:
            ev.record(cur)
        upload(0)
        for k in range(k_pipe):
            if k + 1 < k_pipe:
                upload(k + 1)
            cur.wait_event(up[k % 2])
            rt.primitives.copy_(stagebuf[k % 2], non_blocking=True)
            free[k % 2].record(cur)
            jb = one_step(jb)
    ms_pipe = timed(pipelined)
    del stagebuf
    e2e = {"value": cells_global * k_pipe / (ms_pipe * 1e-3) / 1e6, "unit": "MCUPS",
           "h2d_bytes_per_step": int(host_state.numel() * 8), "d2h_bytes_per_step": 40,
           "steps": k_pipe, "ms_per_step": ms_pipe / k_pipe,
           "what": "per step: H2D of the block's halo'd primitive buffer from pinned host memory (copy stream, "
                   "device staging buffer; the upload of step k+1 overlaps step k, the first upload overlaps "
                   "nothing), device copy into the state, prim->cons on the device, "
                   "SimulationManager.do_integration_step, D2H of (t, dt, max speed, min rho, min p); "
                   "PCIe-bound (5.7 GB per step)",
           "serial": {"value": cells_global * k_e2e / (ms_serial * 1e-3) / 1e6, "steps": k_e2e,
                      "ms_per_step": ms_serial / k_e2e, "what": "same without any overlap (upload, then step)"}}

    return e2e


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=int(os.environ.get("JXF_BENCH_CELLS", "512")),
                    help="cells per GPU per axis (default 512: BASELINE config TGV 512^3 per GPU)")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--viscous", action="store_true",
                    help="widened workload (not the BASELINE line): TGV Re=1600 with viscous + heat flux")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = args.gpus
    split = SPLITS.get(n_gpus)
    if split is None:
        raise SystemExit(f"--gpus must be one of {sorted(SPLITS)}")
    config = {"workload": f"3D Taylor-Green vortex {args.cells}^3 cells per GPU (examples_3D/01_tgv case, SYMMETRY), "
                          f"WENO5-Z CHAR-PRIMITIVE + HLLC/EINFELDT + SSP-RK3, CFL 0.5, halo_cells 5, fp64",
              "cells_per_gpu": args.cells ** 3, "global_cells": args.cells ** 3 * n_gpus,
              "decomposition": list(split), "dx_fixed": True,
              "l2": "inputs larger than L2 (5.7 GB per field buffer at 512^3), no flush needed"}
    if args.viscous:
        config["workload"] += " + viscous and heat flux (CENTRAL4), Re 1600, Pr 0.71 [widened workload]"

    if args.impl == "reference":
        if rank != 0:
            return
        budget = float(os.environ.get("JXF_REF_BUDGET_S", "150"))
        cb, ms = cpu_reference_run(args.steps, args.warmup, budget)
        line = {"impl": "reference", "metric": "cell-updates/sec (MCUPS) per RK3 step, 3D TGV fp64", "value": cb["value"],
                "unit": "MCUPS", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "cpu_baseline": cb,
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

    case, num = tgv_case(args.cells, split, end_step=10 ** 9, viscous=args.viscous)
    im = InputManager(case, num)
    init = InitializationManager(im)
    buffers = init.initialization()
    sim = SimulationManager(im)
    rt = sim.runtime
    cells_global = args.cells ** 3 * n_gpus

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region ------------------------------------------------
    tcv = buffers.time_control_variables
    rt.set_time_control(tcv.physical_simulation_time, tcv.physical_timestep_size)
    sampler = ClockSampler(dev)
    sampler.start()                       # runs through warm-up + timed region (same workload throughout)
    for _ in range(args.warmup):
        rt.step()
    barrier()
    rt.solver.profile_read(reset=True)
    rt.solver.profile_enable(True)
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
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    mcups = cells_global * args.steps / (ms_total * 1e-3) / 1e6
    time_now, dt_now, _, min_rho, min_p = rt.read_step_scalars()
    assert np.isfinite([time_now, dt_now, min_rho, min_p]).all() and min_rho > 0 and min_p > 0, "solution blew up"

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
    dom_ms = sweeps[dom][0] / sweeps[dom][1]
    cells_local = args.cells ** 3
    achieved = cells_local * KERNEL_BYTES[dom] / (dom_ms * 1e-3) / 1e9
    kernel_ms = {k: (round(v[0] / v[1], 4) if v[1] else None) for k, v in prof.items()}
    share = {k: round(v[0] / (ms_total if world == 1 else sum(x[0] for x in prof.values())), 4) for k, v in prof.items()}
    t_hbm = ALG_BYTES_PER_CELL_STEP / (hbm_peak * 1e9)
    t_fp64 = ALG_FLOPS_PER_CELL_STEP / (fp64_tflops * 1e12)
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak,
        "traffic": (NCU_TRAFFIC_512.get(dom) if args.cells == 512 else None),
        "traffic_source": "profiles/r01b_ncu_full_sweeps.txt (ncu --set full, one launch, bytes)",
        "peak_source": peak_src,
        "bytes_per_cell_launch": KERNEL_BYTES[dom], "launch_ms": dom_ms,
        "note": "the path is FP64-pipe bound (AI ~22 flop/B): see step_roofline for the binding bound",
        "kernel_ms": kernel_ms, "kernel_share_of_step": share,
        "step_roofline": {
            "alg_bytes_per_cell_step": ALG_BYTES_PER_CELL_STEP, "alg_flops_per_cell_step": ALG_FLOPS_PER_CELL_STEP,
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
    if not args.no_e2e:
        try:
            e2e = measure_e2e(args, rt, sim, buffers, tcv, time_now, dt_now, cells_global, world, barrier)
        except Exception as exc:                      # never lose the device-resident line to the host-buffer leg
            if world > 1:
                raise                                 # ranks must stay in step: fail loudly instead of hanging
            e2e = {"value": None, "unit": "MCUPS", "error": f"{type(exc).__name__}: {exc}"}

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------
    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = cpu_reference_run(steps=3, warmup=1, budget_s=float(os.environ.get("JXF_CPU_BUDGET_S", "100")))

    if rank == 0:
        line = {"metric": "cell-updates/sec (MCUPS) per RK3 step, 3D TGV fp64", "value": mcups, "unit": "MCUPS",
                "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "cpu_baseline": cpu_baseline,
                "state": {"time": time_now, "dt": dt_now, "min_density": min_rho, "min_pressure": min_p}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
