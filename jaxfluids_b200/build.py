"""In-tree build of libjxf_b200.so (nvcc, sm_100a only)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
MAIN = os.path.join(CSRC, "jxf_b200.cu")
INST = os.path.join(CSRC, "sweep_inst.cu")        # compiled once per (axis, RECON) pair: the sweep kernel instantiations
SRC = [MAIN, INST]
DEPS = [os.path.join(CSRC, n) for n in ("numerics.cuh", "dissipative.cuh", "sweep_kernels.cuh", "plan.cuh")] + \
       [os.path.join(os.path.dirname(HERE), "include", "jxf_b200.h")]
OUT = os.path.join(HERE, "lib", "libjxf_b200.so")
OBJ = os.path.join(HERE, "lib", "obj")            # object files (git-ignored, gpurun-ignored: only the .so travels)
STAMP = OUT + ".srchash"          # written after a successful build; travels with the .so

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xfatbin", "-compress-all",      # 19 cubins with line info: 117 MB uncompressed
]
# (axis, RECON) pairs of the production library; tuning builds (-DJXF_TUNE_ONLY) keep the bench variant only
PAIRS = [(a, r) for a in range(3) for r in range(6)]
TUNE_PAIRS = [(a, 1) for a in range(3)]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def source_hash() -> str:
    """Content hash of everything the library is built from (sources, headers, flags)."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in SRC + DEPS:
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def up_to_date() -> bool:
    """Content-based, not mtime-based: a copy of the tree (the GPU box snapshot, a git checkout, a stash/pop) changes
    mtimes but must never trigger a multi-minute nvcc run inside a test or a bench."""
    if not os.path.exists(OUT):
        return False
    if not os.path.exists(STAMP):      # a library built before the stamp existed: fall back to modification times
        t = os.path.getmtime(OUT)
        return all(os.path.getmtime(p) <= t for p in SRC + DEPS)
    with open(STAMP) as fh:
        return fh.read().strip() == source_hash()


def _compile_all(out: str, defines, verbose: bool, tag: str) -> None:
    """One nvcc -c per translation unit, in parallel over the host cores, then one link."""
    from concurrent.futures import ThreadPoolExecutor
    nvcc = find_nvcc()
    tune = any(d.split("=")[0] == "JXF_TUNE_ONLY" for d in defines)
    limit = next((int(d.split("=")[1]) for d in defines if d.split("=")[0] == "JXF_RECON_LIMIT"), 6)
    objdir = os.path.join(OBJ, tag)
    os.makedirs(objdir, exist_ok=True)
    base = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c"]
    jobs = [(base + ["-o", os.path.join(objdir, "main.o"), MAIN], os.path.join(objdir, "main.o"))]
    # the generic instantiations (RECON 4 / 5) take longest: start them first
    for a, r in sorted(TUNE_PAIRS if tune else [p for p in PAIRS if p[1] < limit], key=lambda p: -p[1]):
        o = os.path.join(objdir, f"sweep_a{a}_r{r}.o")
        jobs.append((base + [f"-DJXF_INST_A={a}", f"-DJXF_INST_RECON={r}", "-o", o, INST], o))

    def run(job):
        cmd, _ = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stdout.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    workers = int(os.environ.get("JXF_BUILD_JOBS", str(max(1, min(len(jobs), os.cpu_count() or 1)))))
    print(f"[jaxfluids_b200.build] {len(jobs)} translation units, {workers} parallel nvcc jobs -> {out}", flush=True)
    with ThreadPoolExecutor(workers) as ex:
        list(ex.map(run, jobs))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + [o for _, o in jobs]
    subprocess.run(link, check=True)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    digest = source_hash()             # before the compile: an edit during the build must not be stamped as built
    if os.path.exists(STAMP):
        os.remove(STAMP)
    _compile_all(OUT, [], verbose, "prod")
    with open(STAMP, "w") as fh:
        fh.write(digest + "\n")
    return OUT


def build_variant(name: str, defines, verbose: bool = False) -> str:
    """Tuning / checking builds (e.g. JXF_TUNE_ONLY, JXF_REFERENCE_ORDER) next to the production library; selected at run
    time with JXF_LIB_VARIANT=<name>.  `defines` may hold nvcc flags too (entries starting with '-')."""
    out = OUT.replace(".so", f"_{name}.so")
    flags = [d for d in defines if d.startswith("-")]
    defs = [d for d in defines if not d.startswith("-")]
    global NVCC_FLAGS
    saved = NVCC_FLAGS
    NVCC_FLAGS = NVCC_FLAGS + flags
    try:
        _compile_all(out, defs, verbose, name)
    finally:
        NVCC_FLAGS = saved
    return out


def build_reforder(force: bool = False) -> str:
    """The checking build of tests/test_gpu_production.py: the reference's operations in the reference's order, no FMA
    contraction (-DJXF_REFERENCE_ORDER -fmad=false); the WENO5-Z instantiations (PRIMITIVE / CHAR-PRIMITIVE, every Riemann
    solver) are enough for the fixtures it is run on.  Selected with JXF_LIB_VARIANT=reforder."""
    out = OUT.replace(".so", "_reforder.so")
    stamp = out + ".srchash"
    digest = source_hash()
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return out
    build_variant("reforder", ["JXF_REFERENCE_ORDER", "JXF_RECON_LIMIT=2", "-fmad=false"])
    with open(stamp, "w") as fh:
        fh.write(digest + "\n")
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        rest = [x for x in sys.argv[i + 2:] if x != "--ptxas-v"]
        build_variant(sys.argv[i + 1], rest, verbose="--ptxas-v" in sys.argv)
    else:
        build(force="--force" in sys.argv, verbose="-v" in sys.argv)
