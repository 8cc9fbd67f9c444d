"""In-tree build of libjxf_b200.so (nvcc, sm_100a only)."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", "jxf_b200.cu")]
DEPS = [os.path.join(HERE, "csrc", "numerics.cuh"), os.path.join(HERE, "csrc", "dissipative.cuh"), os.path.join(os.path.dirname(HERE), "include", "jxf_b200.h")]
OUT = os.path.join(HERE, "lib", "libjxf_b200.so")
STAMP = OUT + ".srchash"          # written after a successful build; travels with the .so

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


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


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + SRC
    print("[jaxfluids_b200.build]", " ".join(cmd), flush=True)
    digest = source_hash()             # before the compile: an edit during the build must not be stamped as built
    if os.path.exists(STAMP):
        os.remove(STAMP)
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as fh:
        fh.write(digest + "\n")
    return OUT


def build_variant(name: str, defines) -> str:
    """Tuning builds (e.g. JXF_MIN_BLOCKS=4) next to the production library; selected at run time with
    JXF_LIB_VARIANT=<name>.  Not used by tests or the default bench."""
    out = OUT.replace(".so", f"_{name}.so")
    cmd = [find_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", out] + SRC
    print("[jaxfluids_b200.build]", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
