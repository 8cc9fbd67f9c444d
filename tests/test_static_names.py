"""CPU: every global name referenced anywhere in the package, bench.py and __graft_entry__.py resolves.

Most of the host code only runs with a CUDA device; a misspelt name there would otherwise first show up on
the GPU box.  Walks the byte code of every function (LOAD_GLOBAL / LOAD_NAME) and checks the name against the
module's globals and the builtins."""
import builtins
import dis
import importlib
import os
import pkgutil
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _modules():
    import jaxfluids_b200
    names = ["jaxfluids_b200"] + [m.name for m in pkgutil.walk_packages(jaxfluids_b200.__path__, "jaxfluids_b200.")]
    names = [n for n in names if n != "jaxfluids_b200.jax_ffi"]        # raises ImportError by design without jax
    return names + ["bench", "__graft_entry__"]


def _code_objects(code):
    yield code
    for c in code.co_consts:
        if isinstance(c, types.CodeType):
            yield from _code_objects(c)


@pytest.mark.parametrize("modname", _modules())
def test_global_names_resolve(modname):
    mod = importlib.import_module(modname)
    src = getattr(mod, "__file__", None)
    assert src and src.endswith(".py")
    with open(src) as fh:
        top = compile(fh.read(), src, "exec")
    known = set(vars(mod)) | set(vars(builtins))
    missing = []
    for code in _code_objects(top):
        stored = {i.argval for i in dis.get_instructions(code) if i.opname in ("STORE_NAME", "STORE_GLOBAL")}
        for ins in dis.get_instructions(code):
            if ins.opname == "LOAD_GLOBAL" and ins.argval not in known:
                missing.append((code.co_name, ins.argval, ins.positions.lineno if ins.positions else None))
            elif ins.opname == "LOAD_NAME" and ins.argval not in known and ins.argval not in stored \
                    and not ins.argval.startswith("__"):
                missing.append((code.co_name, ins.argval, ins.positions.lineno if ins.positions else None))
    assert not missing, f"unresolved names in {modname}: {missing}"


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under jaxfluids_b200/ (Python or CUDA sources) imports, names or opens
    anything under oracle/; bench.py may -- in its CPU legs only (cpu_reference_run)."""
    import ast
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "jaxfluids_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                continue
            text = open(os.path.join(dirpath, f), encoding="utf-8").read()
            assert not re.search(r"\b(import|from)\s+oracle\b", text), f
            assert "oracle/" not in text and "port_mt" not in text and "refharness" not in text, f
    # bench.py: every use of the oracle sits inside cpu_reference_run (the cpu_baseline / --impl reference legs) or the
    # parity leg, which runs BEFORE the timed region as the checker
    tree = ast.parse(open(os.path.join(root, "bench.py"), encoding="utf-8").read())
    allowed = {"cpu_reference_run", "parity_leg", "_oracle_tgv", "_oracle_hit"}     # the helpers of those two legs
    for node in ast.walk(tree):
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            names = [a.name for a in node.names] + [getattr(node, "module", "") or ""]
            if any(n == "oracle" or n.startswith("oracle.") for n in names):
                owner = next((fn.name for fn in ast.walk(tree) if isinstance(fn, ast.FunctionDef)
                              and any(child is node for child in ast.walk(fn))), None)
                assert owner in allowed, f"bench.py imports the oracle in {owner}"
    # ... and those helpers are reached from the two legs only
    src = open(os.path.join(root, "bench.py"), encoding="utf-8").read()
    for helper in ("_oracle_tgv", "_oracle_hit"):
        for fn in ast.walk(tree):
            if isinstance(fn, ast.FunctionDef) and fn.name not in allowed:
                body = ast.get_source_segment(src, fn) or ""
                assert helper + "(" not in body and ("= " + helper) not in body, (fn.name, helper)
