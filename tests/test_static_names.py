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
