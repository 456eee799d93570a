"""CPU: every global name a function of the package (or of the bench / entry scripts) refers to exists — the GPU-only
code paths cannot be executed in the build container, so a missing import there would first show on the B200 box."""
import builtins
import importlib
import os
import symtable
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "handwriting_line_generation_b200"
MODULES = sorted(f"{PKG}.{f[:-3]}" for f in os.listdir(os.path.join(ROOT, PKG)) if f.endswith(".py") and f != "__init__.py")
SCRIPTS = ["bench", "bench_gan_train", "bench_hwr_train", "__graft_entry__"]


def _unresolved(table, module, path=""):
    bad = []
    for child in table.get_children():
        bad += _unresolved(child, module, f"{path}{child.get_name()}.")
    if table.get_type() == "module":
        return bad
    for sym in table.get_symbols():
        if sym.is_referenced() and sym.is_global() and not sym.is_assigned():
            name = sym.get_name()
            if not hasattr(module, name) and not hasattr(builtins, name):
                bad.append(f"{path[:-1]}: {name}")
    return bad


@pytest.mark.parametrize("modname", MODULES + SCRIPTS)
def test_every_global_name_resolves(modname, hwg_lib):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    module = importlib.import_module(modname)
    src = open(module.__file__).read()
    bad = _unresolved(symtable.symtable(src, module.__file__, "exec"), module)
    assert not bad, f"{modname}: unresolved global names: {bad}"
