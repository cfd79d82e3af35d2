"""CPU: every name loaded anywhere in the Python sources resolves (tools/lint_names.py).  The GPU-only code paths
cannot be executed in the build container; this at least catches typos in them before a GPU box does."""
import os
import subprocess
import sys

from conftest import ROOT


def test_no_undefined_names():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "lint_names.py")], cwd=ROOT,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr


def test_every_referenced_library_symbol_exists():
    """Every entry point called on `lib` and every attribute read from the `_lib` module anywhere in the Python sources
    must exist in the ctypes shim: the code paths that use them only run on a GPU box, an AttributeError there
    costs a paid-for run."""
    import glob
    import re
    from machin_b200 import _lib
    files = [f for pat in ("machin_b200/**/*.py", "tools/*.py", "tests/*.py", "bench.py", "__graft_entry__.py")
             for f in glob.glob(os.path.join(ROOT, pat), recursive=True)]
    missing = []
    for f in files:
        src = open(f).read()
        for name in set(re.findall(r"\blib\.(mb_[a-z0-9_]+)", src)):
            if name not in _lib.SIGNATURES:
                missing.append((os.path.relpath(f, ROOT), "lib." + name))
        for name in set(re.findall(r"\b_lib\.([A-Za-z_][A-Za-z0-9_]*)", src)) - {"py"}:        # ("_lib.py" in prose)
            if not hasattr(_lib, name):
                missing.append((os.path.relpath(f, ROOT), "_lib." + name))
    assert not missing, missing


def test_every_library_call_has_the_declared_number_of_arguments():
    """Static arity check of every `lib.mb_*(...)` / `create(...)`-style direct call in the Python sources against the
    ctypes signatures (calls with *args are skipped): tools/ and the GPU-only branches get this without running."""
    import ast
    import glob
    from machin_b200 import _lib
    files = [f for pat in ("machin_b200/**/*.py", "tools/*.py", "tests/*.py", "bench.py", "__graft_entry__.py")
             for f in glob.glob(os.path.join(ROOT, pat), recursive=True)]
    wrong, seen = [], 0
    for f in files:
        for node in ast.walk(ast.parse(open(f).read())):
            if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute)):
                continue
            name = node.func.attr
            if not name.startswith("mb_") or name not in _lib.SIGNATURES or node.keywords:
                continue
            if any(isinstance(a, ast.Starred) for a in node.args):
                continue
            seen += 1
            want = len(_lib.SIGNATURES[name][1])
            if len(node.args) != want:
                wrong.append((os.path.relpath(f, ROOT), node.lineno, name, len(node.args), want))
    assert seen > 100 and not wrong, (seen, wrong)
