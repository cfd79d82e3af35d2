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
