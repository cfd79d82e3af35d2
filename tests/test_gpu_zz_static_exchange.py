"""GPU, >= 2 devices: the fixed-shape APEX update (one CUDA graph, NCCL captured) against the dynamic exchange."""
import os
import subprocess
import sys

import pytest
import torch as t

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4])
def test_static_exchange_matches_dynamic_over_nccl(world):
    if t.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29631 + world),
           os.path.join(ROOT, "tests", "dist_gpu_static_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "static ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
