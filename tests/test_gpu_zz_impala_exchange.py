"""GPU, >= 2 devices: IMPALA episode sampling across replay shards over NCCL."""
import os
import sys

import pytest
import torch as t

from conftest import ROOT, run_group

pytestmark = [pytest.mark.gpu]


@pytest.mark.timeout(900, method="thread")      # above the 600 s / 300 s limit of the torchrun launch itself
@pytest.mark.parametrize("world", [2, 4])
def test_impala_episode_exchange_over_nccl(world):
    if t.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29651 + world),
           os.path.join(ROOT, "tests", "dist_gpu_impala_worker.py")]
    out = run_group(cmd, timeout=300)
    assert out.returncode == 0 and "impala ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
