"""GPU, >= 2 devices: the fixed-shape APEX update (one CUDA graph, NCCL captured) against the dynamic exchange."""
import os
import sys

import pytest
import torch as t

from conftest import ROOT, run_group

pytestmark = [pytest.mark.gpu]


@pytest.mark.timeout(900, method="thread")      # above the 600 s / 300 s limit of the torchrun launch itself
@pytest.mark.parametrize("world", [2, 4])
def test_static_exchange_matches_dynamic_over_nccl(world):
    if t.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29631 + world),
           os.path.join(ROOT, "tests", "dist_gpu_static_worker.py")]
    out = run_group(cmd, timeout=300)
    assert out.returncode == 0 and "static ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    # the peer-memory exchange must really have run (a box whose GPUs cannot be peer-mapped reports UNAVAILABLE)
    assert f"peer exchange ok (world {world})" in out.stdout, out.stdout[-3000:]
    assert f"sharded DDPG fixed-shape update ok (world {world})" in out.stdout, out.stdout[-3000:]


def test_device_apportion_matches_host_apportion_single_member():
    """Group of one: the device-side apportioning (shard sum all-gathered and apportioned inside the graph, host
    running ahead) must reproduce the host-apportioned fixed-shape update draw for draw -- with one member the
    apportioned count is the whole batch, so both consume np.random identically."""
    import numpy as np
    from machin_b200.frame.algorithms import DQNApex
    from machin_b200.parallel import LocalGroup
    S, A, N, B = 12, 5, 3000, 64

    class QNet(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2, self.fc3 = t.nn.Linear(S, 32), t.nn.Linear(32, 32), t.nn.Linear(32, A)

        def forward(self, state):
            return self.fc3(t.relu(self.fc2(t.relu(self.fc1(state)))))

    def make(device_apportion):
        t.manual_seed(0)
        q, qt = QNet().cuda(), QNet().cuda()
        learner = DQNApex(q, qt, t.optim.SGD, t.nn.MSELoss(reduction="none"), LocalGroup(device="cuda:0"), (), batch_size=B,
                          learning_rate=0.05, discount=0.9, update_rate=0.1, replay_size=4096)
        rng = np.random.RandomState(3)
        learner.replay_buffer.store_episode_tensors(
            {"state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))},
             "action": {"action": t.from_numpy(rng.randint(0, A, size=(N, 1)))},
             "next_state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))}},
            {"reward": t.from_numpy(rng.randn(N, 1).astype(np.float32)), "terminal": t.from_numpy(rng.rand(N, 1) < 0.1)},
            priorities=rng.exponential(size=N))
        learner.batched_draw = device_apportion          # the one-launch draw for all requesters rides along
        learner.enable_cuda_graph(static_exchange=True, device_apportion=device_apportion)
        return learner

    flat = lambda m: t.cat([p.detach().reshape(-1) for p in m.parameters()]).cpu().numpy()
    a, b = make(False), make(True)
    for step in range(8):
        np.random.seed(100 + step)
        la = a.update_device().item()
        np.random.seed(100 + step)
        lb = b.update_device().item()
        assert la == lb, (step, la, lb)
        assert np.array_equal(a.replay_buffer.wt_tree.weights, b.replay_buffer.wt_tree.weights)
        assert np.array_equal(flat(a.qnet), flat(b.qnet))
        assert a.replay_buffer.curr_beta == b.replay_buffer.curr_beta
    b.replay_buffer.wt_tree.check()
