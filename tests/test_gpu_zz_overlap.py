"""GPU: the opt-in stream-level parallelism of one update (DQNPer.overlap_priority_update: priority write-back
beside backward + optimizer; DQNPer.parallel_forwards: the two no-grad forwards beside the online one) must not
change the trajectory -- eager and captured."""
import numpy as np
import pytest
import torch as t



pytestmark = [pytest.mark.gpu]


class QNet(t.nn.Module):
    def __init__(self, s, a, h):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = t.nn.Linear(s, h), t.nn.Linear(h, h), t.nn.Linear(h, a)

    def forward(self, state):
        return self.fc3(t.relu(self.fc2(t.relu(self.fc1(state)))))


def _flat(m):
    return t.cat([p.detach().reshape(-1) for p in m.parameters()]).cpu().numpy()


@pytest.mark.parametrize("flags", [("overlap_priority_update",), ("parallel_forwards",), ("fuse_optimizer_and_target",),
                                   ("overlap_priority_update", "parallel_forwards", "fuse_optimizer_and_target")])
@pytest.mark.parametrize("graph", [False, True])
def test_stream_parallel_update_same_trajectory(graph, flags):
    from machin_b200.frame.algorithms import DQNPer
    from machin_b200.model import FusedAdam
    S, A, N, B = 12, 5, 3000, 64

    def make(overlap):
        t.manual_seed(0)
        q, qt = QNet(S, A, 32).cuda(), QNet(S, A, 32).cuda()
        learner = DQNPer(q, qt, FusedAdam, t.nn.MSELoss(reduction="none"), batch_size=B, learning_rate=1e-3,
                         discount=0.9, update_rate=0.1, replay_size=4096)
        rng = np.random.RandomState(3)
        learner.replay_buffer.store_episode_tensors(
            {"state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))},
             "action": {"action": t.from_numpy(rng.randint(0, A, size=(N, 1)))},
             "next_state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))}},
            {"reward": t.from_numpy(rng.randn(N, 1).astype(np.float32)), "terminal": t.from_numpy(rng.rand(N, 1) < 0.1)},
            priorities=rng.exponential(size=N))
        for f in flags:
            assert hasattr(learner, f)
            setattr(learner, f, overlap)
        if graph:
            learner.enable_cuda_graph()
        return learner

    a, b = make(False), make(True)
    for step in range(8):
        np.random.seed(50 + step)
        la = a.update_device().item()
        np.random.seed(50 + step)
        lb = b.update_device().item()
        # same kernels on the same inputs: equal up to cuBLAS picking another algorithm when a second stream is busy
        assert abs(la - lb) <= 1e-6 * max(1.0, abs(la)), (step, la, lb)
        np.testing.assert_allclose(b.replay_buffer.wt_tree.weights, a.replay_buffer.wt_tree.weights, rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(_flat(b.qnet), _flat(a.qnet), rtol=1e-5, atol=1e-7)
    b.replay_buffer.wt_tree.check()
