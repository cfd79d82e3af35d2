"""GPU parity of whole learner updates against outputs of the unmodified reference
(tests/golden/ref_td.npz, made by oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch as t

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


class QNet(t.nn.Module):
    def __init__(self, s, a, h=16):
        super().__init__()
        self.fc1 = t.nn.Linear(s, h)
        self.fc2 = t.nn.Linear(h, h)
        self.fc3 = t.nn.Linear(h, a)

    def forward(self, state):
        x = t.relu(self.fc1(state))
        return self.fc3(t.relu(self.fc2(x)))


def _load_flat(module, flat):
    o = 0
    with t.no_grad():
        for p in module.parameters():
            n = p.numel()
            p.copy_(t.from_numpy(flat[o:o + n].copy()).view_as(p))
            o += n


def _flat(module):
    return np.concatenate([p.detach().cpu().numpy().reshape(-1) for p in module.parameters()])


@pytest.mark.parametrize("criterion", ["mse_fused", "mse_generic", "mse_fused_graph"])
def test_dqn_per_update_vs_reference(criterion):
    from machin_b200.frame.algorithms import DQNPer
    g = np.load(os.path.join(GOLDEN, "ref_td.npz"))
    S, A, N, B = 6, 4, 200, 32
    qnet, qtgt = QNet(S, A).cuda(), QNet(S, A).cuda()
    _load_flat(qtgt, g["dqn_init_qt"])     # the ctor copies target -> online, like the reference
    crit = t.nn.MSELoss(reduction="sum")
    dqn = DQNPer(qnet, qtgt, t.optim.SGD, crit, batch_size=B, learning_rate=0.01, discount=0.97,
                 update_rate=0.05, replay_size=N)
    if criterion == "mse_generic":
        dqn._fused_mse = False
    assert np.array_equal(_flat(dqn.qnet), g["dqn_init_q"])
    ep = [dict(state={"state": t.from_numpy(g["dqn_state"][i:i + 1])}, action={"action": t.tensor([[g["dqn_action"][i]]])},
               next_state={"state": t.from_numpy(g["dqn_next_state"][i:i + 1])}, reward=float(g["dqn_reward"][i]),
               terminal=bool(g["dqn_terminal"][i])) for i in range(N)]
    dqn.replay_buffer.store_episode(ep, priorities=list(g["dqn_priority"]))
    assert np.array_equal(dqn.replay_buffer.wt_tree.weights, g["dqn_tree0"])
    if criterion == "mse_fused_graph":
        dqn.enable_cuda_graph()          # whole update replayed as one CUDA graph; state restored after warm-up
        assert np.array_equal(dqn.replay_buffer.wt_tree.weights, g["dqn_tree0"])
        assert np.array_equal(_flat(dqn.qnet), g["dqn_init_q"])
    np.random.seed(11)
    for step in range(3):
        loss = dqn.update()
        ref = float(g[f"dqn_loss{step}"])
        assert abs(loss - ref) <= 1e-5 * abs(ref), (step, loss, ref)
        np.testing.assert_allclose(dqn.replay_buffer.wt_tree.weights, g[f"dqn_tree{step + 1}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(_flat(dqn.qnet), g[f"dqn_q{step + 1}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(_flat(dqn.qnet_target), g[f"dqn_qt{step + 1}"], rtol=1e-5, atol=1e-6)
