"""GPU: BASELINE config C1 as a parity case -- DQN with the uniform replay buffer (state 4, 2 actions, ring 10 000, batch
32), the three modes, five updates each against the unmodified reference (tests/golden/ref_dqn.npz).  Slots come from
Python's random stream exactly like buffer.py:234; rows are gathered on the device."""
import os
import random

import numpy as np
import pytest
import torch as t

from conftest import GOLDEN
from test_gpu_algorithms import QNet, _flat, _load_flat

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("mode", ["vanilla", "fixed_target", "double"])
def test_dqn_uniform_replay_vs_reference(mode):
    from machin_b200.frame.algorithms import DQN
    g = np.load(os.path.join(GOLDEN, "ref_dqn.npz"))
    S, A, N, B = 4, 2, 300, 32
    qnet, qtgt = QNet(S, A).cuda(), QNet(S, A).cuda()
    _load_flat(qtgt, g[f"{mode}_init_qt"])
    if mode == "vanilla":
        _load_flat(qnet, g[f"{mode}_init_q"])
    dqn = DQN(qnet, qtgt, t.optim.SGD, t.nn.MSELoss(), batch_size=B, learning_rate=0.2, discount=0.97, update_rate=0.1,
              replay_size=10000, mode=mode)
    assert np.array_equal(_flat(dqn.qnet), g[f"{mode}_init_q"]) and dqn.replay_buffer.buffer_device.type == "cuda"
    ep = [dict(state={"state": t.from_numpy(g["state"][i:i + 1])}, action={"action": t.tensor([[int(g["action"][i])]])},
               next_state={"state": t.from_numpy(g["next_state"][i:i + 1])}, reward=float(g["reward"][i]),
               terminal=bool(g["terminal"][i])) for i in range(N)]
    for lo in range(0, N, 50):
        dqn.store_episode(ep[lo:lo + 50])
    t.manual_seed(17); np.random.seed(17); random.seed(17)
    for step in range(5):
        loss = dqn.update()
        want = float(g[f"{mode}_loss{step}"])
        assert abs(loss - want) <= 1e-5 * abs(want), (mode, step, loss, want)
        np.testing.assert_allclose(_flat(dqn.qnet), g[f"{mode}_q{step + 1}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(_flat(dqn.qnet_target), g[f"{mode}_qt{step + 1}"], rtol=1e-5, atol=1e-6)
    assert dqn.act_discrete({"state": t.zeros(3, S, device="cuda")}).shape == (3, 1)
