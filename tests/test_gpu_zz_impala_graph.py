"""GPU: IMPALA's update for trajectories of one length captured as one CUDA graph (BASELINE C3's shape: T = 20)
reproduces the unmodified reference (tests/golden/ref_vtrace.npz, tag "fixed": 12 episodes of 20 steps) like the eager
update in tests/test_gpu_algorithms.py; a batch of another shape falls back to the eager path."""
import os
import random

import numpy as np
import pytest
import torch as t

from conftest import GOLDEN
from test_gpu_algorithms import PolicyNet, ValueNet, _flat, _load_flat

pytestmark = [pytest.mark.gpu]


def test_impala_captured_update_vs_reference():
    from machin_b200.frame.algorithms import IMPALA
    from machin_b200.parallel import LocalGroup
    g = np.load(os.path.join(GOLDEN, "ref_vtrace.npz"))
    tag = "fixed"
    lengths = [int(x) for x in g[f"{tag}_lengths"]]
    clip_c, clip_rho, discount = (float(x) for x in g[f"{tag}_clip"])
    S, A = 8, 3
    actor, critic = PolicyNet(S, A).cuda(), ValueNet(S).cuda()
    _load_flat(actor, g[f"{tag}_actor0"])
    _load_flat(critic, g[f"{tag}_critic0"])
    captured = {}

    def criterion(value, vs):
        captured["vs"] = vs            # (a static buffer of the graph: refreshed by every replay)
        return ((value - vs) ** 2).sum()

    imp = IMPALA(actor, critic, t.optim.SGD, criterion, LocalGroup(device="cuda:0"), (), batch_size=len(lengths),
                 learning_rate=0.01, isw_clip_c=clip_c, isw_clip_rho=clip_rho, discount=discount, replay_size=4 * sum(lengths))
    pos = 0
    for n in lengths:
        imp.store_episode([dict(state={"state": t.from_numpy(g[f"{tag}_state"][i:i + 1])},
                                action={"action": t.tensor([[int(g[f"{tag}_action"][i])]])},
                                next_state={"state": t.from_numpy(g[f"{tag}_next_state"][i:i + 1])},
                                reward=float(g[f"{tag}_reward"][i]), action_log_prob=float(g[f"{tag}_logmu"][i]),
                                terminal=bool(g[f"{tag}_terminal"][i])) for i in range(pos, pos + n)])
        pos += n
    t.manual_seed(3); np.random.seed(3); random.seed(3)
    imp.enable_cuda_graph(episode_length=lengths[0])
    assert np.array_equal(_flat(imp.actor), g[f"{tag}_actor0"]) and np.array_equal(_flat(imp.critic), g[f"{tag}_critic0"])
    assert imp.graph_launches >= 3                                # gather, unpack, V-trace
    pol, val = imp.update()
    ref_pol, ref_val = g[f"{tag}_losses"]
    np.testing.assert_allclose(captured["vs"].cpu().numpy().reshape(-1), g[f"{tag}_vs"], rtol=1e-5, atol=1e-6)
    assert abs(pol - ref_pol) <= 1e-5 * max(1.0, abs(ref_pol)) and abs(val - ref_val) <= 1e-5 * max(1.0, abs(ref_val))
    np.testing.assert_allclose(_flat(imp.actor), g[f"{tag}_actor1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(_flat(imp.critic), g[f"{tag}_critic1"], rtol=1e-5, atol=1e-6)
    # an episode of another length: batches that contain it run eagerly, the others keep replaying
    imp.store_episode([dict(state={"state": t.zeros(1, S)}, action={"action": t.tensor([[0]])}, next_state={"state": t.zeros(1, S)},
                            reward=0.0, action_log_prob=-1.0, terminal=i == 6) for i in range(7)])
    # (four updates only: plain SGD at lr 0.01 on a summed squared error diverges on this toy problem -- the same
    # schedule over host doubles reaches 1e9 at the sixth update and inf at the eighth)
    for _ in range(4):
        pol, val = imp.update()
        assert np.isfinite(pol) and np.isfinite(val)
