"""GPU: DDPGPer's whole update captured as one CUDA graph reproduces the reference trajectory (ref_td.npz), like the
eager update in tests/test_gpu_algorithms.py; enabling the graph leaves networks, tree and beta untouched; after
release_graphs() the eager path continues the same trajectory."""
import os

import numpy as np
import pytest
import torch as t

from conftest import GOLDEN
from test_gpu_algorithms import Actor, Critic, _flat, _load_flat

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("release_after", [None, 1])
def test_ddpg_per_captured_update_vs_reference(release_after):
    from machin_b200.frame.algorithms import DDPGPer
    g = np.load(os.path.join(GOLDEN, "ref_td.npz"))
    S, A, N, B = 5, 2, 150, 16
    actor, actor_t, critic, critic_t = Actor(S, A).cuda(), Actor(S, A).cuda(), Critic(S, A).cuda(), Critic(S, A).cuda()
    _load_flat(actor_t, g["ddpg_init_actor_t"])
    _load_flat(critic_t, g["ddpg_init_critic_t"])
    ddpg = DDPGPer(actor, actor_t, critic, critic_t, t.optim.SGD, t.nn.MSELoss(reduction="sum"), batch_size=B,
                   actor_learning_rate=0.005, critic_learning_rate=0.01, discount=0.95, update_rate=0.1, replay_size=N)
    ep = [dict(state={"state": t.from_numpy(g["ddpg_state"][i:i + 1])}, action={"action": t.from_numpy(g["ddpg_action"][i:i + 1])},
               next_state={"state": t.from_numpy(g["ddpg_next_state"][i:i + 1])}, reward=float(g["ddpg_reward"][i]),
               terminal=bool(g["ddpg_terminal"][i])) for i in range(N)]
    ddpg.replay_buffer.store_episode(ep, priorities=list(g["ddpg_priority"]))
    tree0 = ddpg.replay_buffer.wt_tree.weights.copy()
    ddpg.enable_cuda_graph()
    assert np.array_equal(ddpg.replay_buffer.wt_tree.weights, tree0) and ddpg.replay_buffer.curr_beta == 0.4
    assert np.array_equal(_flat(ddpg.actor), g["ddpg_init_actor"]) and np.array_equal(_flat(ddpg.critic), g["ddpg_init_critic"])
    assert ddpg.graph_launches >= 6           # draw, gather, 2 x target, write-back, 2 x soft update
    np.random.seed(12)
    for step in range(3):
        if release_after == step:
            t.cuda.synchronize()
            ddpg.release_graphs()
        pol, val = ddpg.update()
        ref_pol, ref_val = g[f"ddpg_loss{step}"]
        assert abs(pol - ref_pol) <= 1e-5 * max(1.0, abs(ref_pol)) and abs(val - ref_val) <= 1e-5 * max(1.0, abs(ref_val))
        np.testing.assert_allclose(ddpg.replay_buffer.wt_tree.weights, g[f"ddpg_tree{step + 1}"], rtol=1e-5, atol=1e-7)
        for name, net in (("actor", ddpg.actor), ("critic", ddpg.critic), ("actor_t", ddpg.actor_target),
                          ("critic_t", ddpg.critic_target)):
            np.testing.assert_allclose(_flat(net), g[f"ddpg_{name}{step + 1}"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("device_apportion", [False, True])
def test_ddpg_apex_fixed_shape_update_matches_dynamic_single_member(device_apportion):
    """DDPGApex.enable_cuda_graph (the sharded DDPG update as ONE graph over the fixed-shape exchange) on a group of
    one against the dynamic, eager exchange: same draws, same losses, same networks, same tree.  (The merged batch of
    the fixed-shape update carries one padding row per member: the critic loss weighs it 0, the policy loss masks it.)"""
    from machin_b200.frame.algorithms import DDPGApex
    from machin_b200.parallel import LocalGroup
    S, A, N, B = 7, 3, 2000, 48

    def make(static):
        t.manual_seed(1)
        nets = [Actor(S, A).cuda(), Actor(S, A).cuda(), Critic(S, A).cuda(), Critic(S, A).cuda()]
        learner = DDPGApex(*nets, t.optim.SGD, t.nn.MSELoss(reduction="none"), LocalGroup(device="cuda:0"), (), batch_size=B,
                           actor_learning_rate=0.01, critic_learning_rate=0.02, discount=0.9, update_rate=0.1, replay_size=4096)
        rng = np.random.RandomState(5)
        learner.replay_buffer.store_episode_tensors(
            {"state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))},
             "action": {"action": t.from_numpy(rng.rand(N, A).astype(np.float32) * 2 - 1)},
             "next_state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))}},
            {"reward": t.from_numpy(rng.randn(N, 1).astype(np.float32)), "terminal": t.from_numpy(rng.rand(N, 1) < 0.1)},
            priorities=rng.exponential(size=N))
        if static:
            learner.enable_cuda_graph(device_apportion=device_apportion, peer_exchange=False)
        return learner

    a, b = make(False), make(True)
    assert np.array_equal(a.replay_buffer.wt_tree.weights, b.replay_buffer.wt_tree.weights)
    for net_a, net_b in zip((a.actor, a.critic), (b.actor, b.critic)):
        assert np.array_equal(_flat(net_a), _flat(net_b))
    for step in range(6):
        np.random.seed(40 + step)
        pa, va = (x.item() for x in a.update_device())
        np.random.seed(40 + step)
        pb, vb = (x.item() for x in b.update_device())
        assert abs(pa - pb) <= 1e-5 * max(1.0, abs(pa)) and abs(va - vb) <= 1e-5 * max(1.0, abs(va)), (step, pa, pb, va, vb)
        np.testing.assert_allclose(b.replay_buffer.wt_tree.weights, a.replay_buffer.wt_tree.weights, rtol=1e-5, atol=1e-9)
        for net_a, net_b in ((a.actor, b.actor), (a.critic, b.critic), (a.actor_target, b.actor_target),
                             (a.critic_target, b.critic_target)):
            np.testing.assert_allclose(_flat(net_b), _flat(net_a), rtol=1e-4, atol=1e-6)
    b.replay_buffer.wt_tree.check()
    t.cuda.synchronize()
    b.release_graphs()
