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
        dqn.fuse_online_forwards = True     # one forward over [s; s'] must not change the trajectory
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


class Actor(t.nn.Module):
    def __init__(self, s, a, h=16):
        super().__init__()
        self.fc1 = t.nn.Linear(s, h)
        self.fc2 = t.nn.Linear(h, a)

    def forward(self, state):
        return t.tanh(self.fc2(t.relu(self.fc1(state))))


class Critic(t.nn.Module):
    def __init__(self, s, a, h=16):
        super().__init__()
        self.fc1 = t.nn.Linear(s + a, h)
        self.fc2 = t.nn.Linear(h, 1)

    def forward(self, state, action):
        return self.fc2(t.relu(self.fc1(t.cat([state, action], dim=1))))


def test_ddpg_per_update_vs_reference():
    from machin_b200.frame.algorithms import DDPGPer
    g = np.load(os.path.join(GOLDEN, "ref_td.npz"))
    S, A, N, B = 5, 2, 150, 16
    actor, actor_t, critic, critic_t = Actor(S, A).cuda(), Actor(S, A).cuda(), Critic(S, A).cuda(), Critic(S, A).cuda()
    _load_flat(actor_t, g["ddpg_init_actor_t"])
    _load_flat(critic_t, g["ddpg_init_critic_t"])
    ddpg = DDPGPer(actor, actor_t, critic, critic_t, t.optim.SGD, t.nn.MSELoss(reduction="sum"), batch_size=B,
                   actor_learning_rate=0.005, critic_learning_rate=0.01, discount=0.95, update_rate=0.1, replay_size=N)
    assert np.array_equal(_flat(ddpg.actor), g["ddpg_init_actor"]) and np.array_equal(_flat(ddpg.critic), g["ddpg_init_critic"])
    ep = [dict(state={"state": t.from_numpy(g["ddpg_state"][i:i + 1])}, action={"action": t.from_numpy(g["ddpg_action"][i:i + 1])},
               next_state={"state": t.from_numpy(g["ddpg_next_state"][i:i + 1])}, reward=float(g["ddpg_reward"][i]),
               terminal=bool(g["ddpg_terminal"][i])) for i in range(N)]
    ddpg.replay_buffer.store_episode(ep, priorities=list(g["ddpg_priority"]))
    np.random.seed(12)
    for step in range(3):
        pol, val = ddpg.update()
        ref_pol, ref_val = g[f"ddpg_loss{step}"]
        assert abs(pol - ref_pol) <= 1e-5 * max(1.0, abs(ref_pol)) and abs(val - ref_val) <= 1e-5 * max(1.0, abs(ref_val))
        np.testing.assert_allclose(ddpg.replay_buffer.wt_tree.weights, g[f"ddpg_tree{step + 1}"], rtol=1e-5, atol=1e-7)
        for name, net in (("actor", ddpg.actor), ("critic", ddpg.critic), ("actor_t", ddpg.actor_target),
                          ("critic_t", ddpg.critic_target)):
            np.testing.assert_allclose(_flat(net), g[f"ddpg_{name}{step + 1}"], rtol=1e-5, atol=1e-6)


class PolicyNet(t.nn.Module):
    def __init__(self, s, a, h=16):
        super().__init__()
        self.fc1 = t.nn.Linear(s, h)
        self.fc2 = t.nn.Linear(h, a)

    def forward(self, state, action=None):
        probs = t.softmax(self.fc2(t.relu(self.fc1(state))), dim=1)
        dist = t.distributions.Categorical(probs=probs)
        act = action if action is not None else dist.sample()
        return act.view(-1, 1), dist.log_prob(act.flatten()).view(-1, 1), dist.entropy().view(-1, 1)


class ValueNet(t.nn.Module):
    def __init__(self, s, h=16):
        super().__init__()
        self.fc1 = t.nn.Linear(s, h)
        self.fc2 = t.nn.Linear(h, 1)

    def forward(self, state):
        return self.fc2(t.relu(self.fc1(state)))


@pytest.mark.parametrize("tag", ["fixed", "ragged"])
def test_impala_update_vs_reference(tag):
    """Whole IMPALA.update() (episode sampling with Python's random, V-trace kernel, losses, SGD step)
    against the unmodified reference (tests/golden/ref_vtrace.npz)."""
    import random
    from machin_b200.frame.algorithms import IMPALA
    from machin_b200.parallel import LocalGroup
    g = np.load(os.path.join(GOLDEN, "ref_vtrace.npz"))
    lengths = [int(x) for x in g[f"{tag}_lengths"]]
    clip_c, clip_rho, discount = (float(x) for x in g[f"{tag}_clip"])
    S, A = 8, 3
    actor, critic = PolicyNet(S, A).cuda(), ValueNet(S).cuda()
    _load_flat(actor, g[f"{tag}_actor0"])
    _load_flat(critic, g[f"{tag}_critic0"])
    captured = {}

    def criterion(value, vs):
        captured["vs"] = vs.detach().clone()
        return ((value - vs) ** 2).sum()

    imp = IMPALA(actor, critic, t.optim.SGD, criterion, LocalGroup(device="cuda:0"), (), batch_size=len(lengths),
                 learning_rate=0.01, isw_clip_c=clip_c, isw_clip_rho=clip_rho, discount=discount,
                 replay_size=4 * sum(lengths))
    pos = 0
    for n in lengths:
        imp.store_episode([dict(state={"state": t.from_numpy(g[f"{tag}_state"][i:i + 1])},
                                action={"action": t.tensor([[int(g[f"{tag}_action"][i])]])},
                                next_state={"state": t.from_numpy(g[f"{tag}_next_state"][i:i + 1])},
                                reward=float(g[f"{tag}_reward"][i]), action_log_prob=float(g[f"{tag}_logmu"][i]),
                                terminal=bool(g[f"{tag}_terminal"][i])) for i in range(pos, pos + n)])
        pos += n
    t.manual_seed(3); np.random.seed(3); random.seed(3)
    pol, val = imp.update()
    ref_pol, ref_val = g[f"{tag}_losses"]
    np.testing.assert_allclose(captured["vs"].cpu().numpy().reshape(-1), g[f"{tag}_vs"], rtol=1e-5, atol=1e-6)
    assert abs(pol - ref_pol) <= 1e-5 * max(1.0, abs(ref_pol)) and abs(val - ref_val) <= 1e-5 * max(1.0, abs(ref_val))
    np.testing.assert_allclose(_flat(imp.actor), g[f"{tag}_actor1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(_flat(imp.critic), g[f"{tag}_critic1"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("lam", [1.0, 0.0, 0.95])
def test_a2c_store_episode_vs_reference(lam):
    """value / gae attached by store_episode (batched critic + reverse-scan kernel) vs the reference's
    per-step Python loop (tests/golden/ref_gae.npz)."""
    from machin_b200.frame.algorithms import A2C
    g = np.load(os.path.join(GOLDEN, "ref_gae.npz"))
    S, A = 8, 3
    actor, critic = PolicyNet(S, A).cuda(), ValueNet(S).cuda()
    _load_flat(critic, g["critic0"])
    a2c = A2C(actor, critic, t.optim.SGD, t.nn.MSELoss(reduction="sum"), batch_size=8, discount=0.97, gae_lambda=lam)
    for e in range(4):
        k = f"lam{lam}_ep{e}_"
        n = len(g[k + "reward"])
        ep = [dict(state={"state": t.from_numpy(g[k + "state"][i:i + 1])}, action={"action": t.tensor([[0]])},
                   next_state={"state": t.from_numpy(g[k + "next_state"][i:i + 1])}, reward=float(g[k + "reward"][i]),
                   terminal=bool(g[k + "terminal"][i])) for i in range(n)]
        a2c.store_episode(ep)
        assert [tr["value"] for tr in ep] == list(g[k + "value"])              # float64 returns: bit-exact
        np.testing.assert_allclose([tr["gae"] for tr in ep], g[k + "gae"], rtol=1e-5, atol=2e-6)
    bs, (state, adv) = a2c.replay_buffer.sample_batch(16, sample_attrs=["state", "gae"], additional_concat_custom_attrs=["gae"])
    assert adv.shape == (bs, 1) and adv.dtype == t.float32
    pol, val = a2c.update()
    assert np.isfinite(pol) and np.isfinite(val) and a2c.replay_buffer.size() == 0


# ----------------------------------------------------------------------------- fused MLP / Adam
# whole-K panels + cp.async (default) / 32 x 32 tiles, K in steps of 32 / 64 x 64 tiles + grid split-K
@pytest.mark.parametrize("legacy", ["wk", "small", "legacy"])
@pytest.mark.parametrize("sizes,batch", [([84, 256, 256, 6], 512), ([6, 16, 16, 4], 32), ([393, 300, 1], 100), ([5, 7], 3),
                                         ([84, 256, 256, 6], 1536), ([33, 65, 31], 8192), ([256, 256], 33)])
def test_fused_mlp_matches_torch(sizes, batch, legacy, monkeypatch):
    from machin_b200.model import FusedMLP
    monkeypatch.setenv("MB_GEMM", legacy)
    t.manual_seed(0)
    net = FusedMLP(sizes).cuda()
    ref = t.nn.Sequential()
    for i, layer in enumerate(net.layers):
        lin = t.nn.Linear(layer.in_features, layer.out_features).cuda()
        lin.load_state_dict(layer.state_dict())
        ref.append(lin)
        if i < len(net.layers) - 1:
            ref.append(t.nn.ReLU())
    x = t.randn(batch, sizes[0], device="cuda", requires_grad=True)
    x2 = x.detach().clone().requires_grad_()
    y, y2 = net(x), ref(x2)
    np.testing.assert_allclose(y.detach().cpu().numpy(), y2.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
    gy = t.randn_like(y)
    y.backward(gy)
    y2.backward(gy)
    np.testing.assert_allclose(x.grad.cpu().numpy(), x2.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
    # weight gradients are float32 sums over the batch in a blocking that differs from cuBLAS': the absolute rounding
    # noise of such a sum grows with its length (entries near zero next to entries of magnitude ~80 at batch 8192)
    atol = 1e-5 if batch <= 2048 else 5e-5
    for a, b in zip(net.parameters(), ref.parameters()):
        np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.cpu().numpy(), rtol=1e-4, atol=atol)
    with t.no_grad():
        assert net(x).shape == (batch, sizes[-1])


def test_fused_adam_matches_torch():
    from machin_b200.model import FusedAdam
    t.manual_seed(1)
    ps = [t.randn(s, device="cuda") for s in ((256, 84), (256,), (6, 256), (6,))]
    a = [p.clone().requires_grad_() for p in ps]
    b = [p.clone().requires_grad_() for p in ps]
    oa, ob = FusedAdam(a, lr=1e-3, weight_decay=0.01), t.optim.Adam(b, lr=1e-3, weight_decay=0.01)
    for step in range(5):
        gs = [t.randn_like(p) for p in ps]
        for x, y, g_ in zip(a, b, gs):
            x.grad, y.grad = g_.clone(), g_.clone()
        oa.step()
        ob.step()
    for x, y in zip(a, b):
        np.testing.assert_allclose(x.detach().cpu().numpy(), y.detach().cpu().numpy(), rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("use_program", [False, True])
def test_dqn_per_fused_net_vs_reference(use_program):
    """The reference fixture again, with the Q-nets as FusedMLP (csrc/mlp.cu) and the update in a CUDA graph."""
    from machin_b200.frame.algorithms import DQNPer
    from machin_b200.model import FusedMLP
    g = np.load(os.path.join(GOLDEN, "ref_td.npz"))
    S, A, N, B = 6, 4, 200, 32
    qnet, qtgt = FusedMLP([S, 16, 16, A]).cuda(), FusedMLP([S, 16, 16, A]).cuda()
    _load_flat(qtgt, g["dqn_init_qt"])
    dqn = DQNPer(qnet, qtgt, t.optim.SGD, t.nn.MSELoss(reduction="sum"), batch_size=B, learning_rate=0.01,
                 discount=0.97, update_rate=0.05, replay_size=N)
    ep = [dict(state={"state": t.from_numpy(g["dqn_state"][i:i + 1])}, action={"action": t.tensor([[g["dqn_action"][i]]])},
               next_state={"state": t.from_numpy(g["dqn_next_state"][i:i + 1])}, reward=float(g["dqn_reward"][i]),
               terminal=bool(g["dqn_terminal"][i])) for i in range(N)]
    dqn.replay_buffer.store_episode(ep, priorities=list(g["dqn_priority"]))
    dqn.use_fused_program = use_program
    dqn.enable_cuda_graph()
    assert (dqn._program is not None and dqn._program is not False) == use_program
    np.random.seed(11)
    for step in range(3):
        loss = dqn.update()
        ref = float(g[f"dqn_loss{step}"])
        assert abs(loss - ref) <= 1e-5 * abs(ref), (step, loss, ref)
        np.testing.assert_allclose(dqn.replay_buffer.wt_tree.weights, g[f"dqn_tree{step + 1}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(_flat(dqn.qnet), g[f"dqn_q{step + 1}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(_flat(dqn.qnet_target), g[f"dqn_qt{step + 1}"], rtol=1e-5, atol=1e-6)


def test_apex_static_exchange_matches_dynamic_single_member():
    """The fixed-shape single-graph update (padded wire blocks, device-side counts) against the dynamic
    exchange, on a group of one: same seeds -> same draws, same losses, same trees, same weights."""
    from machin_b200.frame.algorithms import DQNApex
    from machin_b200.parallel import LocalGroup
    S, A, N, B = 12, 5, 3000, 64

    def make(seed):
        t.manual_seed(seed)
        q, qt = QNet(S, A, 32).cuda(), QNet(S, A, 32).cuda()
        learner = DQNApex(q, qt, t.optim.SGD, t.nn.MSELoss(reduction="none"), LocalGroup(device="cuda:0"), (), batch_size=B,
                          learning_rate=0.05, discount=0.9, update_rate=0.1, replay_size=4096)
        rng = np.random.RandomState(3)
        learner.replay_buffer.store_episode_tensors(
            {"state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))},
             "action": {"action": t.from_numpy(rng.randint(0, A, size=(N, 1)))},
             "next_state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))}},
            {"reward": t.from_numpy(rng.randn(N, 1).astype(np.float32)), "terminal": t.from_numpy(rng.rand(N, 1) < 0.1)},
            priorities=rng.exponential(size=N))
        return learner

    a, b = make(0), make(0)
    b.enable_cuda_graph(static_exchange=True)
    assert np.array_equal(a.replay_buffer.wt_tree.weights, b.replay_buffer.wt_tree.weights)
    for step in range(6):
        np.random.seed(100 + step)
        la = a.update_device().item()
        np.random.seed(100 + step)
        lb = b.update_device().item()
        assert abs(la - lb) <= 1e-6 * max(1.0, abs(la)), (step, la, lb)
        np.testing.assert_allclose(b.replay_buffer.wt_tree.weights, a.replay_buffer.wt_tree.weights, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(_flat(b.qnet), _flat(a.qnet), rtol=1e-5, atol=1e-7)
        assert a.replay_buffer.curr_beta == b.replay_buffer.curr_beta
