"""CPU: whole learner updates through the product's own algorithm classes, with host-memory test doubles under
them -- the ring (tests/fake_ring.py) and, for the CUDA kernels, THE KERNELS' DEVICE SOURCE compiled for the host
(tests/cpu_emul) -- against fixtures of the unmodified reference (tests/golden/ref_vtrace.npz: IMPALA.update()).
The same test runs on the device in tests/test_gpu_algorithms.py; this one keeps the Python between the kernels
(episode sampling with Python's random, loss assembly, optimizer steps) honest where no GPU exists."""
import ctypes as C
import os
import random
import subprocess
from threading import RLock

import numpy as np
import pytest
import torch as t

from conftest import GOLDEN, ROOT
from fake_ring import make_cpu_buffer

EMUL = os.path.join(ROOT, "tests", "cpu_emul")


@pytest.fixture(scope="module")
def scan_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "scan_emul.so")
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread", "-shared", "-fPIC",
           "-I" + os.path.join(ROOT, "machin_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(EMUL, "scan_harness.cpp"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    return C.CDLL(out)


def _emulated_vtrace(lib):
    def vtrace(value, next_value, reward, terminal, log_pi, log_mu, discount, clip_c=1.0, clip_rho=1.0, offsets=None,
               n_episodes=None, max_len=0):
        f = lambda x: np.ascontiguousarray(x.detach().numpy().reshape(-1), dtype=np.float32)
        p = lambda a: C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)
        n = value.numel()
        off = None if offsets is None else np.ascontiguousarray(offsets.numpy(), dtype=np.int64)
        n_ep = int(n_episodes) if off is None else len(off) - 1
        term = np.ascontiguousarray(terminal.numpy().reshape(-1).astype(np.uint8))
        out = {k: np.full(n, np.nan, np.float32) for k in ("vs", "vss", "rho", "adv")}
        assert lib.emul_vtrace(p(off), C.c_int64(n_ep), C.c_int64(n), p(f(value)), p(f(next_value)), p(f(reward)), p(term),
                               p(f(log_pi)), p(f(log_mu)), C.c_float(discount), C.c_float(clip_c), C.c_float(clip_rho),
                               p(out["vs"]), p(out["vss"]), p(out["rho"]), p(out["adv"]), C.c_int64(int(max_len))) == 0
        return {k: t.from_numpy(v).view(value.shape) for k, v in out.items()}
    return vtrace


@pytest.mark.parametrize("tag", ["fixed", "ragged"])
def test_impala_update_over_doubles_vs_reference(tag, scan_lib, monkeypatch):
    import machin_b200.frame.algorithms.impala as impala_mod
    from machin_b200.frame.buffers import IMPALABuffer
    from machin_b200.parallel import LocalGroup
    from test_gpu_algorithms import PolicyNet, ValueNet, _flat, _load_flat

    def buffer_double(buffer_name, group, buffer_size, requesters=None):
        return make_cpu_buffer(IMPALABuffer, buffer_size, buffer_name=buffer_name, group=group, wr_lock=RLock(),
                               requesters=[0] if requesters is None else requesters, _schema_agreed=False)

    monkeypatch.setattr(impala_mod, "IMPALABuffer", buffer_double)
    monkeypatch.setattr(impala_mod.ops, "vtrace", _emulated_vtrace(scan_lib))
    g = np.load(os.path.join(GOLDEN, "ref_vtrace.npz"))
    lengths = [int(x) for x in g[f"{tag}_lengths"]]
    clip_c, clip_rho, discount = (float(x) for x in g[f"{tag}_clip"])
    S, A = 8, 3
    actor, critic = PolicyNet(S, A), ValueNet(S)
    _load_flat(actor, g[f"{tag}_actor0"])
    _load_flat(critic, g[f"{tag}_critic0"])
    captured = {}

    def criterion(value, vs):
        captured["vs"] = vs.detach().clone()
        return ((value - vs) ** 2).sum()

    imp = impala_mod.IMPALA(actor, critic, t.optim.SGD, criterion, LocalGroup(device="cpu"), (), batch_size=len(lengths),
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
    np.testing.assert_allclose(captured["vs"].numpy().reshape(-1), g[f"{tag}_vs"], rtol=1e-5, atol=1e-6)
    assert abs(pol - ref_pol) <= 1e-5 * max(1.0, abs(ref_pol)) and abs(val - ref_val) <= 1e-5 * max(1.0, abs(ref_val))
    np.testing.assert_allclose(_flat(imp.actor), g[f"{tag}_actor1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(_flat(imp.critic), g[f"{tag}_critic1"], rtol=1e-5, atol=1e-6)


def _emulated_gae(lib):
    def gae(reward, terminal, critic_value, next_critic_value, discount, gae_lambda, offsets=None, n_episodes=None,
            max_len=0):
        p = lambda a: C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)
        r = np.ascontiguousarray(reward.detach().numpy().reshape(-1), dtype=np.float64)
        n = r.shape[0]
        term = np.ascontiguousarray(terminal.numpy().reshape(-1).astype(np.uint8))
        cv = np.ascontiguousarray(critic_value.detach().numpy().reshape(-1), dtype=np.float32)
        ncv = np.ascontiguousarray(next_critic_value.detach().numpy().reshape(-1), dtype=np.float32)
        off = None if offsets is None else np.ascontiguousarray(offsets.numpy(), dtype=np.int64)
        n_ep = int(n_episodes) if off is None else len(off) - 1
        ret, adv = np.full(n, np.nan), np.full(n, np.nan)
        assert lib.emul_gae(p(off), C.c_int64(n_ep), C.c_int64(n), p(r), p(term), p(cv), p(ncv), C.c_double(discount),
                            C.c_double(gae_lambda), p(ret), p(adv), C.c_int64(int(max_len))) == 0
        return {"returns": t.from_numpy(ret).view(reward.shape), "gae": t.from_numpy(adv).view(reward.shape)}
    return gae


@pytest.mark.parametrize("lam", [1.0, 0.0, 0.95])
def test_a2c_store_episode_and_update_over_doubles_vs_reference(lam, scan_lib, monkeypatch):
    """A2C.store_episode (batched critic + returns / GAE kernel source under emulation) against the reference's
    per-step Python loop (tests/golden/ref_gae.npz, a2c.py:269-326), then one update through the buffer double."""
    import machin_b200.frame.algorithms.a2c as a2c_mod
    from machin_b200.frame.buffers import Buffer
    from test_gpu_algorithms import PolicyNet, ValueNet, _load_flat
    monkeypatch.setattr(a2c_mod.ops, "gae", _emulated_gae(scan_lib))
    g = np.load(os.path.join(GOLDEN, "ref_gae.npz"))
    S, A = 8, 3
    actor, critic = PolicyNet(S, A), ValueNet(S)
    _load_flat(critic, g["critic0"])
    a2c = a2c_mod.A2C(actor, critic, t.optim.SGD, t.nn.MSELoss(reduction="sum"), batch_size=8, discount=0.97, gae_lambda=lam,
                      replay_buffer=make_cpu_buffer(Buffer, 5000))
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


@pytest.fixture(scope="module")
def td_lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "td_emul.so")
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread", "-shared", "-fPIC",
           "-I" + os.path.join(ROOT, "machin_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(EMUL, "td_harness.cpp"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    return C.CDLL(out)


def _emulated_td_dqn(lib):
    def td_dqn(q, q_next, q_tgt_next, action, reward, terminal, discount, is_weight=None, mode="double", want_grad=False):
        p = lambda a: C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)
        f = lambda x: None if x is None else np.ascontiguousarray(x.detach().numpy(), dtype=np.float32)
        qa, qn, qt = f(q), f(q_next), f(q_tgt_next)
        B, A = qt.shape
        act = None if action is None else np.ascontiguousarray(action.numpy().reshape(-1), dtype=np.int64)
        rew = np.ascontiguousarray(reward.detach().numpy().reshape(-1), dtype=np.float32)
        term = np.ascontiguousarray(terminal.numpy().reshape(-1).astype(np.uint8))
        isw = None if is_weight is None else np.ascontiguousarray(is_weight.numpy().reshape(-1), dtype=np.float64)
        y, av, ae = (np.full(B, np.nan, np.float32) for _ in range(3))
        le = np.full(B, np.nan)
        gq = np.full((B, A), np.nan, np.float32) if want_grad else None
        assert lib.emul_td_dqn(p(qa), p(qn), p(qt), C.c_int64(B), A, p(act), p(rew), p(term), p(isw), C.c_float(discount),
                               0 if mode == "double" else 1, p(y), p(av), p(ae), p(le), p(gq)) == 0
        return {"y": t.from_numpy(y).view(B, 1), "action_value": t.from_numpy(av).view(B, 1), "abs_error": t.from_numpy(ae),
                "loss_elem": t.from_numpy(le).view(B, 1), "grad_q": None if gq is None else t.from_numpy(gq)}
    return td_dqn


@pytest.mark.parametrize("criterion", ["mse_fused", "mse_generic"])
def test_dqn_per_update_over_doubles_vs_reference(criterion, td_lib, monkeypatch):
    """Whole DQNPer.update() -- draw (oracle restatement as the double of the draw kernel), gather, double-DQN target /
    priority math (the kernel's row function compiled for the host), PER write-back, SGD step, soft target update --
    against the unmodified reference (tests/golden/ref_td.npz; dqn_per.py:89-161)."""
    import machin_b200.frame.algorithms.dqn_per as dqn_mod
    import machin_b200.frame.algorithms.utils as utils_mod
    from fake_ring import make_cpu_prioritized
    from machin_b200.frame.buffers import PrioritizedBuffer
    from oracle import td as otd
    from test_gpu_algorithms import QNet, _flat, _load_flat

    class CpuPer(PrioritizedBuffer):
        def update_priority(self, priorities, indexes, valid=None):     # device tensors in the product: host arrays here
            to_np = lambda x: x.detach().numpy() if t.is_tensor(x) else x
            return PrioritizedBuffer.update_priority(self, to_np(priorities), to_np(indexes))

    monkeypatch.setattr(dqn_mod.ops, "td_dqn", _emulated_td_dqn(td_lib))
    soft = lambda tgt, src, rate: otd.soft_update_(list(tgt.parameters()), list(src.parameters()), rate)
    monkeypatch.setattr(dqn_mod, "soft_update", soft)
    monkeypatch.setattr(utils_mod, "soft_update", soft)
    g = np.load(os.path.join(GOLDEN, "ref_td.npz"))
    S, A, N, B = 6, 4, 200, 32
    qnet, qtgt = QNet(S, A), QNet(S, A)
    _load_flat(qtgt, g["dqn_init_qt"])
    dqn = dqn_mod.DQNPer(qnet, qtgt, t.optim.SGD, t.nn.MSELoss(reduction="sum"), batch_size=B, learning_rate=0.01,
                         discount=0.97, update_rate=0.05, replay_buffer=make_cpu_prioritized(CpuPer, N))
    if criterion == "mse_generic":
        dqn._fused_mse = False
        dqn.fuse_online_forwards = True
    assert np.array_equal(_flat(dqn.qnet), g["dqn_init_q"])
    ep = [dict(state={"state": t.from_numpy(g["dqn_state"][i:i + 1])}, action={"action": t.tensor([[g["dqn_action"][i]]])},
               next_state={"state": t.from_numpy(g["dqn_next_state"][i:i + 1])}, reward=float(g["dqn_reward"][i]),
               terminal=bool(g["dqn_terminal"][i])) for i in range(N)]
    dqn.replay_buffer.store_episode(ep, priorities=list(g["dqn_priority"]))
    assert np.array_equal(dqn.replay_buffer.wt_tree.weights, g["dqn_tree0"])
    np.random.seed(11)
    for step in range(3):
        loss = dqn.update()
        ref = float(g[f"dqn_loss{step}"])
        assert abs(loss - ref) <= 1e-5 * abs(ref), (step, loss, ref)
        np.testing.assert_allclose(dqn.replay_buffer.wt_tree.weights, g[f"dqn_tree{step + 1}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(_flat(dqn.qnet), g[f"dqn_q{step + 1}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(_flat(dqn.qnet_target), g[f"dqn_qt{step + 1}"], rtol=1e-5, atol=1e-6)


def test_ddpg_per_update_over_doubles_vs_reference(monkeypatch):
    """Whole DDPGPer.update() over the ring / tree doubles against the unmodified reference (ref_td.npz;
    ddpg_per.py:92-185): critic step on the IS-weighted loss, actor step through the updated critic, priorities from
    Q(s, mu(s)) of the updated critic, both soft updates.  The target / |error| kernel (validated on the device) is
    stood in for by its oracle restatement."""
    import machin_b200.frame.algorithms.ddpg_per as ddpg_mod
    import machin_b200.frame.algorithms.utils as utils_mod
    from fake_ring import make_cpu_prioritized
    from machin_b200.frame.buffers import PrioritizedBuffer
    from oracle import td as otd
    from test_gpu_algorithms import Actor, Critic, _flat, _load_flat

    class CpuPer(PrioritizedBuffer):
        def update_priority(self, priorities, indexes, valid=None):
            to_np = lambda x: x.detach().numpy() if t.is_tensor(x) else x
            return PrioritizedBuffer.update_priority(self, to_np(priorities), to_np(indexes))

    def td_target(next_value, reward, terminal, discount, value=None):
        y = otd.ddpg_target(next_value, reward, terminal, discount)
        return {"y": y, "abs_error": None if value is None else otd.abs_error(value, y)}

    monkeypatch.setattr(ddpg_mod.ops, "td_target", td_target)
    soft = lambda tgt, src, rate: otd.soft_update_(list(tgt.parameters()), list(src.parameters()), rate)
    monkeypatch.setattr(ddpg_mod, "soft_update", soft)
    monkeypatch.setattr(utils_mod, "soft_update", soft)
    g = np.load(os.path.join(GOLDEN, "ref_td.npz"))
    S, A, N, B = 5, 2, 150, 16
    actor, actor_t, critic, critic_t = Actor(S, A), Actor(S, A), Critic(S, A), Critic(S, A)
    _load_flat(actor_t, g["ddpg_init_actor_t"])
    _load_flat(critic_t, g["ddpg_init_critic_t"])
    ddpg = ddpg_mod.DDPGPer(actor, actor_t, critic, critic_t, t.optim.SGD, t.nn.MSELoss(reduction="sum"), batch_size=B,
                            actor_learning_rate=0.005, critic_learning_rate=0.01, discount=0.95, update_rate=0.1,
                            replay_buffer=make_cpu_prioritized(CpuPer, N))
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


@pytest.mark.parametrize("mode", ["vanilla", "fixed_target", "double"])
def test_dqn_uniform_replay_over_doubles_vs_reference(mode, td_lib, monkeypatch):
    """BASELINE config C1's shapes: DQN with the uniform Buffer (slots from Python's random stream, buffer.py:234), all
    three modes, five updates each against the unmodified reference (tests/golden/ref_dqn.npz; dqn.py:333-463)."""
    import machin_b200.frame.algorithms.dqn as dqn_mod
    import machin_b200.frame.algorithms.dqn_per as per_mod
    import machin_b200.frame.algorithms.utils as utils_mod
    from machin_b200.frame.buffers import Buffer
    from oracle import td as otd
    from test_gpu_algorithms import QNet, _flat, _load_flat
    monkeypatch.setattr(dqn_mod.ops, "td_dqn", _emulated_td_dqn(td_lib))
    soft = lambda tgt, src, rate: otd.soft_update_(list(tgt.parameters()), list(src.parameters()), rate)
    for mod in (dqn_mod, per_mod, utils_mod):
        monkeypatch.setattr(mod, "soft_update", soft)
    g = np.load(os.path.join(GOLDEN, "ref_dqn.npz"))
    S, A, N, B = 4, 2, 300, 32
    qnet, qtgt = QNet(S, A), QNet(S, A)
    _load_flat(qtgt, g[f"{mode}_init_qt"])
    if mode == "vanilla":
        _load_flat(qnet, g[f"{mode}_init_q"])           # no target net in this mode: the online net is its own target
    dqn = dqn_mod.DQN(qnet, qtgt, t.optim.SGD, t.nn.MSELoss(), batch_size=B, learning_rate=0.2, discount=0.97, update_rate=0.1,
                      replay_buffer=make_cpu_buffer(Buffer, 10000), mode=mode)
    assert np.array_equal(_flat(dqn.qnet), g[f"{mode}_init_q"]) and (dqn.qnet_target is dqn.qnet) == (mode == "vanilla")
    assert dqn.criterion.reduction == "mean"                              # the user's criterion is left alone
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
    with pytest.raises(ValueError, match="Unknown DQN mode"):
        dqn_mod.DQN(qnet, qtgt, t.optim.SGD, t.nn.MSELoss(), mode="triple", replay_buffer=make_cpu_buffer(Buffer, 10))
    with pytest.raises(RuntimeError, match="nothing to capture"):
        dqn.enable_cuda_graph()
