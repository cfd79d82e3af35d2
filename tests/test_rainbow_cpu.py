"""CPU: RAINBOW's learner math.  (1) the oracle restatement (oracle/rainbow.py) against fixtures
of the UNMODIFIED reference (tests/golden/ref_rainbow.npz from oracle/make_golden.py::gen_rainbow): n-step values,
greedy actions, projected target distribution BIT FOR BIT -- including rows whose target sits exactly on an atom and
loses its mass --, priorities and the scalar loss; (2) the DEVICE SOURCE of the fused kernel
(machin_b200/csrc/dist_device.cuh) compiled with g++ against tests/cpu_emul/cuda_shim.h, against the oracle and
the same fixtures."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch as t

from conftest import GOLDEN, ROOT
from oracle import rainbow as orb

EMUL = os.path.join(ROOT, "tests", "cpu_emul")


class DistQNet(t.nn.Module):
    def __init__(self, s, a, atoms, h=16):
        super().__init__()
        self.fc1 = t.nn.Linear(s, h)
        self.fc2 = t.nn.Linear(h, a * atoms)
        self.a, self.atoms = a, atoms

    def forward(self, state):
        return t.softmax(self.fc2(t.relu(self.fc1(state))).view(-1, self.a, self.atoms), dim=-1)


def load_flat(net, flat):
    o = 0
    with t.no_grad():
        for p in net.parameters():
            p.copy_(t.from_numpy(flat[o:o + p.numel()]).view_as(p))
            o += p.numel()


@pytest.fixture(scope="module")
def ref():
    z = np.load(os.path.join(GOLDEN, "ref_rainbow.npz"))
    S, A, N, B, NSTEP = (int(x) for x in z["config"][:5])
    cfg = dict(S=S, A=A, N=N, B=B, NSTEP=NSTEP, VMIN=float(z["config"][5]), VMAX=float(z["config"][6]),
               DISC=float(z["config"][7]), LR=float(z["config"][8]), TAU=float(z["config"][9]))
    return z, cfg


def net_outputs(z, cfg, step):
    q, qt = DistQNet(cfg["S"], cfg["A"], cfg["N"]), DistQNet(cfg["S"], cfg["A"], cfg["N"])
    load_flat(q, z["init_q"] if step == 0 else z[f"q{step}"])
    load_flat(qt, z["init_qt"] if step == 0 else z[f"qt{step}"])
    st, nst = t.from_numpy(z[f"b_state{step}"]), t.from_numpy(z[f"b_next{step}"])
    with t.no_grad():
        return q(st).numpy(), q(nst).numpy(), qt(nst).numpy()


def test_nstep_values_match_reference(ref):
    z, cfg = ref
    o, vals = 0, []
    for n in z["lens"]:
        vals += orb.nstep_values([float(x) for x in z["reward"][o:o + n]], cfg["DISC"], cfg["NSTEP"])
        o += n
    assert np.array_equal(np.array(vals), z["value"])


@pytest.mark.parametrize("step", [0, 1, 2])
def test_oracle_update_matches_reference(ref, step):
    z, cfg = ref
    B = cfg["B"]
    qd, nq, tq = net_outputs(z, cfg, step)
    sup = orb.support(cfg["VMIN"], cfg["VMAX"], cfg["N"])
    nd = tq[np.arange(B), orb.act_discrete(nq, sup)]
    td = orb.project(z[f"b_value{step}"].reshape(-1), z[f"b_terminal{step}"].reshape(-1), nd, cfg["VMIN"], cfg["VMAX"],
                     cfg["DISC"] ** cfg["NSTEP"])
    assert np.array_equal(td, z[f"target_dist{step}"])                   # bit for bit
    if step == 0:
        assert (td.sum(1) == 0).any() and (td.sum(1) < 0.95).any()      # the fixture holds dropped-mass rows
    rows = orb.loss_rows(td, qd[np.arange(B), z[f"b_action{step}"].reshape(-1)])
    np.testing.assert_allclose(orb.abs_error(rows), z[f"abs_error{step}"], rtol=1e-5)
    np.testing.assert_allclose(orb.weighted_loss(rows, z[f"isw{step}"]), float(z[f"loss{step}"]), rtol=1e-5)


def test_oracle_gradient_matches_reference_step(ref):
    """One SGD step with the oracle's d loss / d q_dist lands on the reference's updated weights."""
    z, cfg = ref
    B, A, N = cfg["B"], cfg["A"], cfg["N"]
    q, qt = DistQNet(cfg["S"], A, N), DistQNet(cfg["S"], A, N)
    load_flat(q, z["init_q"])
    load_flat(qt, z["init_qt"])
    out = q(t.from_numpy(z["b_state0"]))
    act = z["b_action0"].reshape(-1)
    grad = np.zeros((B, A, N), dtype=np.float32)
    grad[np.arange(B), act] = orb.grad_q_sel(z["target_dist0"], out.detach().numpy()[np.arange(B), act], z["isw0"])
    out.backward(t.from_numpy(grad))
    with t.no_grad():
        new = np.concatenate([(p - cfg["LR"] * p.grad).numpy().reshape(-1) for p in q.parameters()])
    np.testing.assert_allclose(new, z["q1"], rtol=1e-5, atol=1e-7)


def test_oracle_acting_matches_reference(ref):
    z, cfg = ref
    q, qt = DistQNet(cfg["S"], cfg["A"], cfg["N"]), DistQNet(cfg["S"], cfg["A"], cfg["N"])
    load_flat(q, z["q3"])
    load_flat(qt, z["qt3"])
    sup = orb.support(cfg["VMIN"], cfg["VMAX"], cfg["N"])
    with t.no_grad():
        x = t.from_numpy(z["act_state"])
        assert np.array_equal(orb.act_discrete(q(x).numpy(), sup), z["act_online"].reshape(-1))
        assert np.array_equal(orb.act_discrete(qt(x).numpy(), sup), z["act_target"].reshape(-1))


# ------------------------------------------------------------------ the kernel's device source, emulated
@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "td_emul.so")
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread", "-shared", "-fPIC",
           "-I" + os.path.join(ROOT, "machin_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(EMUL, "td_harness.cpp"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    return C.CDLL(out)


def run_kernel(lib, qd, nq, tq, action, value, terminal, isw, v_min, v_max, disc_n, scale=None):
    B, A, N = qd.shape
    c = lambda a, dt: np.ascontiguousarray(a, dtype=dt)
    qd, nq, tq = c(qd, np.float32), c(nq, np.float32), c(tq, np.float32)
    sup = c(orb.support(v_min, v_max, N), np.float32)
    action, value = c(action.reshape(-1), np.int64), c(value.reshape(-1), np.float32)
    terminal = c(terminal.reshape(-1), np.uint8)
    isw = None if isw is None else c(isw, np.float64)
    out = dict(target_dist=np.full((B, N), -7, np.float32), loss_row=np.zeros(B, np.float32), abs_error=np.zeros(B, np.float32),
               loss_elem=np.zeros(B, np.float64), grad_q=np.full((B, A, N), -7, np.float32))
    p = lambda a: C.c_void_p(0) if a is None else C.c_void_p(a.ctypes.data)
    rc = lib.emul_td_rainbow(p(qd), p(nq), p(tq), C.c_int64(B), A, N, p(sup), p(action), p(value), p(terminal), p(isw),
                             p(None if scale is None else np.array([scale], dtype=np.float64)), C.c_double(v_min), C.c_double(v_max), C.c_double(disc_n), p(out["target_dist"]),
                             p(out["loss_row"]), p(out["abs_error"]), p(out["loss_elem"]), p(out["grad_q"]))
    assert rc == 0
    return out


@pytest.mark.parametrize("step", [0, 1, 2])
def test_emulated_kernel_matches_reference_fixture(emul, ref, step):
    z, cfg = ref
    B = cfg["B"]
    qd, nq, tq = net_outputs(z, cfg, step)
    out = run_kernel(emul, qd, nq, tq, z[f"b_action{step}"], z[f"b_value{step}"], z[f"b_terminal{step}"], z[f"isw{step}"],
                     cfg["VMIN"], cfg["VMAX"], cfg["DISC"] ** cfg["NSTEP"])
    assert np.array_equal(out["target_dist"], z[f"target_dist{step}"])            # bit for bit with the reference
    np.testing.assert_allclose(out["abs_error"], z[f"abs_error{step}"], rtol=1e-5)
    np.testing.assert_allclose(out["loss_elem"].mean(), float(z[f"loss{step}"]), rtol=1e-5)
    act = z[f"b_action{step}"].reshape(-1)
    want = np.zeros_like(out["grad_q"])
    want[np.arange(B), act] = orb.grad_q_sel(z[f"target_dist{step}"], qd[np.arange(B), act], z[f"isw{step}"])
    np.testing.assert_allclose(out["grad_q"], want, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("B,A,N,v_min,v_max", [(1, 1, 2, 0.0, 1.0), (7, 2, 51, -10.0, 10.0), (33, 18, 51, -10.0, 10.0),
                                               (5, 200, 10, 0.0, 40.0), (3, 4, 256, -3.0, 5.0), (130, 3, 129, -1.0, 1.0)])
def test_emulated_kernel_matches_oracle(emul, B, A, N, v_min, v_max):
    rng = np.random.RandomState(B * 1000 + N)
    soft = lambda x: np.exp(x) / np.exp(x).sum(-1, keepdims=True)
    qd, nq, tq = (soft(rng.randn(B, A, N)).astype(np.float32) for _ in range(3))
    if A > 1:
        nq[0, 1] = nq[0, 0]                                        # a tie between two actions: the first one wins
    action = rng.randint(0, A, size=B)
    delta = (v_max - v_min) / (N - 1)
    value = np.where(rng.rand(B) < 0.4, v_min + delta * rng.randint(-2, N + 2, size=B), rng.randn(B) * (v_max - v_min))
    terminal = rng.rand(B) < 0.5
    isw = rng.rand(B) + 0.1
    disc_n = 0.99 ** 3
    out = run_kernel(emul, qd, nq, tq, action, value, terminal, isw, v_min, v_max, disc_n)
    sup = orb.support(v_min, v_max, N)
    na = orb.act_discrete(nq, sup)
    td = orb.project(value.astype(np.float32), terminal, tq[np.arange(B), na], v_min, v_max, disc_n)
    assert np.array_equal(out["target_dist"], td)
    rows = orb.loss_rows(td, qd[np.arange(B), action])
    np.testing.assert_allclose(out["loss_row"], rows, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(out["abs_error"], orb.abs_error(rows), rtol=2e-6)
    np.testing.assert_allclose(out["loss_elem"].mean(), orb.weighted_loss(rows, isw), rtol=1e-6)
    want = np.zeros((B, A, N), dtype=np.float32)
    want[np.arange(B), action] = orb.grad_q_sel(td, qd[np.arange(B), action], isw)
    np.testing.assert_allclose(out["grad_q"], want, rtol=2e-6, atol=1e-12)
    # the caller may supply sum(is_weight) / B^2 (large batches): same outputs
    out2 = run_kernel(emul, qd, nq, tq, action, value, terminal, isw, v_min, v_max, disc_n, scale=float(np.sum(isw) / (B * B)))
    np.testing.assert_allclose(out2["loss_elem"], out["loss_elem"], rtol=1e-12)
    np.testing.assert_allclose(out2["grad_q"], out["grad_q"], rtol=1e-6, atol=1e-12)
    assert np.array_equal(out2["target_dist"], out["target_dist"])
    # without IS weights: all ones
    out1 = run_kernel(emul, qd, nq, tq, action, value, terminal, None, v_min, v_max, disc_n)
    np.testing.assert_allclose(out1["loss_elem"].mean(), orb.weighted_loss(rows, np.ones(B)), rtol=1e-6)


# ------------------------------------------------------------------ the learner class over host-memory doubles
def _emulated_td_rainbow(lib):
    def td_rainbow(q_dist, next_q_dist, target_next_q_dist, action, value, terminal, v_min, v_max, discount_n,
                   is_weight=None, want_grad=False):
        out = run_kernel(lib, q_dist.detach().numpy(), next_q_dist.numpy(), target_next_q_dist.numpy(), action.numpy(),
                         value.numpy(), terminal.numpy(), None if is_weight is None else is_weight.numpy(), v_min, v_max,
                         discount_n)
        return {k: t.from_numpy(v) for k, v in out.items()}
    return td_rainbow


def test_rainbow_learner_over_doubles_vs_reference(emul, ref, monkeypatch):
    """Whole RAINBOW.store_episode / update() / act_discrete through the product's class -- ring and tree as host-memory
    doubles, the value block as the kernel's device source compiled for the host -- against the unmodified reference:
    n-step values, sampled slots, projected targets (bit for bit), losses, tree after the write-back, weights after
    the SGD step and the soft target update, greedy actions."""
    import machin_b200.frame.algorithms.dqn_per as dqn_mod
    import machin_b200.frame.algorithms.rainbow as rb_mod
    import machin_b200.frame.algorithms.utils as utils_mod
    from fake_ring import make_cpu_prioritized
    from machin_b200.frame.buffers import PrioritizedBuffer
    from oracle import td as otd

    class CpuPer(PrioritizedBuffer):
        def update_priority(self, priorities, indexes, valid=None):     # device tensors in the product: host arrays here
            to_np = lambda x: x.detach().numpy() if t.is_tensor(x) else x
            return PrioritizedBuffer.update_priority(self, to_np(priorities), to_np(indexes))

    z, cfg = ref
    monkeypatch.setattr(rb_mod.ops, "td_rainbow", _emulated_td_rainbow(emul))
    soft = lambda tgt, src, rate: otd.soft_update_(list(tgt.parameters()), list(src.parameters()), rate)
    monkeypatch.setattr(dqn_mod, "soft_update", soft)
    monkeypatch.setattr(utils_mod, "soft_update", soft)
    S, A, N, B = cfg["S"], cfg["A"], cfg["N"], cfg["B"]
    qnet, qtgt = DistQNet(S, A, N), DistQNet(S, A, N)
    load_flat(qtgt, z["init_qt"])
    rb = rb_mod.RAINBOW(qnet, qtgt, t.optim.SGD, cfg["VMIN"], cfg["VMAX"], batch_size=B, learning_rate=cfg["LR"],
                        discount=cfg["DISC"], update_rate=cfg["TAU"], reward_future_steps=cfg["NSTEP"],
                        replay_buffer=make_cpu_prioritized(CpuPer, 400))
    assert np.array_equal(np.concatenate([p.detach().numpy().reshape(-1) for p in rb.qnet.parameters()]), z["init_q"])
    o = 0
    for n in z["lens"]:
        ep = [dict(state={"state": t.from_numpy(z["state"][i:i + 1])}, action={"action": t.tensor([[int(z["action"][i])]])},
                   next_state={"state": t.from_numpy(z["next_state"][i:i + 1])}, reward=float(z["reward"][i]),
                   terminal=bool(z["terminal"][i])) for i in range(o, o + n)]
        rb.store_episode(ep)
        assert [tr["value"] for tr in ep] == list(z["value"][o:o + n])          # Python floats: bit-exact
        o += n
    assert np.array_equal(rb.replay_buffer.wt_tree.weights, z["tree0"])
    flat = lambda net: np.concatenate([p.detach().numpy().reshape(-1) for p in net.parameters()])
    np.random.seed(13)
    for step in range(3):
        loss = rb.update()
        assert np.array_equal(rb.last_target_dist.numpy(), z[f"target_dist{step}"]), step
        want = float(z[f"loss{step}"])
        assert abs(loss - want) <= 1e-5 * abs(want), (step, loss, want)
        np.testing.assert_allclose(rb.replay_buffer.wt_tree.weights, z[f"tree{step + 1}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(flat(rb.qnet), z[f"q{step + 1}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(flat(rb.qnet_target), z[f"qt{step + 1}"], rtol=1e-5, atol=1e-6)
    probe = {"state": t.from_numpy(z["act_state"])}
    assert np.array_equal(rb.act_discrete(probe).numpy(), z["act_online"])
    assert np.array_equal(rb.act_discrete(probe, use_target=True).numpy(), z["act_target"])
    # a tensor reward gives the same stored value (f32 arithmetic, as the reference would do with it)
    ep = [dict(state={"state": t.zeros(1, S)}, action={"action": t.tensor([[0]])}, next_state={"state": t.zeros(1, S)},
               reward=t.tensor([[0.5]]), terminal=i == 2) for i in range(3)]
    rb.store_episode(ep)
    assert isinstance(ep[0]["value"], float) and abs(ep[0]["value"] - (0.5 + 0.9 * (0.5 + 0.9 * 0.5))) < 1e-6
