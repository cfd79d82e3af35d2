"""GPU parity: row gather/scatter, TD target math, V-trace and GAE kernels against the oracle and the
reference-generated fixtures."""
import os

import numpy as np
import pytest
import torch as t

from conftest import GOLDEN
from oracle import gae as ogae
from oracle import td as otd
from oracle import vtrace as ovt

pytestmark = pytest.mark.gpu


def _ops():
    from machin_b200 import ops
    return ops


# ----------------------------------------------------------------------------- rows
@pytest.mark.parametrize("state_dim,cap,n", [(84, 1000, 512), (5, 64, 200), (376, 333, 1000)])
def test_ring_store_and_gather(state_dim, cap, n):
    from machin_b200.frame.buffers import Buffer
    rng = np.random.RandomState(0)
    buf = Buffer(buffer_size=cap)
    mirror = {}
    pos_all = []
    stored = 0
    while stored < n:
        m = int(rng.randint(1, 40))
        ep = []
        for _ in range(m):
            ep.append(dict(state={"state": t.from_numpy(rng.randn(1, state_dim).astype(np.float32))},
                           action={"action": t.tensor([[int(rng.randint(0, 6))]])},
                           next_state={"state": t.from_numpy(rng.randn(1, state_dim).astype(np.float32))},
                           reward=float(np.float32(rng.randn())), terminal=bool(rng.rand() < 0.3),
                           tag=int(stored + len(ep))))
        pos = buf.store_episode(ep)
        for p, tr in zip(pos, ep):
            mirror[int(p)] = tr
        pos_all.append(pos)
        stored += m
    assert buf.size() == min(cap, stored)
    live = buf._sampleable_handles()
    assert set(live) <= set(mirror)
    import random
    random.seed(0)
    bs, (state, action, reward, next_state, terminal, others) = buf.sample_batch(
        64, sample_attrs=["state", "action", "reward", "next_state", "terminal", "*"])
    random.seed(0)
    picks = random.sample(live, k=min(64, len(live)))
    assert bs == len(picks)
    assert state["state"].shape == (bs, state_dim) and action["action"].dtype == t.int64
    assert reward.shape == (bs, 1) and reward.dtype == t.float32 and terminal.dtype == t.bool
    for i, p in enumerate(picks):
        tr = mirror[p]
        assert t.equal(state["state"][i].cpu(), tr["state"]["state"][0])
        assert t.equal(next_state["state"][i].cpu(), tr["next_state"]["state"][0])
        assert int(action["action"][i]) == int(tr["action"]["action"])
        assert float(reward[i]) == tr["reward"] and bool(terminal[i]) == tr["terminal"]
        assert others["tag"][i] == tr["tag"]


# ----------------------------------------------------------------------------- TD
@pytest.mark.parametrize("B,A", [(1, 2), (32, 4), (512, 6), (4096, 18)])
def test_td_dqn_bit_exact(B, A):
    ops = _ops()
    g = t.Generator().manual_seed(B)
    q, qn, qt = (t.randn(B, A, generator=g) for _ in range(3))
    qn[::3, 1] = qn[::3, 0]  # ties -> first maximum
    action = t.randint(0, A, (B, 1), generator=g)
    reward = t.randn(B, 1, generator=g)
    terminal = t.rand(B, 1, generator=g) < 0.3
    isw = t.rand(B, generator=g, dtype=t.float64)
    av, y, abs_err = otd.dqn_per_target(q, qn, qt, action, reward, terminal, 0.99)
    out = ops.td_dqn(q.cuda(), qn.cuda(), qt.cuda(), action.cuda(), reward.cuda(), terminal.cuda(), 0.99,
                     is_weight=isw.cuda(), want_grad=True)
    assert t.equal(out["y"].cpu(), y) and t.equal(out["action_value"].cpu(), av)
    assert t.equal(out["abs_error"].cpu(), abs_err)
    loss = otd.per_weighted_loss((av - y) ** 2, isw)
    np.testing.assert_allclose(out["loss_elem"].mean().item(), loss.item(), rtol=1e-12)
    qg = q.clone().requires_grad_()
    otd.per_weighted_loss((qg.gather(1, action) - y) ** 2, isw).backward()
    np.testing.assert_allclose(out["grad_q"].cpu().numpy(), qg.grad.numpy(), rtol=1e-6, atol=1e-9)


def test_td_target_and_soft_update_bit_exact():
    ops = _ops()
    g = t.Generator().manual_seed(1)
    B = 1000
    nv, r, v = t.randn(B, 1, generator=g), t.randn(B, 1, generator=g), t.randn(B, 1, generator=g)
    term = t.rand(B, 1, generator=g) < 0.5
    y = otd.ddpg_target(nv, r, term, 0.95)
    out = ops.td_target(nv.cuda(), r.cuda(), term.cuda(), 0.95, value=v.cuda())
    assert t.equal(out["y"].cpu(), y)
    assert t.equal(out["abs_error"].cpu(), otd.abs_error(v, y))
    tgt, src = t.randn(10007, generator=g), t.randn(10007, generator=g)
    want = [tgt.clone()]
    otd.soft_update_(want, [src], 0.005)
    d_tgt = tgt.cuda()
    ops.soft_update_flat(d_tgt, src.cuda(), 0.005)
    assert t.equal(d_tgt.cpu(), want[0])


# ----------------------------------------------------------------------------- V-trace
def _vtrace_case(lengths, seed, clip_c=1.0, clip_rho=1.0, discount=0.99):
    g = t.Generator().manual_seed(seed)
    n = int(sum(lengths))
    value, next_value, reward = (t.randn(n, 1, generator=g) for _ in range(3))
    logp = -t.rand(n, 1, generator=g) * 2
    logmu = -t.rand(n, 1, generator=g) * 2
    term = t.zeros(n, 1, dtype=t.bool)
    ends = np.cumsum(lengths) - 1
    term[ends[::2]] = True
    c, rho, delta = ovt.vtrace_prepare(logp, logmu, value, next_value, reward, term, discount, clip_c, clip_rho)
    return dict(value=value, next_value=next_value, reward=reward, logp=logp, logmu=logmu, term=term,
                c=c, rho=rho, delta=delta, lengths=list(lengths), discount=discount, clip=(clip_c, clip_rho))


def _run_vtrace(case, offsets=True):
    ops = _ops()
    off = None
    if offsets:
        off = t.tensor(np.concatenate(([0], np.cumsum(case["lengths"]))), dtype=t.int64).cuda()
    return ops.vtrace(case["value"].cuda(), case["next_value"].cuda(), case["reward"].cuda(), case["term"].cuda(),
                      case["logp"].cuda(), case["logmu"].cuda(), case["discount"], case["clip"][0], case["clip"][1],
                      offsets=off, n_episodes=len(case["lengths"]))


@pytest.mark.parametrize("lengths,seed", [([20] * 64, 0), ([1, 7, 2, 33, 5, 1, 12, 40, 3], 1), ([1], 2), ([2], 3),
                                          ([5000, 3, 4000], 4), (list(np.random.RandomState(5).randint(1, 41, 700)), 5)])
def test_vtrace_vs_reference_loop(lengths, seed):
    case = _vtrace_case(lengths, seed, clip_c=0.9, clip_rho=1.2)
    vs, vss = ovt.vtrace_loop(case["value"], case["delta"], case["c"], case["lengths"], case["discount"])
    out = _run_vtrace(case)
    # exp() is the only operation whose rounding may differ (CUDA expf <= 2 ulp): tolerance 1e-5 rel
    np.testing.assert_allclose(out["rho"].cpu().numpy(), case["rho"].numpy(), rtol=1e-6)
    np.testing.assert_allclose(out["vs"].cpu().numpy(), vs.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["vss"].cpu().numpy(), vss.numpy(), rtol=1e-5, atol=1e-6)
    adv = case["reward"] + case["discount"] * vss - case["value"]
    np.testing.assert_allclose(out["adv"].cpu().numpy(), adv.numpy(), rtol=1e-5, atol=1e-6)
    ends = np.cumsum(lengths) - 1
    assert np.all(out["vs"].cpu().numpy()[ends] == 0) and np.all(out["vss"].cpu().numpy()[ends] == 0)


def test_vtrace_fixed_length_full_config_c3():
    """BASELINE config C3: 1024 episodes x T=20; fixed-length path == CSR path bit-for-bit; matches the
    vectorised oracle (itself pinned to the reference loop)."""
    case = _vtrace_case([20] * 1024, 9)
    a, b = _run_vtrace(case, offsets=False), _run_vtrace(case, offsets=True)
    for k in ("vs", "vss", "rho", "adv"):
        assert t.equal(a[k], b[k])
    vs, vss = ovt.vtrace_batched(case["value"], case["delta"], case["c"], case["lengths"], case["discount"])
    np.testing.assert_allclose(a["vs"].cpu().numpy(), vs.numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("tag", ["fixed", "ragged"])
def test_vtrace_vs_reference_fixture(tag):
    from test_oracle_golden import _vtrace_inputs
    g = np.load(os.path.join(GOLDEN, "ref_vtrace.npz"))
    clip_c, clip_rho, discount = (float(x) for x in g[f"{tag}_clip"])
    ls, logp, logmu, value, next_value, rew, term = _vtrace_inputs(g, tag)
    case = dict(value=value, next_value=next_value, reward=rew, term=term, logp=logp, logmu=logmu, lengths=ls,
                discount=discount, clip=(clip_c, clip_rho))
    out = _run_vtrace(case)
    np.testing.assert_allclose(out["vs"].cpu().numpy().reshape(-1), g[f"{tag}_vs"], rtol=1e-5, atol=1e-6)


# ----------------------------------------------------------------------------- GAE
@pytest.mark.parametrize("lam", [1.0, 0.0, 0.95])
def test_gae_vs_reference_fixture(lam):
    ops = _ops()
    g = np.load(os.path.join(GOLDEN, "ref_gae.npz"))
    rew, term, cv, ncv, lens = [], [], [], [], []
    for e in range(4):
        k = f"lam{lam}_ep{e}_"
        rew.append(g[k + "reward"].astype(np.float64)); term.append(g[k + "terminal"])
        cv.append(g[k + "critic_v"]); ncv.append(g[k + "critic_nv"]); lens.append(len(g[k + "reward"]))
    off = t.tensor(np.concatenate(([0], np.cumsum(lens))), dtype=t.int64).cuda()
    out = ops.gae(t.from_numpy(np.concatenate(rew)).cuda(), t.from_numpy(np.concatenate(term)).cuda(),
                  t.from_numpy(np.concatenate(cv)).cuda(), t.from_numpy(np.concatenate(ncv)).cuda(), 0.97, lam,
                  offsets=off, n_episodes=4)
    ret, adv = out["returns"].cpu().numpy(), out["gae"].cpu().numpy()
    pos = 0
    for e in range(4):
        k = f"lam{lam}_ep{e}_"
        assert np.array_equal(ret[pos:pos + lens[e]], g[k + "value"])           # float64, bit-exact
        want = ogae.gae([float(x) for x in rew[e]], [bool(x) for x in term[e]], [float(x) for x in cv[e]],
                        [float(x) for x in ncv[e]], list(g[k + "value"]), 0.97, lam)
        assert np.array_equal(adv[pos:pos + lens[e]], np.array(want))             # bit-exact vs oracle
        np.testing.assert_allclose(adv[pos:pos + lens[e]], g[k + "gae"], rtol=1e-5, atol=2e-6)
        pos += lens[e]
