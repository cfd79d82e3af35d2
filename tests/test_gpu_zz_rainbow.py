"""GPU: RAINBOW.  The fused categorical value block (csrc/dist_device.cuh through ops.td_rainbow)
against the oracle and against fixtures of the unmodified reference -- projected target distribution bit for bit --,
then whole learner updates (eager and as one CUDA graph) against the reference's trajectory
(tests/golden/ref_rainbow.npz).  The CPU twin of this file is tests/test_rainbow_cpu.py."""
import numpy as np
import pytest
import torch as t


from oracle import rainbow as orb
from test_rainbow_cpu import DistQNet, load_flat, net_outputs, ref  # noqa: F401  (ref: fixture)

pytestmark = [pytest.mark.gpu]


def run_device(qd, nq, tq, action, value, terminal, isw, v_min, v_max, disc_n):
    from machin_b200 import ops
    dev = lambda a, dt: t.from_numpy(np.ascontiguousarray(a)).to(device="cuda", dtype=dt)
    out = ops.td_rainbow(dev(qd, t.float32), dev(nq, t.float32), dev(tq, t.float32), dev(action.reshape(-1, 1), t.int64),
                         dev(value.reshape(-1, 1), t.float32), dev(terminal.reshape(-1, 1), t.bool), v_min, v_max, disc_n,
                         is_weight=None if isw is None else dev(isw, t.float64), want_grad=True)
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("step", [0, 1, 2])
def test_value_block_vs_reference_fixture(ref, step):  # noqa: F811
    z, cfg = ref
    B = cfg["B"]
    qd, nq, tq = net_outputs(z, cfg, step)                       # network outputs computed on the CPU, like the reference
    out = run_device(qd, nq, tq, z[f"b_action{step}"], z[f"b_value{step}"], z[f"b_terminal{step}"], z[f"isw{step}"],
                     cfg["VMIN"], cfg["VMAX"], cfg["DISC"] ** cfg["NSTEP"])
    assert np.array_equal(out["target_dist"], z[f"target_dist{step}"])          # bit for bit with the reference
    np.testing.assert_allclose(out["abs_error"], z[f"abs_error{step}"], rtol=1e-5)
    np.testing.assert_allclose(out["loss_elem"].mean(), float(z[f"loss{step}"]), rtol=1e-5)
    act = z[f"b_action{step}"].reshape(-1)
    want = np.zeros_like(out["grad_q"])
    want[np.arange(B), act] = orb.grad_q_sel(z[f"target_dist{step}"], qd[np.arange(B), act], z[f"isw{step}"])
    np.testing.assert_allclose(out["grad_q"], want, rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("B,A,N,v_min,v_max", [(1, 1, 2, 0.0, 1.0), (7, 2, 51, -10.0, 10.0), (33, 18, 51, -10.0, 10.0),
                                               (5, 200, 10, 0.0, 40.0), (3, 4, 256, -3.0, 5.0), (512, 6, 51, -10.0, 10.0),
                                               (4100, 2, 10, 0.0, 40.0)])   # > 4096 rows: sum(is_weight) reduced once up front
def test_value_block_vs_oracle(B, A, N, v_min, v_max):
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
    out = run_device(qd, nq, tq, action, value, terminal, isw, v_min, v_max, disc_n)
    sup = orb.support(v_min, v_max, N)
    td = orb.project(value.astype(np.float32), terminal, tq[np.arange(B), orb.act_discrete(nq, sup)], v_min, v_max, disc_n)
    assert np.array_equal(out["target_dist"], td)
    rows = orb.loss_rows(td, qd[np.arange(B), action])
    np.testing.assert_allclose(out["loss_row"], rows, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["abs_error"], orb.abs_error(rows), rtol=1e-5)
    np.testing.assert_allclose(out["loss_elem"].mean(), orb.weighted_loss(rows, isw), rtol=1e-5)
    want = np.zeros((B, A, N), dtype=np.float32)
    want[np.arange(B), action] = orb.grad_q_sel(td, qd[np.arange(B), action], isw)
    np.testing.assert_allclose(out["grad_q"], want, rtol=1e-5, atol=1e-10)


def test_value_block_rejects_bad_shapes():
    from machin_b200 import ops
    z = lambda *s: t.zeros(*s, device="cuda")
    with pytest.raises(RuntimeError):                                # more atoms than the kernel's shared arrays hold
        ops.td_rainbow(z(2, 3, 300), z(2, 3, 300), z(2, 3, 300), z(2, 1).long(), z(2, 1), z(2, 1).bool(), 0.0, 1.0, 0.9)
    with pytest.raises(ValueError):
        ops.td_rainbow(z(2, 3, 11), z(2, 4, 11), z(2, 3, 11), z(2, 1).long(), z(2, 1), z(2, 1).bool(), 0.0, 1.0, 0.9)


@pytest.mark.parametrize("graph", [False, True])
def test_rainbow_learner_vs_reference(ref, graph):  # noqa: F811
    from machin_b200.frame.algorithms import RAINBOW
    z, cfg = ref
    S, A, N, B = cfg["S"], cfg["A"], cfg["N"], cfg["B"]
    qnet, qtgt = DistQNet(S, A, N).cuda(), DistQNet(S, A, N).cuda()
    load_flat(qtgt, z["init_qt"])
    rb = RAINBOW(qnet, qtgt, t.optim.SGD, cfg["VMIN"], cfg["VMAX"], batch_size=B, learning_rate=cfg["LR"],
                 discount=cfg["DISC"], update_rate=cfg["TAU"], reward_future_steps=cfg["NSTEP"], replay_size=400)
    flat = lambda net: np.concatenate([p.detach().cpu().numpy().reshape(-1) for p in net.parameters()])
    assert np.array_equal(flat(rb.qnet), z["init_q"])
    o = 0
    for n in z["lens"]:
        ep = [dict(state={"state": t.from_numpy(z["state"][i:i + 1])}, action={"action": t.tensor([[int(z["action"][i])]])},
                   next_state={"state": t.from_numpy(z["next_state"][i:i + 1])}, reward=float(z["reward"][i]),
                   terminal=bool(z["terminal"][i])) for i in range(o, o + n)]
        rb.store_episode(ep)
        assert [tr["value"] for tr in ep] == list(z["value"][o:o + n])
        o += n
    assert np.array_equal(rb.replay_buffer.wt_tree.weights, z["tree0"])
    if graph:
        rb.enable_cuda_graph()
        assert np.array_equal(rb.replay_buffer.wt_tree.weights, z["tree0"]) and np.array_equal(flat(rb.qnet), z["init_q"])
    np.random.seed(13)
    for step in range(3):
        loss = rb.update()
        want = float(z[f"loss{step}"])
        assert abs(loss - want) <= 1e-5 * abs(want), (step, loss, want)
        # the network runs on the GPU here (cuBLAS float32): targets agree to rounding, dropped-mass rows exactly
        got = rb.last_target_dist.cpu().numpy()
        np.testing.assert_allclose(got, z[f"target_dist{step}"], rtol=1e-4, atol=1e-6)
        assert np.array_equal(got.sum(1) == 0, z[f"target_dist{step}"].sum(1) == 0)
        np.testing.assert_allclose(rb.replay_buffer.wt_tree.weights, z[f"tree{step + 1}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(flat(rb.qnet), z[f"q{step + 1}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(flat(rb.qnet_target), z[f"qt{step + 1}"], rtol=1e-5, atol=1e-6)
    probe = {"state": t.from_numpy(z["act_state"]).cuda()}
    assert np.array_equal(rb.act_discrete(probe).cpu().numpy(), z["act_online"])
    assert np.array_equal(rb.act_discrete(probe, use_target=True).cpu().numpy(), z["act_target"])
