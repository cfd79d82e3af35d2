"""GPU: the row-tile fused value math (csrc/mlp.cu: dqn_rows_kernel + single-phase dW program, DQNPer.fused_kind =
"rows") against the generic autograd path over the same FusedMLP nets, and against the reference fixture."""
import os

import numpy as np
import pytest
import torch as t

from conftest import GOLDEN

pytestmark = [pytest.mark.gpu]


def _flat(m):
    return t.cat([p.detach().reshape(-1) for p in m.parameters()]).cpu().numpy()


def _load_flat(m, flat):
    o = 0
    with t.no_grad():
        for p in m.parameters():
            n = p.numel()
            p.copy_(t.from_numpy(flat[o:o + n]).view_as(p))
            o += n


@pytest.mark.parametrize("sizes,batch", [([6, 16, 16, 4], 32), ([84, 256, 256, 6], 512), ([7, 33, 10, 3], 50),
                                          ([12, 64, 5], 9), ([20, 8, 8, 8, 2], 64)])
@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("cpt", ["4", "2"])
def test_rows_program_matches_autograd_path(sizes, batch, graph, cpt, monkeypatch):
    monkeypatch.setenv("MB_ROWS_CPT", cpt)              # thread layout of the row-tile kernel, read at program creation
    from machin_b200.frame.algorithms import DQNPer
    from machin_b200.model import FusedMLP
    S, A, N = sizes[0], sizes[-1], 3000

    def make(rows):
        t.manual_seed(1)
        q, qt = FusedMLP(sizes).cuda(), FusedMLP(sizes).cuda()
        learner = DQNPer(q, qt, t.optim.SGD, t.nn.MSELoss(reduction="none"), batch_size=batch, learning_rate=0.05,
                         discount=0.9, update_rate=0.1, replay_size=4096)
        rng = np.random.RandomState(3)
        learner.replay_buffer.store_episode_tensors(
            {"state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))},
             "action": {"action": t.from_numpy(rng.randint(0, A, size=(N, 1)))},
             "next_state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))}},
            {"reward": t.from_numpy(rng.randn(N, 1).astype(np.float32)), "terminal": t.from_numpy(rng.rand(N, 1) < 0.1)},
            priorities=rng.exponential(size=N))
        if rows:
            learner.use_fused_program, learner.fused_kind = True, "rows"
        if graph:
            learner.enable_cuda_graph()
        return learner

    a, b = make(False), make(True)
    for step in range(4):
        np.random.seed(70 + step)
        la = a.update_device().item()
        np.random.seed(70 + step)
        lb = b.update_device().item()
        assert b._program, "row-tile program was not built"
        assert abs(la - lb) <= 1e-5 * max(1e-3, abs(la)), (step, la, lb)
        np.testing.assert_allclose(b.replay_buffer.wt_tree.weights, a.replay_buffer.wt_tree.weights, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(_flat(b.qnet), _flat(a.qnet), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(_flat(b.qnet_target), _flat(a.qnet_target), rtol=1e-5, atol=1e-6)


def test_rows_program_vs_reference_fixture():
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
    dqn.use_fused_program, dqn.fused_kind = True, "rows"
    dqn.enable_cuda_graph()
    assert dqn._program
    np.random.seed(11)
    for step in range(3):
        loss = dqn.update()
        ref = float(g[f"dqn_loss{step}"])
        assert abs(loss - ref) <= 1e-5 * abs(ref), (step, loss, ref)
        np.testing.assert_allclose(dqn.replay_buffer.wt_tree.weights, g[f"dqn_tree{step + 1}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(_flat(dqn.qnet), g[f"dqn_q{step + 1}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(_flat(dqn.qnet_target), g[f"dqn_qt{step + 1}"], rtol=1e-5, atol=1e-6)


def test_rows_program_rejects_wide_layers():
    from machin_b200.frame.algorithms import DQNPer
    from machin_b200.model import FusedMLP
    q, qt = FusedMLP([8, 512, 4]).cuda(), FusedMLP([8, 512, 4]).cuda()
    dqn = DQNPer(q, qt, t.optim.SGD, t.nn.MSELoss(reduction="none"), batch_size=16, replay_size=64)
    dqn.replay_buffer.store_episode_tensors(
        {"state": {"state": t.randn(32, 8)}, "action": {"action": t.zeros(32, 1, dtype=t.int64)}, "next_state": {"state": t.randn(32, 8)}},
        {"reward": t.randn(32, 1), "terminal": t.zeros(32, 1, dtype=t.bool)})
    dqn.use_fused_program, dqn.fused_kind = True, "rows"
    with pytest.raises(RuntimeError, match="libmachin_b200"):
        dqn.update()


@pytest.mark.parametrize("kind", ["rows", "program"])
def test_fused_program_inside_fixed_shape_exchange_single_member(kind):
    """DQNApex's fixed-shape update (group of one) with the fused value-math program over the merged, padded batch
    against the same update through autograd."""
    from machin_b200.frame.algorithms import DQNApex
    from machin_b200.model import FusedMLP
    from machin_b200.parallel import LocalGroup
    sizes, N, B = [12, 32, 32, 5], 3000, 64
    S, A = sizes[0], sizes[-1]

    def make(fused):
        t.manual_seed(0)
        q, qt = FusedMLP(sizes).cuda(), FusedMLP(sizes).cuda()
        learner = DQNApex(q, qt, t.optim.SGD, t.nn.MSELoss(reduction="none"), LocalGroup(device="cuda:0"), (), batch_size=B,
                          learning_rate=0.05, discount=0.9, update_rate=0.1, replay_size=4096)
        rng = np.random.RandomState(3)
        learner.replay_buffer.store_episode_tensors(
            {"state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))},
             "action": {"action": t.from_numpy(rng.randint(0, A, size=(N, 1)))},
             "next_state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))}},
            {"reward": t.from_numpy(rng.randn(N, 1).astype(np.float32)), "terminal": t.from_numpy(rng.rand(N, 1) < 0.1)},
            priorities=rng.exponential(size=N))
        if fused:
            learner.use_fused_program, learner.fused_kind = True, kind
        learner.enable_cuda_graph(static_exchange=True)
        return learner

    a, b = make(False), make(True)
    for step in range(5):
        np.random.seed(100 + step)
        la = a.update_device().item()
        np.random.seed(100 + step)
        lb = b.update_device().item()
        assert abs(la - lb) <= 1e-5 * max(1e-3, abs(la)), (step, la, lb)
        np.testing.assert_allclose(b.replay_buffer.wt_tree.weights, a.replay_buffer.wt_tree.weights, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(_flat(b.qnet), _flat(a.qnet), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("graph", [False, True])
def test_rows_program_with_overlapped_priority_update(graph):
    """overlap_priority_update in the fused step: the write-back starts after part 0 (forward, TD, backward-input) on a
    second stream, beside part 1 (dW), the optimizer and the target update -- same trajectory."""
    from machin_b200.frame.algorithms import DQNPer
    from machin_b200.model import FusedAdam, FusedMLP
    sizes, N, B = [12, 32, 32, 5], 3000, 64
    S, A = sizes[0], sizes[-1]

    def make(overlap):
        t.manual_seed(1)
        q, qt = FusedMLP(sizes).cuda(), FusedMLP(sizes).cuda()
        learner = DQNPer(q, qt, FusedAdam, t.nn.MSELoss(reduction="none"), batch_size=B, learning_rate=1e-3,
                         discount=0.9, update_rate=0.1, replay_size=4096)
        rng = np.random.RandomState(3)
        learner.replay_buffer.store_episode_tensors(
            {"state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))},
             "action": {"action": t.from_numpy(rng.randint(0, A, size=(N, 1)))},
             "next_state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))}},
            {"reward": t.from_numpy(rng.randn(N, 1).astype(np.float32)), "terminal": t.from_numpy(rng.rand(N, 1) < 0.1)},
            priorities=rng.exponential(size=N))
        learner.use_fused_program, learner.fused_kind, learner.overlap_priority_update = True, "rows", overlap
        if graph:
            learner.enable_cuda_graph()
        return learner

    a, b = make(False), make(True)
    for step in range(6):
        np.random.seed(70 + step)
        la = a.update_device().item()
        np.random.seed(70 + step)
        lb = b.update_device().item()
        assert la == lb, (step, la, lb)                    # the same kernels on the same inputs
        assert np.array_equal(a.replay_buffer.wt_tree.weights, b.replay_buffer.wt_tree.weights)
        assert np.array_equal(_flat(a.qnet), _flat(b.qnet))
    b.replay_buffer.wt_tree.check()
