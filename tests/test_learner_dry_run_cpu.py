"""CPU: a DRY RUN of DQNApex's fixed-shape update (frame/algorithms/apex.py: _enable_static_graph, its body(), the
host preparation and update_device) -- every Python line executes, CUDA streams / graphs / pinned memory are no-op
stand-ins, the ring and tree are host-memory doubles, and the library is a recorder that converts every argument
the way ctypes would and checks it against the declared signature.  No numbers are checked (the kernels do not
run): this catches what a GPU box would otherwise be the first to see -- a wrong argument count or type, a
missing attribute, a shape error -- in the default N > 1 path and its opt-in variants (device apportioning, one-launch
draw, fused value-math program)."""
import contextlib
import ctypes as C
from threading import RLock

import numpy as np
import pytest
import torch as t

from fake_ring import FakeRing, FakeTree, make_cpu_buffer
from test_library_call_sites_cpu import _Recorder


def _patch_library(monkeypatch, rec):
    """Every module of the package that holds a reference to the ctypes library gets the recorder instead."""
    import importlib
    import pkgutil
    import sys
    import machin_b200
    import machin_b200._lib as lib_mod
    real = lib_mod.lib
    for info in pkgutil.walk_packages(machin_b200.__path__, "machin_b200."):
        if not info.name.endswith("libmachin_b200"):              # (the shared library sits in the package directory)
            importlib.import_module(info.name)
    for name, mod in list(sys.modules.items()):
        if name.startswith("machin_b200") and getattr(mod, "lib", None) is real:
            monkeypatch.setattr(mod, "lib", rec)


class _Stream:
    def wait_stream(self, other):
        pass


class _Event:
    def synchronize(self):
        pass

    def record(self):
        pass


class _Graph:
    released = 0

    def replay(self):
        pass

    def reset(self):
        type(self).released += 1


class _Tree(FakeTree):
    handle = C.c_void_p(1)

    def snapshot(self):
        return self.ora.weights.copy(), self.ora.max_leaf

    def restore(self, snap):
        self.ora.weights[:] = snap[0]
        self.ora.max_leaf = snap[1]


class _Ring(FakeRing):
    def gather_into_wire(self, index, wire, stream=None):
        wire[:, :self.pitch] = self.data[index.clamp(0, self.max_size - 1)]


@pytest.mark.parametrize("variant", ["default", "device_apportion+batched_draw", "fused_rows", "fused_program", "dynamic"])
def test_fixed_shape_update_dry_run(variant, monkeypatch):
    import machin_b200._lib as lib_mod
    import machin_b200.frame.algorithms.apex as apex_mod
    import machin_b200.frame.algorithms.dqn_per as dqn_mod
    import machin_b200.model as model_mod
    import machin_b200.ops as ops
    from machin_b200.frame.buffers import DistributedPrioritizedBuffer
    from machin_b200.parallel import LocalGroup
    rec = _Recorder()
    _patch_library(monkeypatch, rec)
    monkeypatch.setattr(lib_mod, "stream_ptr", lambda stream=None: C.c_void_p(0))
    monkeypatch.setattr(ops, "_c", lambda x, dtype=None: None if x is None else (x if dtype is None else x.to(dtype)).contiguous())
    monkeypatch.setattr(t.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(t, "empty", t.zeros)              # what a kernel would have filled in reads as zeros, not garbage
    monkeypatch.setattr(t.Tensor, "is_cuda", property(lambda self: True))          # "device" tensors of the dry run
    monkeypatch.setattr(t.cuda, "Event", _Event)
    monkeypatch.setattr(t.cuda, "Stream", lambda device=None: _Stream())
    monkeypatch.setattr(t.cuda, "current_stream", lambda device=None: _Stream())
    monkeypatch.setattr(t.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(t.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(t.cuda, "synchronize", lambda device=None: None)
    monkeypatch.setattr(t.cuda, "CUDAGraph", _Graph)
    monkeypatch.setattr(t.cuda, "graph", lambda g, **kw: contextlib.nullcontext())
    dev = t.device("cpu", 0)

    def buffer_double(buffer_name, group, buffer_size, requesters=None, **kw):
        buf = make_cpu_buffer(DistributedPrioritizedBuffer, buffer_size, epsilon=1e-2, alpha=0.6, beta=0.4, curr_beta=0.4,
                              beta_increment_per_sampling=1e-3, wt_tree=_Tree(buffer_size), buffer_name=buffer_name, group=group,
                              wr_lock=RLock(), _requesters=requesters, _sampler=None,
                              buffer_version_table=t.zeros(buffer_size, dtype=t.int64))
        buf.storage = _Ring(buffer_size)
        buf.buffer_device = dev
        buf.weight_sum_tensor = lambda: t.tensor([buf.wt_tree.get_weight_sum()], dtype=t.float64)
        return buf

    monkeypatch.setattr(apex_mod, "DistributedPrioritizedBuffer", buffer_double)
    S, A, B, N = 6, 4, 16, 200
    fused = variant.startswith("fused")
    Net = (lambda: model_mod.FusedMLP([S, 16, 16, A])) if fused else (lambda: t.nn.Sequential(t.nn.Linear(S, 16), t.nn.ReLU(), t.nn.Linear(16, A)))

    class Q(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.net = Net()

        def forward(self, state):
            return self.net(state)

    q, qt = (Net(), Net()) if fused else (Q(), Q())
    if fused:                                   # FusedMLP.forward launches kernels: the dry run computes it with torch
        monkeypatch.setattr(model_mod.FusedMLP, "forward", lambda self, state, action=None: t.nn.functional.linear(
            t.relu(t.nn.functional.linear(t.relu(t.nn.functional.linear(state, self.layers[0].weight, self.layers[0].bias)),
                                          self.layers[1].weight, self.layers[1].bias)), self.layers[2].weight, self.layers[2].bias))
    group = LocalGroup(device="cpu")
    group.device = dev
    learner = apex_mod.DQNApex(q, qt, model_mod.FusedAdam if fused else t.optim.SGD, t.nn.MSELoss(reduction="none"), group, (),
                               batch_size=B, learning_rate=0.01, discount=0.9, update_rate=0.1, replay_size=N)
    learner.replay_buffer.store_episode(
        [dict(state={"state": t.randn(1, S)}, action={"action": t.tensor([[i % A]])}, next_state={"state": t.randn(1, S)},
              reward=float(i), terminal=False) for i in range(50)], priorities=list(np.random.RandomState(0).exponential(size=50)))
    if variant == "device_apportion+batched_draw":
        learner.batched_draw = True
    if fused:
        learner.use_fused_program, learner.fused_kind = True, ("rows" if variant == "fused_rows" else "program")
    if variant == "dynamic":                    # exactly-sized eager exchanges between two captured graphs
        learner.enable_cuda_graph()
    else:
        learner.enable_cuda_graph(static_exchange=True, device_apportion=variant.startswith("device_apportion"))
    for _ in range(3):
        loss = learner.update_device()
        assert loss.shape == ()
    before, n_calls = _Graph.released, len(rec.calls)
    learner.release_graphs()                    # afterwards: the eager exchange, launch by launch
    assert _Graph.released - before == (2 if variant == "dynamic" else 1)
    assert learner.update_device().shape == () and len(rec.calls) > n_calls
    called = {name for name, _ in rec.calls}
    want = {"mb_rows_pack_meta", "mb_tree_update_priority"}
    if variant == "dynamic":
        assert {"mb_rows_unpack_meta", "mb_per_sample", "mb_td_dqn"} | want <= called
        return
    want |= {"mb_wire_compact", "mb_wire_scatter_update", "mb_wire_update_mask"}
    want |= {"mb_per_sample_multi", "mb_wire_apportion"} if variant.startswith("device") else {"mb_per_sample"}
    want |= {"mb_dqn_mlp_rows_create" if variant == "fused_rows" else "mb_dqn_mlp_program_create", "mb_mlp_program_run",
             "mb_adam_step_multi"} if fused else {"mb_td_dqn"}
    assert want <= called, sorted(want - called)
    assert learner.graph_launches >= 0


@pytest.mark.parametrize("flags", [(), ("overlap_priority_update",), ("parallel_forwards",), ("fuse_optimizer_and_target",),
                                   ("fused_rows",), ("fused_rows", "overlap_priority_update", "fuse_optimizer_and_target"),
                                   ("fused_program",)])
@pytest.mark.parametrize("graph", [False, True])
def test_single_learner_update_dry_run(flags, graph, monkeypatch):
    """The same for DQNPer's captured single-GPU update and its stream-level / fusion options."""
    import machin_b200._lib as lib_mod
    import machin_b200.frame.algorithms.dqn_per as dqn_mod
    import machin_b200.frame.buffers.prioritized_buffer as per_mod
    import machin_b200.model as model_mod
    import machin_b200.ops as ops
    from machin_b200.frame.buffers import PrioritizedBuffer
    rec = _Recorder()
    _patch_library(monkeypatch, rec)
    monkeypatch.setattr(lib_mod, "stream_ptr", lambda stream=None: C.c_void_p(0))
    monkeypatch.setattr(ops, "_c", lambda x, dtype=None: None if x is None else (x if dtype is None else x.to(dtype)).contiguous())
    monkeypatch.setattr(t.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(t, "empty", t.zeros)
    monkeypatch.setattr(t.Tensor, "is_cuda", property(lambda self: True))
    monkeypatch.setattr(t.cuda, "Event", _Event)
    monkeypatch.setattr(t.cuda, "Stream", lambda device=None: _Stream())
    monkeypatch.setattr(t.cuda, "current_stream", lambda device=None: _Stream())
    monkeypatch.setattr(t.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(t.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(t.cuda, "synchronize", lambda device=None: None)
    monkeypatch.setattr(t.cuda, "CUDAGraph", _Graph)
    monkeypatch.setattr(t.cuda, "graph", lambda g, **kw: contextlib.nullcontext())
    S, A, B, N = 6, 4, 16, 200
    fused = any(f.startswith("fused_") for f in flags)
    Net = lambda: model_mod.FusedMLP([S, 16, 16, A])
    monkeypatch.setattr(model_mod.FusedMLP, "forward", lambda self, state, action=None: t.nn.functional.linear(
        t.relu(t.nn.functional.linear(t.relu(t.nn.functional.linear(state, self.layers[0].weight, self.layers[0].bias)),
                                      self.layers[1].weight, self.layers[1].bias)), self.layers[2].weight, self.layers[2].bias))
    buf = make_cpu_buffer(PrioritizedBuffer, N, epsilon=1e-2, alpha=0.6, beta=0.4, curr_beta=0.4, beta_increment_per_sampling=1e-3,
                          wt_tree=_Tree(N))
    buf.storage = _Ring(N)
    buf.buffer_device = t.device("cpu", 0)
    learner = dqn_mod.DQNPer(Net(), Net(), model_mod.FusedAdam, t.nn.MSELoss(reduction="none"), batch_size=B, learning_rate=0.01,
                             discount=0.9, update_rate=0.1, replay_buffer=buf)
    buf.store_episode([dict(state={"state": t.randn(1, S)}, action={"action": t.tensor([[i % A]])},
                            next_state={"state": t.randn(1, S)}, reward=float(i), terminal=False) for i in range(50)],
                      priorities=list(np.random.RandomState(0).exponential(size=50)))
    for f in flags:
        if f == "fused_rows":
            learner.use_fused_program, learner.fused_kind = True, "rows"
        elif f == "fused_program":
            learner.use_fused_program, learner.fused_kind = True, "program"
        else:
            assert hasattr(learner, f)
            setattr(learner, f, True)
    if graph:
        learner.enable_cuda_graph()
    for _ in range(3):
        loss = learner.update_device()
        assert loss.shape == ()
    if graph:
        before, n_calls = _Graph.released, len(rec.calls)
        learner.release_graphs()
        assert _Graph.released - before == 1 and learner._graph is None
        assert learner.update_device().shape == () and len(rec.calls) > n_calls       # eager launches again
    called = {name for name, _ in rec.calls}
    want = {"mb_per_sample", "mb_tree_update_priority"}
    want |= {"mb_mlp_program_run_part"} if fused else {"mb_td_dqn"}
    want |= {"mb_adam_soft_step_multi"} if "fuse_optimizer_and_target" in flags else {"mb_adam_step_multi", "mb_soft_update_multi"}
    assert want <= called, sorted(want - called)


def _dry_run_patches(monkeypatch):
    import machin_b200._lib as lib_mod
    import machin_b200.ops as ops
    rec = _Recorder()
    _patch_library(monkeypatch, rec)
    monkeypatch.setattr(lib_mod, "stream_ptr", lambda stream=None: C.c_void_p(0))
    monkeypatch.setattr(ops, "_c", lambda x, dtype=None: None if x is None else (x if dtype is None else x.to(dtype)).contiguous())
    monkeypatch.setattr(t.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(t, "empty", t.zeros)
    monkeypatch.setattr(t.Tensor, "is_cuda", property(lambda self: True))
    monkeypatch.setattr(t.cuda, "Event", _Event)
    monkeypatch.setattr(t.cuda, "Stream", lambda device=None: _Stream())
    monkeypatch.setattr(t.cuda, "current_stream", lambda device=None: _Stream())
    monkeypatch.setattr(t.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(t.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(t.cuda, "synchronize", lambda device=None: None)
    monkeypatch.setattr(t.cuda, "CUDAGraph", _Graph)
    monkeypatch.setattr(t.cuda, "graph", lambda g, **kw: contextlib.nullcontext())
    return rec


def _per_double(N):
    from machin_b200.frame.buffers import PrioritizedBuffer
    buf = make_cpu_buffer(PrioritizedBuffer, N, epsilon=1e-2, alpha=0.6, beta=0.4, curr_beta=0.4, beta_increment_per_sampling=1e-3,
                          wt_tree=_Tree(N))
    buf.storage = _Ring(N)
    buf.buffer_device = t.device("cpu", 0)
    return buf


@pytest.mark.parametrize("mode", ["eager", "graph_sgd", "graph_fused_adam"])
def test_ddpg_per_update_dry_run(mode, monkeypatch):
    """DDPGPer.update (ddpg_per.py:92-185 mirror): actor / critic / two targets over the prioritized buffer double,
    eager and with the whole update captured (then released again)."""
    import machin_b200.frame.algorithms.ddpg_per as ddpg_mod
    import machin_b200.model as model_mod
    rec = _dry_run_patches(monkeypatch)
    S, A, B, N = 5, 2, 8, 100

    class Actor(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = t.nn.Linear(S, A)

        def forward(self, state):
            return t.tanh(self.fc(state))

    class Critic(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = t.nn.Linear(S + A, 1)

        def forward(self, state, action):
            return self.fc(t.cat([state, action], dim=1))

    optimizer = model_mod.FusedAdam if mode == "graph_fused_adam" else t.optim.SGD
    learner = ddpg_mod.DDPGPer(Actor(), Actor(), Critic(), Critic(), optimizer, t.nn.MSELoss(reduction="none"), batch_size=B,
                               replay_buffer=_per_double(N))
    learner.replay_buffer.store_episode(
        [dict(state={"state": t.randn(1, S)}, action={"action": t.randn(1, A)}, next_state={"state": t.randn(1, S)},
              reward=float(i), terminal=(i == 29)) for i in range(30)], priorities=list(np.ones(30)))
    if mode != "eager":
        before = [p.detach().clone() for p in learner.actor.parameters()]
        learner.enable_cuda_graph()
        assert all(t.equal(a, b) for a, b in zip(before, learner.actor.parameters()))      # warm-up undone in place
        assert learner.replay_buffer.curr_beta == 0.4
    for _ in range(2):
        out = learner.update()
        assert len(out) == 2 and all(np.isfinite(x) for x in out)
    if mode != "eager":
        assert learner.replay_buffer.curr_beta == pytest.approx(0.402)
        learner.release_graphs()
        assert learner._graph is None and len(learner.update()) == 2
        with pytest.raises(RuntimeError, match="soft target updates"):
            hard = ddpg_mod.DDPGPer(Actor(), Actor(), Critic(), Critic(), t.optim.SGD, t.nn.MSELoss(reduction="none"),
                                    batch_size=B, update_rate=None, update_steps=5, replay_buffer=_per_double(N))
            hard.enable_cuda_graph()
    called = {name for name, _ in rec.calls}
    assert {"mb_per_sample", "mb_td_target", "mb_tree_update_priority", "mb_soft_update_multi"} <= called, sorted(called)


def test_ddpg_apex_update_dry_run(monkeypatch):
    """DDPGApex.update over the sharded-buffer double on a group of one (the path tools/ddpg_apex_bench.py times)."""
    import machin_b200.frame.algorithms.apex as apex_mod
    import machin_b200.model as model_mod
    from machin_b200.frame.buffers import DistributedPrioritizedBuffer
    from machin_b200.parallel import LocalGroup
    rec = _dry_run_patches(monkeypatch)
    dev = t.device("cpu", 0)

    def buffer_double(buffer_name, group, buffer_size, requesters=None, **kw):
        buf = make_cpu_buffer(DistributedPrioritizedBuffer, buffer_size, epsilon=1e-2, alpha=0.6, beta=0.4, curr_beta=0.4,
                              beta_increment_per_sampling=1e-3, wt_tree=_Tree(buffer_size), buffer_name=buffer_name, group=group,
                              wr_lock=RLock(), _requesters=requesters, _sampler=None,
                              buffer_version_table=t.zeros(buffer_size, dtype=t.int64))
        buf.storage = _Ring(buffer_size)
        buf.buffer_device = dev
        buf.weight_sum_tensor = lambda: t.tensor([max(buf.wt_tree.get_weight_sum(), 1.0)], dtype=t.float64)
        return buf

    monkeypatch.setattr(apex_mod, "DistributedPrioritizedBuffer", buffer_double)
    S, A, B, N = 5, 2, 8, 100

    class Actor(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = t.nn.Linear(S, A)

        def forward(self, state):
            return t.tanh(self.fc(state))

    class Critic(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = t.nn.Linear(S + A, 1)

        def forward(self, state, action):
            return self.fc(t.cat([state, action], dim=1))

    group = LocalGroup(device="cpu")
    group.device = dev
    learner = apex_mod.DDPGApex(Actor(), Actor(), Critic(), Critic(), model_mod.FusedAdam, t.nn.MSELoss(reduction="none"), group, (),
                                batch_size=B, replay_size=N)
    learner.replay_buffer.store_episode(
        [dict(state={"state": t.randn(1, S)}, action={"action": t.randn(1, A)}, next_state={"state": t.randn(1, S)},
              reward=float(i), terminal=(i == 29)) for i in range(30)], priorities=list(np.ones(30)))
    for _ in range(2):
        out = learner.update()
        assert len(out) == 2
    called = {name for name, _ in rec.calls}
    assert {"mb_per_sample", "mb_rows_pack_meta", "mb_rows_unpack_meta", "mb_td_target", "mb_tree_update_priority",
            "mb_adam_step_multi", "mb_soft_update_multi"} <= called, sorted(called)
    # the same learner with the sharded update as one graph over the fixed-shape exchange (BASELINE C4's path)
    monkeypatch.setattr(t.cuda, "CUDAGraph", _Graph)
    monkeypatch.setattr(t.cuda, "graph", lambda g, **kw: contextlib.nullcontext())
    for overlap in (False, True):
        n0 = len(rec.calls)
        learner.overlap_priority_update = overlap
        learner.batched_draw = overlap
        learner.enable_cuda_graph(device_apportion=overlap, peer_exchange=True)     # (no peer mapping here: NCCL path, P = 1)
        for _ in range(2):
            out = learner.update()
            assert len(out) == 2
        new = {name for name, _ in rec.calls[n0:]}
        assert {"mb_wire_compact", "mb_wire_scatter_update", "mb_wire_update_mask", "mb_td_target", "mb_tree_update_priority",
                "mb_adam_step_multi"} <= new, sorted(new)
        learner.release_graphs()


@pytest.mark.parametrize("graph", [False, True])
def test_rainbow_update_dry_run(graph, monkeypatch):
    """RAINBOW (rainbow.py mirror): n-step store, the captured / eager update with the categorical value block, acting."""
    import machin_b200.frame.algorithms.rainbow as rb_mod
    import machin_b200.model as model_mod
    rec = _dry_run_patches(monkeypatch)
    S, A, ATOMS, B, N = 5, 3, 11, 8, 100

    class DistQ(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = t.nn.Linear(S, A * ATOMS)

        def forward(self, state):
            return t.softmax(self.fc(state).view(-1, A, ATOMS), dim=-1)

    learner = rb_mod.RAINBOW(DistQ(), DistQ(), model_mod.FusedAdam, -2.0, 3.0, batch_size=B, reward_future_steps=3,
                             replay_buffer=_per_double(N))
    for n in (1, 2, 7, 30):
        learner.store_episode([dict(state={"state": t.randn(1, S)}, action={"action": t.tensor([[i % A]])},
                                    next_state={"state": t.randn(1, S)}, reward=float(i % 3), terminal=(i == n - 1))
                               for i in range(n)])
    assert "value" in learner.replay_buffer.storage.fields
    if graph:
        learner.enable_cuda_graph()
    for _ in range(3):
        assert learner.update_device().shape == ()
    assert learner.act_discrete({"state": t.randn(4, S)}).shape == (4, 1)
    called = {name for name, _ in rec.calls}
    assert {"mb_per_sample", "mb_td_rainbow", "mb_tree_update_priority", "mb_adam_step_multi", "mb_soft_update_multi"} <= called
    assert "mb_td_dqn" not in called


def test_real_device_ring_dry_run(monkeypatch):
    """The REAL DeviceRing (not the host-memory double) under the recorder: schema layout, field descriptor arrays and
    every gather / scatter / unpack entry point are driven through Buffer's store and sample paths.  Nothing moves (the
    recorder launches nothing), but descriptor construction and argument types are the product's."""
    from machin_b200.frame.buffers import Buffer
    from machin_b200.frame.buffers.device_ring import DeviceRing
    from machin_b200._lib import MbField
    rec = _dry_run_patches(monkeypatch)
    dev = t.device("cpu", 0)
    buf = make_cpu_buffer(Buffer, 64)
    buf.storage, buf.buffer_device = DeviceRing(64, dev), dev
    ep = [dict(state={"state": t.randn(1, 5), "aux": t.randn(1, 2, 3)}, action={"action": t.tensor([[i % 3]], dtype=t.int)},
               next_state={"state": t.randn(1, 5), "aux": t.randn(1, 2, 3)}, reward=float(i), terminal=i == 6, note="text",
               score=i) for i in range(7)]
    buf.store_episode(ep)
    ring = buf.storage
    assert ring.pitch % 16 == 0 and ring.data.shape == (64, ring.pitch) and len(ring) == 7
    offsets = sorted((f.offset, f.nbytes) for f in ring.fields.values())
    assert all(a + n <= b for (a, n), (b, _) in zip(offsets, offsets[1:])) and offsets[-1][0] + offsets[-1][1] <= ring.pitch
    assert ring.fields["reward"].scalar and ring.fields["action/action"].dtype == t.int32 and "note" not in ring.fields
    n, (state, action, reward, others) = buf.sample_batch(4, sample_attrs=["state", "action", "reward", "*"],
                                                          additional_concat_custom_attrs=["score"])
    assert n == 4 and state["aux"].shape == (4, 2, 3) and action["action"].dtype == t.int32 and reward.shape == (4, 1)
    assert others["score"].shape == (4, 1) and len(others["note"]) == 4
    n, (state, reward) = buf.sample_batch(3, concatenate=False, sample_attrs=["state", "reward"])
    assert n == 3 and len(state["state"]) == 3 and isinstance(reward[0], float)
    buf.store_episode_tensors = Buffer.store_episode_tensors.__get__(buf)
    idx = t.tensor([0, 3, 5])
    packed = ring.gather_packed(idx)
    assert packed.shape == (3, ring.pitch)
    out = {k: t.zeros((3,) + f.shape, dtype=f.dtype) for k, f in ring.fields.items()}
    assert ring.unpack(packed, out=out) is out and set(ring.unpack(packed, names=["reward", "nope"])) == {"reward"}
    wire = t.zeros((3, ring.pitch + 24), dtype=t.uint8)
    ring.gather_into_wire(idx, wire)
    with pytest.raises(RuntimeError, match="schema already fixed"):
        ring.set_schema([])
    calls = [name for name, _ in rec.calls]
    assert {"mb_rows_scatter", "mb_rows_gather", "mb_rows_gather_packed"} <= set(calls)
    gathers = [args for name, args in rec.calls if name == "mb_rows_gather"]
    assert all(isinstance(a[5], C.Array) and a[5]._type_ is MbField for a in gathers)


def test_buffer_constructors_dry_run(monkeypatch):
    """The real constructors (they need a device: tests/fake_ring.py bypasses them everywhere else): argument routing
    through the RNN / prioritized class hierarchy, device normalisation, the loud refusals."""
    import machin_b200._lib as lib_mod
    import machin_b200.frame.buffers.prioritized_buffer as per_mod
    from machin_b200.frame.buffers import (Buffer, DistributedBuffer, DistributedPrioritizedBuffer, PrioritizedBuffer,
                                           RNNBuffer, RNNPrioritizedBuffer)
    from machin_b200.parallel import LocalGroup
    _dry_run_patches(monkeypatch)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Buffer(10, "cuda")                                       # the recorder reports no device: refused loudly
    monkeypatch.setattr(lib_mod, "require_device", lambda index=0: None)
    monkeypatch.setattr(per_mod, "WeightTree", lambda size, device: _Tree(size))
    with pytest.raises(ValueError, match="must be a CUDA device"):
        Buffer(10, "cpu")
    with pytest.raises(ValueError, match="custom storage is not supported"):
        Buffer(10, "cuda", storage=object())
    b = Buffer(10)
    assert b.buffer_device == t.device("cuda", 0) and b.storage.max_size == 10 and b.size() == 0
    p = PrioritizedBuffer(16, "cuda:0", epsilon=0.1, alpha=0.5, beta=0.3, beta_increment_per_sampling=0.01)
    assert (p.epsilon, p.alpha, p.beta, p.curr_beta, p.beta_increment_per_sampling) == (0.1, 0.5, 0.3, 0.3, 0.01)
    assert p.wt_tree.size == 16
    r = RNNBuffer(sample_length=4, buffer_size=20)
    assert (r.sample_length, r.sample_dimension, r.storage.max_size) == (4, 1, 20)
    with pytest.raises(ValueError, match="custom storage is not supported"):
        RNNBuffer(sample_length=4, storage=object())
    rp = RNNPrioritizedBuffer(sample_length=3, sample_dimension=0, buffer_size=32, alpha=0.7)      # the reference cannot
    assert (rp.sample_length, rp.sample_dimension, rp.alpha, rp.wt_tree.size) == (3, 0, 0.7, 32)   # construct this one
    group = LocalGroup("g", device="cpu")
    group.device = t.device("cuda", 0)
    d = DistributedBuffer("buffer", group, buffer_size=12)
    assert d.buffer_name == "buffer" and d.requesters == [0] and d.buffer_device == t.device("cuda", 0)
    real_zeros = t.zeros                                # (the version table is allocated on the device right away)
    monkeypatch.setattr(t, "zeros", lambda *a, device=None, **k: real_zeros(*a, **k))
    dp = DistributedPrioritizedBuffer("buffer", group, buffer_size=12, beta=0.5)
    assert dp.curr_beta == 0.5 and dp.buffer_version_table.shape == (12,) and dp.buffer_version_table.dtype == t.int64


@pytest.mark.parametrize("variant", ["default", "device_apportion+batched_draw"])
def test_fixed_shape_update_dry_run_as_member_of_two(variant, monkeypatch):
    """The P > 1 branches of the fixed-shape update (captured all-to-alls, gradient all-reduce, per-requester draws into
    [P, cap] blocks, one gather and one metadata launch over P * cap rows) executed in ONE process: a group double of
    size 2 whose collectives copy this member's own contribution into every slot."""
    import torch.distributed as dist
    import machin_b200.frame.algorithms.apex as apex_mod
    from machin_b200.frame.buffers import DistributedPrioritizedBuffer
    from machin_b200.parallel import LocalGroup
    rec = _dry_run_patches(monkeypatch)
    dev = t.device("cpu", 0)

    class GroupOfTwo(LocalGroup):
        index, pg = 0, None

        def size(self):
            return 2

        def all_gather_f64(self, value):
            return t.stack([value, value])

    monkeypatch.setattr(dist, "all_to_all_single", lambda out, inp, group=None: out.copy_(inp))
    monkeypatch.setattr(dist, "all_reduce", lambda x, group=None, op=None: None)
    monkeypatch.setattr(dist, "all_gather_into_tensor", lambda out, inp, group=None: out.copy_(inp.repeat(out.numel() // inp.numel())))

    def buffer_double(buffer_name, group, buffer_size, requesters=None, **kw):
        buf = make_cpu_buffer(DistributedPrioritizedBuffer, buffer_size, epsilon=1e-2, alpha=0.6, beta=0.4, curr_beta=0.4,
                              beta_increment_per_sampling=1e-3, wt_tree=_Tree(buffer_size), buffer_name=buffer_name, group=group,
                              wr_lock=RLock(), _requesters=requesters, _sampler=None,
                              buffer_version_table=t.zeros(buffer_size, dtype=t.int64))
        buf.storage, buf.buffer_device = _Ring(buffer_size), dev
        buf.weight_sum_tensor = lambda: t.tensor([buf.wt_tree.get_weight_sum()], dtype=t.float64)
        buf._agree_on_schema = lambda: None                     # (a collective over Python objects; both members stored)
        return buf

    monkeypatch.setattr(apex_mod, "DistributedPrioritizedBuffer", buffer_double)
    S, A, B, N = 6, 4, 16, 200
    net = lambda: t.nn.Sequential(t.nn.Linear(S, 16), t.nn.ReLU(), t.nn.Linear(16, A))

    class Q(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.net = net()

        def forward(self, state):
            return self.net(state)

    group = GroupOfTwo(device="cpu")
    group.device = dev
    learner = apex_mod.DQNApex(Q(), Q(), t.optim.SGD, t.nn.MSELoss(reduction="none"), group, (), batch_size=B, learning_rate=0.01,
                               discount=0.9, update_rate=0.1, replay_size=N, learner_group=group)
    learner.replay_buffer.store_episode(
        [dict(state={"state": t.randn(1, S)}, action={"action": t.tensor([[i % A]])}, next_state={"state": t.randn(1, S)},
              reward=float(i), terminal=False) for i in range(50)], priorities=list(np.random.RandomState(0).exponential(size=50)))
    learner.batched_draw = "batched_draw" in variant
    learner.enable_cuda_graph(static_exchange=True, device_apportion=variant.startswith("device_apportion"))
    st = learner._static
    assert st["P"] == 2 and st["d_idx"].shape == (2, B) and st["send"].shape == (2 * B, st["W"]) and st["capB"] == B + 2
    for _ in range(3):
        assert learner.update_device().shape == ()
    if not variant.startswith("device"):
        assert int(st["counts_pin"].sum()) in (B, B + 1)        # ceil apportioning over two equal shards
    names = [name for name, _ in rec.calls]
    gathers = [args for name, args in rec.calls if name == "mb_rows_pack_meta"]
    assert gathers and all(a[3] == 2 * B for a in gathers)      # ONE metadata launch over both requester blocks
    assert ("mb_per_sample_multi" in names) == ("batched_draw" in variant)
    assert {"mb_wire_compact", "mb_wire_scatter_update", "mb_wire_update_mask", "mb_tree_update_priority"} <= set(names)
    learner.push_weights()                                      # every member learns: nothing to broadcast


def test_impala_fixed_shape_graph_dry_run(monkeypatch):
    """IMPALA.enable_cuda_graph for trajectories of one length (BASELINE C3): staging of the host-picked episode slots,
    replay, the eager fall-back for a batch of another shape (Python's random consumed once either way), release."""
    import random
    from threading import RLock
    import machin_b200.frame.algorithms.impala as impala_mod
    from machin_b200.frame.buffers import IMPALABuffer
    from machin_b200.parallel import LocalGroup
    rec = _dry_run_patches(monkeypatch)
    dev = t.device("cpu", 0)
    S, A, T, B = 5, 3, 4, 6

    def buffer_double(buffer_name, group, buffer_size, requesters=None):
        buf = make_cpu_buffer(IMPALABuffer, buffer_size, buffer_name=buffer_name, group=group, wr_lock=RLock(),
                              requesters=requesters or [0], _schema_agreed=False)
        buf.buffer_device = dev
        return buf

    monkeypatch.setattr(impala_mod, "IMPALABuffer", buffer_double)

    class Policy(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = t.nn.Linear(S, A)

        def forward(self, state, action=None):
            dist = t.distributions.Categorical(logits=self.fc(state))
            act = action if action is not None else dist.sample().view(-1, 1)
            return act, dist.log_prob(act.flatten()).view(-1, 1), dist.entropy().view(-1, 1)

    class Value(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = t.nn.Linear(S, 1)

        def forward(self, state):
            return self.fc(state)

    group = LocalGroup("impala", device="cpu")
    group.device = dev
    imp = impala_mod.IMPALA(Policy(), Value(), t.optim.SGD, t.nn.MSELoss(reduction="sum"), group, (), batch_size=B, replay_size=400)

    def episode(n):
        return [dict(state={"state": t.randn(1, S)}, action={"action": t.tensor([[i % A]])}, next_state={"state": t.randn(1, S)},
                     reward=float(i), action_log_prob=-0.5, terminal=i == n - 1) for i in range(n)]

    for _ in range(9):
        imp.store_episode(episode(T))
    before = [p.detach().clone() for p in imp.actor.parameters()]
    random.seed(5)
    state = random.getstate()
    imp.enable_cuda_graph(episode_length=T)
    assert random.getstate() == state                            # the warm-up's episode picks are put back
    assert all(t.equal(a, b) for a, b in zip(before, imp.actor.parameters()))
    st = imp._graph_state
    assert st["idx"].shape == (B * T,) and st["offsets"].tolist() == list(range(0, B * T + 1, T))
    n_calls = len(rec.calls)
    pol, val = imp.update_device()
    assert pol.shape == () and len(rec.calls) == n_calls         # a replay launches nothing through the entry points
    assert random.getstate() != state and sorted(set(st["idx_pin"].tolist())) == sorted(set(st["idx"].tolist()))
    assert len(imp.update()) == 2
    imp.store_episode(episode(T + 3))                            # a longer episode may now be drawn: eager fall-back
    random.seed(0)
    for _ in range(6):
        out = imp.update_device()
        assert out is not None and out[0].shape == ()
    assert len(rec.calls) > n_calls                              # at least one of those batches ran eagerly
    with pytest.raises(RuntimeError, match="sampled episodes of exactly"):
        other = impala_mod.IMPALA(Policy(), Value(), t.optim.SGD, t.nn.MSELoss(reduction="sum"), group, (), batch_size=B,
                                  replay_size=400)
        other.store_episode(episode(T + 1))
        other.enable_cuda_graph(episode_length=T)
    imp.release_graphs()
    assert imp._graph is None and imp.update_device()[0].shape == ()
