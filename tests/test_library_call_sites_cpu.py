"""CPU: the Python glue of FusedAdam (machin_b200/model.py) with a recording stand-in for the C library: which entry
point is called, with how many tensors, which pointers and hyper-parameters -- the kernels themselves are checked
under emulation (tests/test_optim_kernel_emulated_cpu.py) and on the device."""
import ctypes as C

import torch as t

import machin_b200.model as model_mod


class _Recorder:
    """Stands in for the ctypes library object: records every call and converts the arguments the way ctypes would
    (``argtype.from_param``), so a None where an int is declared, a float for an int64, a wrong argument count ...
    raise here exactly like they would in front of the real library."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        from machin_b200._lib import SIGNATURES
        _res, argtypes = SIGNATURES[name]

        def call(*args):
            assert len(args) == len(argtypes), (name, len(args), len(argtypes))
            for a, tp in zip(args, argtypes):
                tp.from_param(a)
            self.calls.append((name, args))
            return 0
        return call


def _vals(arr, n):
    return [arr[i] for i in range(n)]


def test_fused_adam_calls(monkeypatch):
    import contextlib
    rec = _Recorder()
    monkeypatch.setattr(model_mod, "lib", rec)
    monkeypatch.setattr(t.cuda, "device", lambda d: contextlib.nullcontext())          # no CUDA device in this suite
    monkeypatch.setattr(model_mod._lib, "stream_ptr", lambda stream=None: C.c_void_p(0))
    net, tgt = t.nn.Linear(3, 2), t.nn.Linear(3, 2)
    opt = model_mod.FusedAdam(net.parameters(), lr=2e-3, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.01)
    for p in net.parameters():
        p.grad = t.ones_like(p)
    opt.step()
    name, a = rec.calls[-1]
    assert name == "mb_adam_step_multi" and a[5] == 2
    ps = list(net.parameters())
    assert _vals(a[0], 2) == [p.data_ptr() for p in ps] and _vals(a[1], 2) == [p.grad.data_ptr() for p in ps]
    assert _vals(a[4], 2) == [p.numel() for p in ps]
    assert a[7:12] == (2e-3, 0.8, 0.95, 1e-6, 0.01)
    state = opt.state[ps[0]]
    assert state["step"].dtype == t.float32 and a[6].value == state["step"].data_ptr()
    # the fused step + soft target update
    pairs = {id(p): tp for p, tp in zip(net.parameters(), tgt.parameters())}
    opt.step_with_soft_update(pairs, 0.005)
    name, a = rec.calls[-1]
    assert name == "mb_adam_soft_step_multi" and a[6] == 2
    assert _vals(a[4], 2) == [tp.data_ptr() for tp in tgt.parameters()]
    assert a[7].value == state["step"].data_ptr() and a[8].value == state["ticket"].data_ptr()
    assert state["ticket"].dtype == t.int32 and int(state["ticket"]) == 0
    assert a[9:15] == (2e-3, 0.8, 0.95, 1e-6, 0.01, 0.005)
    # parameters without a gradient are left out of the launch
    ps[1].grad = None
    opt.step()
    assert rec.calls[-1][1][5] == 1


def test_fused_program_creation_arguments(monkeypatch):
    """DQNPer._fused_program(): the argument order handed to mb_dqn_mlp_rows_create / mb_dqn_mlp_program_create must be
    the C prototype's (include/machin_b200.h): device, n_layers, sizes, batch, params, target params, grads, state,
    next_state, action, reward, terminal, is_weight, discount, mode, y, action_value, abs_error, loss_elem, out."""
    import contextlib
    import numpy as np
    import machin_b200._lib as lib_mod
    from fake_ring import FakeTree, make_cpu_buffer
    from machin_b200.frame.algorithms import DQNPer
    from machin_b200.frame.buffers import PrioritizedBuffer
    rec = _Recorder()
    monkeypatch.setattr(lib_mod, "lib", rec)
    monkeypatch.setattr(t.cuda, "device", lambda d: contextlib.nullcontext())
    sizes, B = [6, 16, 16, 4], 32
    q, qt = model_mod.FusedMLP(sizes), model_mod.FusedMLP(sizes)
    buf = make_cpu_buffer(PrioritizedBuffer, 200, epsilon=1e-2, alpha=0.6, beta=0.4, curr_beta=0.4,
                          beta_increment_per_sampling=1e-3, wt_tree=FakeTree(200))
    buf.buffer_device = t.device("cpu", 0)               # (a device WITH an index, like "cuda:0")
    dqn = DQNPer(q, qt, t.optim.SGD, t.nn.MSELoss(reduction="none"), batch_size=B, discount=0.9, update_rate=0.1,
                 replay_buffer=buf)
    buf.store_episode([dict(state={"state": t.zeros(1, 6)}, action={"action": t.tensor([[1]])},
                            next_state={"state": t.zeros(1, 6)}, reward=0.5, terminal=False) for _ in range(5)],
                      priorities=list(np.ones(5)))
    for kind, entry in (("rows", "mb_dqn_mlp_rows_create"), ("program", "mb_dqn_mlp_program_create")):
        dqn.use_fused_program, dqn.fused_kind, dqn._program = True, kind, None
        st = dqn._fused_program()
        assert st, "the program was not built"
        name, a = [c for c in rec.calls if c[0].startswith("mb_dqn_mlp")][-1]
        assert name == entry and len(a) == 20
        assert a[1] == 3 and list(a[2]) == sizes and a[3] == B
        plist = lambda net: [p for layer in net.layers for p in (layer.weight, layer.bias)]
        assert _vals(a[4], 6) == [p.data_ptr() for p in plist(dqn.qnet)]
        assert _vals(a[5], 6) == [p.data_ptr() for p in plist(dqn.qnet_target)]
        assert _vals(a[6], 6) == [p.grad.data_ptr() for p in plist(dqn.qnet)]
        f = st["fields"]
        want = [f["state/state"], f["next_state/state"], f["action/action"], f["reward"], f["terminal"], st["isw"]]
        assert [x.value for x in a[7:13]] == [w.data_ptr() for w in want]
        assert (f["state/state"].shape, f["action/action"].dtype, f["terminal"].dtype) == ((B, 6), t.int64, t.bool)
        assert abs(a[13] - 0.9) < 1e-12 and a[14] == 0                      # discount, mode 0 = double DQN
        assert [x.value for x in a[15:19]] == [st[k].data_ptr() for k in ("y", "av", "abs", "loss")]
        assert (st["y"].dtype, st["abs"].shape, st["loss"].dtype) == (t.float32, (B,), t.float64)


def test_ops_call_sites_convert_like_ctypes(monkeypatch):
    """machin_b200.ops wrappers: argument count / types of the calls into the library (typed recorder), incl. the
    entry points that take the longest episode length."""
    import contextlib
    import machin_b200.ops as ops
    rec = _Recorder()
    monkeypatch.setattr(ops, "lib", rec)
    monkeypatch.setattr(t.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(ops._lib, "stream_ptr", lambda stream=None: C.c_void_p(0))
    monkeypatch.setattr(ops, "_c", lambda x, dtype=None: None if x is None else (x if dtype is None else x.to(dtype)).contiguous())
    n = 12
    f = lambda: t.randn(n, 1)
    term = t.zeros(n, 1, dtype=t.bool)
    off = t.tensor([0, 5, 12])
    out = ops.vtrace(f(), f(), f(), term, -t.rand(n, 1), -t.rand(n, 1), 0.99, 1.0, 1.0, offsets=off, max_len=7)
    name, a = rec.calls[-1]
    assert name == "mb_vtrace_len" and a[1:4] == (2, n, 7) and set(out) == {"vs", "vss", "rho", "adv"}
    ops.vtrace(f(), f(), f(), term, f(), f(), 0.99, n_episodes=3)
    assert rec.calls[-1][1][0].value in (None, 0) and rec.calls[-1][1][1:4] == (3, n, 0)
    out = ops.gae(t.randn(n, dtype=t.float64), term.view(-1), t.randn(n), t.randn(n), 0.97, 0.95, n_episodes=1)
    name, a = rec.calls[-1]
    assert name == "mb_gae_len" and a[1:4] == (1, n, 0) and a[8:10] == (0.97, 0.95) and set(out) == {"returns", "gae"}
    td = ops.td_dqn(t.randn(n, 4), t.randn(n, 4), t.randn(n, 4), t.zeros(n, 1, dtype=t.int64), f(), term, 0.9,
                    is_weight=t.rand(n, dtype=t.float64), want_grad=True)
    name, a = rec.calls[-1]
    assert name == "mb_td_dqn" and a[3:5] == (n, 4) and td["grad_q"].shape == (n, 4)
    ops.td_target(f(), f(), term, 0.9, value=f())
    assert rec.calls[-1][0] == "mb_td_target"
