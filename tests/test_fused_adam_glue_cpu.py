"""CPU: the Python glue of FusedAdam (machin_b200/model.py) with a recording stand-in for the C library: which entry
point is called, with how many tensors, which pointers and hyper-parameters -- the kernels themselves are checked
under emulation (tests/test_optim_kernel_emulated_cpu.py) and on the device."""
import ctypes as C

import torch as t

import machin_b200.model as model_mod


class _Recorder:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def call(*args):
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
