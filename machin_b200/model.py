"""Fused value networks and optimizer for the learner step (optional fast path).

``FusedMLP`` is an ``nn.Module`` for plain ``Linear -> ReLU -> ... -> Linear`` stacks whose forward and
backward run one libmachin_b200 kernel per layer per direction (csrc/mlp.cu) instead of ~8 library
launches per layer; parameters are ordinary ``nn.Linear`` weights/biases, so ``state_dict``, soft/hard
target updates, model servers and checkpoints see a normal module.  ``FusedAdam`` is
``torch.optim.Adam`` (amsgrad off) as one multi-tensor kernel, CUDA-graph capturable.
Any other ``nn.Module`` / optimizer keeps working through the generic autograd path.
"""
import ctypes as C
import os
from typing import Sequence

import torch as t
import torch.nn as nn

from . import _lib
from ._lib import lib


class _Workspace:
    """Split-K scratch shared by the layers of one network (launches are stream-ordered)."""

    def __init__(self, device, floats=1 << 21, counters=256):
        self.buf = t.empty(floats, dtype=t.float32, device=device)
        self.cnt = t.zeros(counters, dtype=t.int32, device=device)

    def args(self):
        return _lib.ptr(self.buf), self.buf.numel(), _lib.ptr(self.cnt), self.cnt.numel()


def _linear_fwd(x, w, b, relu, ws):
    M, K = x.shape
    N = w.shape[0]
    y = t.empty((M, N), dtype=t.float32, device=x.device)
    _lib.check(lib.mb_mlp_linear_forward(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), M, N, K, int(relu),
                                         *ws.args(), _lib.stream_ptr()))
    return y


class _MLPFunction(t.autograd.Function):
    @staticmethod
    def forward(ctx, x, ws, side_ws, sinks, *params):
        x = x.contiguous().float()
        n_layers = len(params) // 2
        acts = [x]
        for l in range(n_layers):
            acts.append(_linear_fwd(acts[-1], params[2 * l], params[2 * l + 1], l < n_layers - 1, ws))
        ctx.ws, ctx.side_ws, ctx.sinks = ws, side_ws, sinks
        ctx.n_layers = n_layers
        ctx.save_for_backward(*acts[:-1], *params[0::2])
        return acts[-1]

    @staticmethod
    def backward(ctx, dy):
        """Input gradients form a dependent chain (layer l needs layer l+1's); weight / bias gradients only feed the
        optimizer and are independent of each other.  So the weight-gradient GEMM of every layer runs on its own side
        stream -- parallel branches of a captured update -- beside the chain, and the main stream joins them at the end
        (`PARALLEL_WEIGHT_GRADS`)."""
        n, ws = ctx.n_layers, ctx.ws
        saved = ctx.saved_tensors
        inputs, weights = saved[:n], saved[n:]
        dy = dy.contiguous().float()
        dev = dy.device
        grads = [None] * (2 * n)
        dx = None
        sides = [_side_stream(dev, l) for l in range(n)] if (PARALLEL_WEIGHT_GRADS and n > 1) else None
        cur = t.cuda.current_stream(dev) if sides is not None else None
        side_ws = ctx.side_ws if sides is not None else None
        keep, used = [], []
        for l in range(n - 1, -1, -1):
            x, w = inputs[l], weights[l]
            M, K = x.shape
            N = w.shape[0]
            # gradient sinks (FusedMLP.set_grad_sinks): the kernel writes dW / db straight into the caller's buffers (e.g.
            # views of one flat all-reduce buffer) and autograd gets None -- no zero-fill, no per-parameter accumulate launches
            sink = ctx.sinks is not None
            dw = ctx.sinks[2 * l] if sink else t.empty_like(w)
            db = ctx.sinks[2 * l + 1] if sink else t.empty(N, dtype=t.float32, device=w.device)
            # the first layer's weight gradient is the LAST kernel of the backward (nothing is left on the main stream to run
            # beside it): on a branch it would only add a fork and a join to the critical path
            on_branch = sides is not None and not (LAST_WEIGHT_GRAD_ON_MAIN and l == 0 and not ctx.needs_input_grad[0])
            if on_branch:
                side = sides[l]                            # one branch per layer: the weight gradients are independent of each other
                used.append(side)
                side.wait_stream(cur)                      # dy of this layer is ready
                with t.cuda.stream(side):
                    _lib.check(lib.mb_mlp_linear_backward_weight(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(dw), _lib.ptr(db), M, N, K,
                                                                 *side_ws[l].args(), _lib.stream_ptr()))
            else:
                _lib.check(lib.mb_mlp_linear_backward_weight(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(dw), _lib.ptr(db), M, N, K,
                                                             *ws.args(), _lib.stream_ptr()))
            if not sink:
                grads[2 * l], grads[2 * l + 1] = dw, db
            if l > 0 or ctx.needs_input_grad[0]:
                dxl = t.empty((M, K), dtype=t.float32, device=w.device)
                _lib.check(lib.mb_mlp_linear_backward_input(_lib.ptr(dy), _lib.ptr(w), _lib.ptr(x) if l > 0 else None,
                                                            _lib.ptr(dxl), M, N, K, *ws.args(), _lib.stream_ptr()))
                keep.append(dy)                            # the side stream may still be reading it: alive until the join
                dy = dxl
                if l == 0:
                    dx = dxl
        for side in used:                                  # (only branches that were forked: a stream outside a capture
            cur.wait_stream(side)                          #  must not be joined into it) every dw / db is complete before
                                                           #  autograd hands them on
        del keep
        return (dx, None, None, None, *grads)


PARALLEL_WEIGHT_GRADS = True       # FusedMLP.backward: weight-gradient GEMMs on a side stream (see _MLPFunction.backward)
LAST_WEIGHT_GRAD_ON_MAIN = "last_grad_main" not in os.environ.get("MB_STEP_OFF", "").split(",")
_SIDE_STREAMS = {}


def _side_stream(device, slot=0):
    key = (device.type, device.index, slot)
    s = _SIDE_STREAMS.get(key)
    if s is None:
        s = _SIDE_STREAMS[key] = t.cuda.Stream(device=device)
    return s


class FusedMLP(nn.Module):
    def __init__(self, sizes: Sequence[int]):
        """``sizes`` = [in, hidden..., out]; ReLU between layers, none after the last."""
        super().__init__()
        self.sizes = list(sizes)
        self.layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(sizes[:-1], sizes[1:])])
        self._ws = None
        self._grad_sinks = None

    def set_grad_sinks(self, sinks):
        """``sinks``: one contiguous float32 CUDA tensor per parameter (weight, bias per layer, parameter order) that the
        backward kernels write the gradients INTO -- typically views of one flat buffer that is all-reduced as a whole;
        autograd then receives no gradient for the parameters (``p.grad`` should be set to the same views by the caller).
        ``None`` restores ordinary autograd accumulation."""
        if sinks is not None:
            sinks = tuple(sinks)
            params = [p for layer in self.layers for p in (layer.weight, layer.bias)]
            if len(sinks) != len(params) or any(s.shape != p.shape or s.dtype != t.float32 or not s.is_contiguous()
                                                for s, p in zip(sinks, params)):
                raise ValueError("grad sinks must match the parameters: contiguous float32 tensors of the same shapes")
        self._grad_sinks = sinks

    def _workspace(self, device, side=0):
        """One split-K scratch per (device, stream): two forwards of the same network may run on different streams at
        the same time (DQNPer.parallel_forwards), and launches are only ordered within a stream.  ``side``: the scratch
        of the weight-gradient branch of this stream's backward."""
        key = (device, getattr(t.cuda.current_stream(device), "cuda_stream", 0), side)
        if self._ws is None:
            self._ws = {}
        ws = self._ws.get(key)
        if ws is None:
            ws = self._ws[key] = _Workspace(device) if not side else _Workspace(device, floats=1 << 19)
        return ws

    def forward(self, state, action=None):
        x = state if action is None else t.cat([state, action.to(state.dtype)], dim=1)
        if not x.is_cuda:
            raise RuntimeError("FusedMLP runs on CUDA tensors only (no CPU fallback)")
        params = [p for layer in self.layers for p in (layer.weight, layer.bias)]
        with t.cuda.device(x.device):
            sides = tuple(self._workspace(x.device, side=l + 1) for l in range(len(self.layers)))   # one per weight-gradient branch
            return _MLPFunction.apply(x, self._workspace(x.device), sides, self._grad_sinks, *params)


class FusedAdam(t.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @t.no_grad()
    def step_with_soft_update(self, target_of, update_rate: float):
        """One Adam step and, in the same launch, ``target = target * (1 - rate) + param * rate`` on the matching
        target-network parameters (``target_of``: dict id(param) -> target parameter).  Every parameter with a
        gradient must have a target; parameters are float32 CUDA tensors."""
        return self.step(_target_of=target_of, _update_rate=update_rate)

    @t.no_grad()
    def step(self, closure=None, _target_of=None, _update_rate=0.0):
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                st = self.state[p]
                if not st:
                    st["exp_avg"] = t.zeros_like(p)
                    st["exp_avg_sq"] = t.zeros_like(p)
            first = self.state[group["params"][0]]
            if "step" not in first:
                first["step"] = t.zeros((), dtype=t.float32, device=ps[0].device)
            n = len(ps)
            arr = lambda xs: (C.c_void_p * n)(*[x.data_ptr() for x in xs])
            sizes = (C.c_int64 * n)(*[p.numel() for p in ps])
            b1, b2 = group["betas"]
            if _target_of is not None and n <= _lib.MB_MAX_SOFT_TENSORS:
                if "ticket" not in first:
                    first["ticket"] = t.zeros(1, dtype=t.int32, device=ps[0].device)
                with t.cuda.device(ps[0].device):
                    _lib.check(lib.mb_adam_soft_step_multi(
                        arr(ps), arr([p.grad for p in ps]), arr([self.state[p]["exp_avg"] for p in ps]),
                        arr([self.state[p]["exp_avg_sq"] for p in ps]), arr([_target_of[id(p)] for p in ps]), sizes, n,
                        _lib.ptr(first["step"]), _lib.ptr(first["ticket"]), float(group["lr"]), float(b1), float(b2),
                        float(group["eps"]), float(group["weight_decay"]), float(_update_rate), _lib.stream_ptr()))
                continue
            with t.cuda.device(ps[0].device):
                _lib.check(lib.mb_adam_step_multi(arr(ps), arr([p.grad for p in ps]),
                                                  arr([self.state[p]["exp_avg"] for p in ps]),
                                                  arr([self.state[p]["exp_avg_sq"] for p in ps]), sizes, n,
                                                  _lib.ptr(first["step"]), float(group["lr"]), float(b1), float(b2),
                                                  float(group["eps"]), float(group["weight_decay"]), _lib.stream_ptr()))
