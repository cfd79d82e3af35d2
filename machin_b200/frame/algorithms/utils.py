"""Helpers shared by the algorithm mirrors (machin/frame/algorithms/utils.py:8-161)."""
import inspect
from typing import Dict

import torch as t
import torch.nn as nn

from ... import ops

_SPEC_CACHE: Dict[tuple, tuple] = {}


def _unwrap(model):
    if isinstance(model, (nn.parallel.DistributedDataParallel, nn.parallel.DataParallel)):
        return model.module
    return model


def model_device(model) -> t.device:
    model = _unwrap(model)
    if hasattr(model, "input_device"):
        return t.device(model.input_device)
    devices = {p.device for p in model.parameters()}
    if len(devices) != 1:
        raise RuntimeError(f"Failed to automatically determine i/o device of your model: {type(model)}; "
                           f"detected devices: {devices}")
    return next(iter(devices))


def safe_call(model, *named_args, method="__call__"):
    """Call ``model`` with the dict entries its forward() names, moved to the model's input device;
    always returns a tuple (utils.py:52-161)."""
    inner = _unwrap(model)
    fn = inner.forward if method == "__call__" else getattr(inner, method)
    key = (type(inner), method)          # the class itself, not its id: ids are reused after a class is collected
    spec = _SPEC_CACHE.get(key)
    if spec is None:
        s = inspect.getfullargspec(fn)
        args = s.args[1:] + s.kwonlyargs
        with_defaults = args[len(s.args[1:]) - len(s.defaults):len(s.args[1:])] if s.defaults else []
        required = set(args) - set(with_defaults) - set((s.kwonlydefaults or {}).keys())
        spec = (set(args), required, s.varargs is not None or s.varkw is not None)
        _SPEC_CACHE[key] = spec
    args, required, open_ended = spec
    dev = model_device(model)
    call_args = {}
    for na in named_args:
        for k, v in na.items():
            if k in args or open_ended:
                call_args[k] = v.to(dev) if t.is_tensor(v) else v
    missing = required - set(call_args)
    if missing:
        raise RuntimeError(f"Required arguments of the forward function of Model {type(inner)} is {required}, "
                           f"missing required arguments: {missing}\n\nCheck your storage functions.")
    result = (model if method == "__call__" else getattr(model, method))(**call_args)
    return result if isinstance(result, tuple) else (result,)


def hard_update(target_net: nn.Module, source_net: nn.Module) -> None:
    with t.no_grad():
        for tb, b in zip(target_net.buffers(), source_net.buffers()):
            tb.data.copy_(b.data)
        for tp, p in zip(target_net.parameters(), source_net.parameters()):
            tp.data.copy_(p.data)


def soft_update(target_net: nn.Module, source_net: nn.Module, update_rate) -> None:
    """target = target * (1 - rate) + source * rate, one kernel per parameter block (csrc/td.cu)."""
    with t.no_grad():
        ops.soft_update(target_net, source_net, update_rate)


def step_option(name: str) -> bool:
    """Stream-level options of the captured update.  All on by default; ``MB_STEP_OFF=capture_stream,last_grad_main,
    third_branch`` switches the named ones off (A/B timing: bench.py)."""
    import os
    return name not in os.environ.get("MB_STEP_OFF", "").split(",")


def capture_graph(graph, warmup_stream, **kwargs):
    """``torch.cuda.graph(graph, ...)`` capturing ON the stream the warm-up iterations ran on.  torch's default is an
    internal capture stream of its own -- a different handle, so everything keyed by the current stream (FusedMLP's
    workspaces: model.py) would be created again inside the capture, and the zero-fill of its counters would be replayed
    in front of the online forward by every update (4 launches of ~2 us on the critical path, profiles/
    r2_launches_default_step.csv IDs 8-11)."""
    if step_option("capture_stream"):
        kwargs["stream"] = warmup_stream
    return t.cuda.graph(graph, **kwargs)


class capture_safe:
    """Context for CUDA-graph capture of a user model: ``torch.distributions`` validates its arguments with a
    host-synchronising ``_is_all_true`` (illegal while a stream is capturing, and a failed capture leaves torch's
    CUDA generator unusable for the rest of the process), so validation is off inside and restored after.
    The warm-up iterations run outside, with validation on."""

    def __enter__(self):
        import torch.distributions as D
        self._was = D.Distribution._validate_args
        D.Distribution.set_default_validate_args(False)
        return self

    def __exit__(self, *exc):
        import torch.distributions as D
        D.Distribution.set_default_validate_args(self._was)
        return False
