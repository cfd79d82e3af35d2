"""Weight and gradient movement between learners and actors over NCCL.

Replaces, on the data path, machin/parallel/server/param_server.py:21-493 and ordered_server.py
(push: deepcopy(state_dict) -> CPU -> RPC into a versioned KV; pull: RPC fetch + load_state_dict).
Here the parameters travel as ONE flat byte buffer ``[all tensors of state_dict | version i64]`` in a
single broadcast from the pushing member; the ordered server's version chain
(ordered_server.py:112-128: a push is accepted iff the stored newest version equals ``prev_version``)
is kept: the stored newest version is replicated on every member and decided from a 16-byte all-gather.

SPMD convention: in a synchronisation round every member calls :meth:`push` or :meth:`pull`.
"""
from typing import List

import torch as t


def _state_tensors(model: t.nn.Module) -> List[t.Tensor]:
    return [v for v in model.state_dict().values() if t.is_tensor(v)]


class PushPullModelServer:
    """Versioned weight exchange (param_server.py:21-91 over ordered_server.py:100-142).

    The reference's server process keeps ``{version: state_dict}`` and accepts a push iff the newest stored
    version equals the pusher's ``prev_version`` (``ordered_server.py:112-128``); a rejected pusher pulls the
    newest parameters instead (``param_server.py:49-71``).  Here that storage is a flat device buffer
    ``[version i64 | pad | tensors...]`` replicated on every member: a synchronisation round is ONE all-gather
    of 16 bytes per member ``(pushing?, prev_version)`` -- the order in which the reference's lock would see
    the requests is member order -- and, if a push is accepted, ONE broadcast of the flat buffer from the
    accepted pusher.  Rejected pushers and pullers load from their local replica.

    SPMD convention: in a round every member calls :meth:`push` or :meth:`pull`."""

    def __init__(self, model_name: str, group, src_index: int = 0, lockstep: bool = False):
        """``lockstep``: every round has exactly one push, by member ``src_index`` whose chain is intact, and
        every other member pulls it -- the outcome is known without the all-gather and without reading anything
        back from the device (no host synchronisation; capturable in a CUDA graph)."""
        self.model_name, self.group, self.src_index = model_name, group, int(src_index)
        self.lockstep = bool(lockstep)
        self._flat = None
        self._newest = None           # version of the stored parameters; None = key not created yet

    @staticmethod
    def _layout(tensors):
        """[version i64 | pad] then every tensor at a 16-byte aligned offset."""
        offs, o = [], 16
        for x in tensors:
            offs.append(o)
            o += (x.numel() * x.element_size() + 15) // 16 * 16
        return offs, o

    def _buffer(self, nbytes):
        if self._flat is None or self._flat.numel() != nbytes:
            self._flat = t.zeros(nbytes, dtype=t.uint8, device=self.group.device)
        return self._flat

    def _views(self, flat, tensors, offs):
        return [flat[o:o + x.numel() * x.element_size()].view(x.dtype).view(x.shape) for x, o in zip(tensors, offs)]

    def _load(self, model, flat, tensors, offs):
        with t.no_grad():
            t._foreach_copy_(tensors, self._views(flat, tensors, offs))   # one multi-tensor launch

    def _round(self, model: t.nn.Module, pushing: bool):
        """One synchronisation round.  Returns (accepted member or None, flat, tensors, offs)."""
        tensors = _state_tensors(model)
        offs, total = self._layout(tensors)
        flat = self._buffer(total)
        me = self.group.index
        prev = int(getattr(model, "pp_version", 0))
        if self.lockstep:
            winner, version = self.src_index, (self._newest or 0) + 1
        else:
            meta = t.tensor([1 if pushing else 0, prev], dtype=t.int64, device=self.group.device)
            seen = self.group.all_gather_f64(meta).tolist()          # (dtype-agnostic gather) [P, 2]
            winner = next((m for m, (is_push, pv) in enumerate(seen)
                           if is_push and (self._newest is None or pv == self._newest)), None)   # :116-126
            version = None if winner is None else seen[winner][1] + 1
        if winner is not None:
            if winner == me:
                t._foreach_copy_(self._views(flat, tensors, offs), [x.detach() for x in tensors])
                flat[0:8].view(t.int64).fill_(version)
            self.group.broadcast(flat, winner)
            self._newest = version
        return winner, flat, tensors, offs

    def push(self, model: t.nn.Module, pull_on_fail=True) -> bool:
        """param_server.py:36-71: True iff the server accepted ``pp_version + 1`` on top of ``pp_version``;
        otherwise (someone else extended the chain first) the newest parameters are loaded into ``model``."""
        if not hasattr(model, "pp_version"):
            model.pp_version = 0
        winner, flat, tensors, offs = self._round(model, True)
        if winner == self.group.index:
            model.pp_version += 1
            return True
        if pull_on_fail:
            self._load(model, flat, tensors, offs)
            model.pp_version = self._newest
        return False

    def pull(self, model: t.nn.Module) -> bool:
        """param_server.py:73-91: load the newest parameters iff they are newer than ``pp_version``.
        False only while nothing has ever been pushed."""
        _, flat, tensors, offs = self._round(model, False)
        if self._newest is None:
            return False
        if self.lockstep or not hasattr(model, "pp_version") or model.pp_version < self._newest:
            self._load(model, flat, tensors, offs)
            model.pp_version = self._newest
        return True


class PushPullGradServer:
    """A3C gradient reduction (param_server.py:148-493): workers push gradients, the reduce tree sums or
    averages ``reduce_batch_size`` of them per reducer, the primary applies its optimizer and publishes
    the parameters.  Over NCCL: every worker accumulates ``pushes_per_reduce`` local pushes, one
    all-reduce combines all workers, the primary member steps and broadcasts.

    With W workers pushing k times each, ``reduce_method="sum"`` applies sum over all W*k gradients and
    ``"mean"`` their mean -- the values the reference's two-level tree produces
    (test/parallel/server/test_param_server.py:108-173)."""

    def __init__(self, server_name: str, group, primary_index: int = 0, reduce_method: str = "sum",
                 pushes_per_reduce: int = 1):
        assert reduce_method in ("mean", "sum")
        self.server_name, self.group, self.primary_index = server_name, group, int(primary_index)
        self.reduce_method, self.pushes_per_reduce = reduce_method, int(pushes_per_reduce)
        self.model = self.optimizer = self.lr_scheduler = None
        self._acc, self._count = None, 0
        self._model_server = PushPullModelServer(server_name + "_params", group, primary_index)

    def manage_model(self, model: t.nn.Module, optimizer, lr_scheduler=None):
        """Primary member only (param_server.py:207-226)."""
        self.model, self.optimizer, self.lr_scheduler = model, optimizer, lr_scheduler
        model.pp_version = 0

    def push(self, model: t.nn.Module, pull_on_fail=True) -> bool:
        """Collective.  Members without gradients (e.g. a pure server process) pass ``model=None``."""
        dev = self.group.device
        ref = model if model is not None else self.model
        params = list(ref.parameters())
        n = sum(p.numel() for p in params)
        if self._acc is None:
            self._acc = t.zeros(n + 1, dtype=t.float32, device=dev)   # last element counts pushes
        if model is not None:
            o = 0
            for p in params:
                if p.grad is not None:
                    self._acc[o:o + p.numel()] += p.grad.detach().reshape(-1).to(dev)
                o += p.numel()
            self._acc[n] += 1
        self._count += 1
        if self._count < self.pushes_per_reduce:
            return True
        self.group.all_reduce_sum(self._acc)
        total = self._acc[n].clamp(min=1)
        if self.model is not None and self.group.index == self.primary_index:
            o = 0
            for p in self.model.parameters():
                g = self._acc[o:o + p.numel()].view_as(p)
                p.grad = (g / total if self.reduce_method == "mean" else g.clone()).to(p.device)
                o += p.numel()
            self.optimizer.step()
            if self.lr_scheduler is not None:
                self.lr_scheduler.step()
            self._model_server.push(self.model)
            if model is not None and model is not self.model:
                # the primary's own worker model is a different object from the managed copy (grad_server_helper builds
                # the managed one from a creator): it receives the published parameters like every other worker's
                with t.no_grad():
                    t._foreach_copy_(_state_tensors(model), [x.detach() for x in _state_tensors(self.model)])
                model.pp_version = self.model.pp_version
        else:
            self._model_server.pull(ref)
        self._acc.zero_()
        self._count = 0
        return True

    def pull(self, model: t.nn.Module) -> bool:
        """The reduced parameters were already loaded at the end of the reducing :meth:`push`."""
        return True
