"""Sharded prioritized replay: one tree + one row ring per GPU, exchanged over NCCL.

Mirror of machin/frame/buffers/prioritized_buffer_d.py:10-303 (DistributedPrioritizedBuffer).  Each
process owns a local shard (sum tree, packed rows, version table -- all in its GPU's HBM).  The
reference reaches the other shards with per-member RPCs carrying pickled ``Transition`` objects over
TCP (:221-245); here every member enters :meth:`sample_batch` / :meth:`update_priority` together and
the shards exchange flat byte rows with one variable all-to-all per direction
(:mod:`machin_b200.parallel.exchange`).  Reference semantics kept: ``+1e-6`` on the global sum, ``ceil``
apportioning (the merged batch may exceed ``batch_size``), per-shard max-normalised IS weights with
``N = len(local storage)``, member-order merge, version ("taint") filter on late priority write-backs,
re-normalised default priorities that compound inside an episode (:123-131).
"""
from collections import OrderedDict
from threading import RLock
from typing import Dict, List, Union

import numpy as np
import torch as t

from ... import _lib
from ..._lib import lib
from ...parallel.exchange import ShardedSampler
from ..transition import TransitionBase
from .buffer import Buffer
from .prioritized_buffer import PrioritizedBuffer


def compounding_default_weights(max_leaf, n: int, epsilon: float, alpha: float) -> np.ndarray:
    """Leaf weights the reference assigns to an episode stored WITHOUT priorities in the distributed
    buffer (prioritized_buffer_d.py:123-131): one ``update_leaf(normalize(get_leaf_max()))`` per transition,
    with the running maximum re-read every time -- it is already normalised, gets normalised again and can
    grow inside the episode.  ``max_leaf`` keeps its dtype (np.float32 after a float32 write-back)."""
    cur, weights = max_leaf, []
    for _ in range(n):
        w = (np.abs(cur) + epsilon) ** alpha
        weights.append(w)
        cur = max(w, cur)              # update_leaf: self.max_leaf = max(weight, self.max_leaf)
    return np.array(weights)


class DistributedPrioritizedBuffer(PrioritizedBuffer):
    def __init__(self, buffer_name: str, group, buffer_size: int = 1000000, epsilon: float = 1e-2,
                 alpha: float = 0.6, beta: float = 0.4, beta_increment_per_sampling: float = 0.001,
                 requesters: List[int] = None, **kwargs):
        kwargs.pop("buffer_device", None)
        super().__init__(buffer_size=buffer_size, buffer_device=group.device, epsilon=epsilon, alpha=alpha,
                         beta=beta, beta_increment_per_sampling=beta_increment_per_sampling, **kwargs)
        assert group.is_member()
        self.buffer_name = buffer_name
        self.group = group
        self.buffer_version_table = t.zeros(int(buffer_size), dtype=t.int64, device=self.buffer_device)
        self.wr_lock = RLock()
        self._requesters = requesters
        self._sampler = None

    # ------------------------------------------------------------------ shard protocol (parallel/exchange.py)
    def weight_sum(self) -> float:
        with self.wr_lock:
            return self.wt_tree.get_weight_sum()                     # _weight_sum_service (:278-280)

    def weight_sum_tensor(self) -> t.Tensor:
        """f64[1] CUDA view of the tree root (no host round trip)."""
        if getattr(self, "_root_view", None) is None:
            self._root_view = self.wt_tree._view(lib.mb_tree_root_ptr(self.wt_tree.handle), (1,), "<f8")
        return self._root_view

    def draw(self, n: int, all_weight_sum: float):
        with self.wr_lock:                                           # _sample_service (:293-303)
            dev = self.buffer_device
            if n <= 0 or len(self.storage) == 0:
                pitch = self.storage.pitch
                return (t.empty((0, pitch), dtype=t.uint8, device=dev), t.empty(0, dtype=t.int64, device=dev),
                        t.empty(0, dtype=t.int64, device=dev), t.empty(0, dtype=t.float64, device=dev))
            d_index, d_isw = self._draw(n, all_weight_sum)
            version = self.buffer_version_table[d_index]
            return self.storage.gather_packed(d_index), d_index, version, d_isw

    # wire protocol (parallel/exchange.py::_sample_wire): payload + metadata written in place
    def serve_into(self, n: int, all_weight_sum: float, wire: t.Tensor):
        with self.wr_lock:
            if n <= 0:
                return
            d_index, d_isw = self._draw(n, all_weight_sum)
            self.storage.gather_into_wire(d_index, wire)
            _lib.check(lib.mb_rows_pack_meta(_lib.ptr(wire), int(wire.shape[1]), self.storage.pitch, n, _lib.ptr(d_index),
                                             None, _lib.ptr(self.buffer_version_table), _lib.ptr(d_isw), None,
                                             _lib.stream_ptr()))

    def unpack_meta(self, recv: t.Tensor, row_width: int):
        n, dev = int(recv.shape[0]), recv.device
        index = t.empty(n, dtype=t.int64, device=dev)
        version = t.empty(n, dtype=t.int64, device=dev)
        isw = t.empty(n, dtype=t.float64, device=dev)
        _lib.check(lib.mb_rows_unpack_meta(_lib.ptr(recv), int(recv.shape[1]), row_width, n, _lib.ptr(index),
                                           _lib.ptr(version), _lib.ptr(isw), None, _lib.stream_ptr()))
        return index, version, isw

    def pack_update(self, index: t.Tensor, version: t.Tensor, priority_f32: t.Tensor) -> t.Tensor:
        n = int(index.shape[0])
        send = t.empty((n, 24), dtype=t.uint8, device=index.device)
        _lib.check(lib.mb_rows_pack_meta(_lib.ptr(send), 24, 0, n, _lib.ptr(index.contiguous()), _lib.ptr(version.contiguous()),
                                         None, None, _lib.ptr(priority_f32), _lib.stream_ptr()))
        return send

    def apply_wire(self, recv: t.Tensor):
        n, dev = int(recv.shape[0]), recv.device
        index = t.empty(n, dtype=t.int64, device=dev)
        version = t.empty(n, dtype=t.int64, device=dev)
        pri = t.empty(n, dtype=t.float32, device=dev)
        _lib.check(lib.mb_rows_unpack_meta(_lib.ptr(recv), 24, 0, n, _lib.ptr(index), _lib.ptr(version), None,
                                           _lib.ptr(pri), _lib.stream_ptr()))
        self.apply(index, version, pri)

    def apply(self, index: t.Tensor, version: t.Tensor, priority: t.Tensor):
        with self.wr_lock:                                           # _update_priority_service (:282-291)
            is_same = self.buffer_version_table[index] == version
            PrioritizedBuffer.update_priority(self, priority, index, valid=is_same)

    def _exchange(self) -> ShardedSampler:
        if self._sampler is None:
            self._agree_on_schema()
            self._sampler = ShardedSampler(self.group, self, self.storage.pitch, self._requesters)
        return self._sampler

    def _agree_on_schema(self):
        """Collective, once: members that never stored anything (pure learners) adopt the row layout of
        the first member that has one, so that every shard packs / unpacks the same byte rows."""
        if self.group.size() == 1:
            if self.storage.data is None:
                raise RuntimeError("the buffer is empty: store an episode before sampling")
            return
        import torch.distributed as dist
        from .device_ring import Field
        mine = None
        if self.storage.data is not None:
            mine = ([(f.name, f.attr, f.key, str(f.dtype), tuple(f.shape), f.kind) for f in self.storage.fields.values()],
                    self._major_attr, self._sub_attr, self._custom_attr)
        specs = [None] * self.group.size()
        dist.all_gather_object(specs, mine, group=self.group.pg)
        first = next((s for s in specs if s is not None), None)
        if first is None:
            raise RuntimeError("no member of the group has stored a transition yet")
        for s in specs:
            if s is not None and s[0] != first[0]:
                raise RuntimeError("members of the group store transitions with different layouts")
        if mine is None:
            fields = [Field(n, a, k, getattr(t, d.split(".")[-1]), sh, kind) for n, a, k, d, sh, kind in first[0]]
            self._major_attr, self._sub_attr, self._custom_attr = list(first[1]), list(first[2]), list(first[3])
            self.storage.set_schema(fields)

    # ------------------------------------------------------------------ store (:111-137)
    def store_episode(self, episode: List[Union[TransitionBase, Dict]], priorities: Union[List[float], None] = None,
                      required_attrs=("state", "action", "next_state", "reward", "terminal")):
        with self.wr_lock:
            positions = Buffer.store_episode(self, episode, required_attrs)
            self._set_new_priorities(positions, priorities)
            return positions

    def store_episode_tensors(self, major, sub, custom=None, priorities=None):
        with self.wr_lock:
            positions = Buffer.store_episode_tensors(self, major, sub, custom)
            if t.is_tensor(priorities) and priorities.is_cuda:
                PrioritizedBuffer.update_priority(self, priorities, t.as_tensor(positions).to(self.buffer_device))
                self.buffer_version_table[t.as_tensor(positions).to(self.buffer_device)] += 1
            else:
                self._set_new_priorities(positions, None if priorities is None else list(np.asarray(priorities)))
            return positions

    def _set_new_priorities(self, positions, priorities):
        if priorities is None:
            # one update_leaf per transition with normalize(get_leaf_max()) re-read every time (:123-131):
            # the (already normalised) maximum is normalised again and can grow inside the episode
            weights = compounding_default_weights(self.wt_tree.get_leaf_max(), len(positions), self.epsilon, self.alpha)
        else:
            weights = np.array([self._normalize_priority(p) for p in priorities])
        self.wt_tree.update_leaf_batch(weights, positions)
        # taint: later write-backs for the overwritten slots are ignored (:131,137)
        self.buffer_version_table[t.as_tensor(np.asarray(positions)).to(self.buffer_device)] += 1

    def size(self):
        with self.wr_lock:
            return super().size()

    def all_size(self):
        sizes = self.group.all_gather_f64(t.tensor([float(self.size())], dtype=t.float64, device=self.group.device))
        return int(sizes.sum().item())

    def clear(self):
        with self.wr_lock:
            super().clear()
            self.buffer_version_table.zero_()

    def all_clear(self):
        self.clear()          # SPMD: every member calls all_clear()
        self.group.barrier()

    # ------------------------------------------------------------------ sample / update (collective)
    def sample_device(self, batch_size: int, names=None):
        """Collective device path: (n, {field: tensor}, index OrderedDict, is_weight f64[n]) on
        requesters, (0, None, None, None) elsewhere."""
        rows, index, isw, _ = self._exchange().sample(batch_size)
        if rows is None:
            return 0, None, None, None
        return int(rows.shape[0]), self.storage.unpack(rows, names=names), index, isw

    def sample_batch(self, batch_size: int, concatenate: bool = True, device: Union[str, t.device] = None,
                     sample_attrs: List[str] = None, additional_concat_custom_attrs: List[str] = None, *_, **__):
        if batch_size <= 0:
            return 0, None, None, None
        rows, index, isw, _ = self._exchange().sample(batch_size)
        if rows is None:
            return 0, None, None, None
        result = self.post_process_batch(None, device, concatenate, sample_attrs, additional_concat_custom_attrs,
                                         packed_rows=rows)
        host_index = OrderedDict((m, (i.cpu().numpy(), v.cpu().numpy().astype(np.uint64))) for m, (i, v) in index.items())
        return int(rows.shape[0]), result, host_index, isw.cpu().numpy()

    def update_priority(self, priorities, indexes: "OrderedDict"):
        """Collective: requesters pass what :meth:`sample_batch` returned; other members pass ``None``."""
        self._exchange().update_priority(priorities, indexes)
