"""Uniform / whole-episode sampling across replay shards over NCCL.

Mirror of machin/frame/buffers/buffer_d.py:14-198 (DistributedBuffer) and of IMPALABuffer
(machin/frame/algorithms/impala.py:28-66).  Every member samples ``ceil(batch_size / members)`` units
locally with the requested sample method (:131-147, :180-198) -- Python's ``random`` stream is consumed
exactly like the reference -- and ships the packed rows to the requesting member(s) in member order.
The reference returns lists of pickled ``Transition`` objects over RPC; here it is one variable
all-to-all of byte rows plus a small all-to-all of unit sizes (episode lengths).
"""
import random
from threading import RLock
from typing import Any, Callable, List, Tuple, Union

import numpy as np
import torch as t

from .buffer import Buffer


def _round_up(num):
    return int(np.ceil(num))


class DistributedBuffer(Buffer):
    def __init__(self, buffer_name: str, group, buffer_size: int = 1000000, requesters: List[int] = None, **kwargs):
        kwargs.pop("buffer_device", None)
        super().__init__(buffer_size=buffer_size, buffer_device=group.device, **kwargs)
        assert group.is_member()
        self.buffer_name, self.group = buffer_name, group
        self.requesters = list(range(group.size())) if requesters is None else list(requesters)
        self.wr_lock = RLock()
        self._schema_agreed = False

    def store_episode(self, episode, required_attrs=("state", "action", "next_state", "reward", "terminal")):
        with self.wr_lock:
            return super().store_episode(episode, required_attrs=required_attrs)

    def clear(self):
        with self.wr_lock:
            return super().clear()

    def all_clear(self):
        self.clear()
        self.group.barrier()

    def size(self):
        with self.wr_lock:
            return super().size()

    def all_size(self):
        sizes = self.group.all_gather_f64(t.tensor([float(self.size())], dtype=t.float64, device=self.group.device))
        return int(sizes.sum().item())

    def _agree_on_schema(self):
        if self._schema_agreed or self.group.size() == 1:
            self._schema_agreed = True
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
        if mine is None:
            fields = [Field(n, a, k, getattr(t, d.split(".")[-1]), sh, kind) for n, a, k, d, sh, kind in first[0]]
            self._major_attr, self._sub_attr, self._custom_attr = list(first[1]), list(first[2]), list(first[3])
            self.storage.set_schema(fields)
        self._schema_agreed = True

    def _local_sample(self, local_batch_size, sample_method):
        """-> (units, handles, unit_sizes): _sample_service (:180-198)."""
        if isinstance(sample_method, str):
            if not hasattr(self, "sample_method_" + sample_method):
                raise RuntimeError(f"Cannot find specified sample method: {sample_method}")
            fn = getattr(self, "sample_method_" + sample_method)
            with self.wr_lock:
                out = fn(local_batch_size)
        else:
            with self.wr_lock:
                out = sample_method(self, local_batch_size)
        units, handles = out[0], out[1]
        sizes = out[2] if len(out) > 2 else [1] * len(handles)
        return units, handles, sizes

    def collect(self, batch_size: int, sample_method="random_unique"):
        """Collective.  -> (units, rows u8[n, pitch], unit_sizes list) on requesters, (0, None, None) elsewhere."""
        self._agree_on_schema()
        g, dev = self.group, self.buffer_device
        P, me = g.size(), g.index
        local_batch_size = _round_up(batch_size / P)                       # :143-144
        if P == 1:                                                         # nothing to exchange: no collectives, no syncs
            units, handles, sizes = self._local_sample(local_batch_size, sample_method)
            if not len(handles):
                return 0, None, None
            idx = t.as_tensor(np.asarray(handles, dtype=np.int64)).to(dev)
            return units, self.storage.gather_packed(idx), [int(x) for x in sizes]
        parts, size_parts, send_rows, send_units = [], [], [0] * P, [0] * P
        for r in self.requesters:                                         # one local draw per requester, member order
            units, handles, sizes = self._local_sample(local_batch_size, sample_method)
            idx = t.as_tensor(np.asarray(handles, dtype=np.int64)).to(dev)
            parts.append(self.storage.gather_packed(idx) if len(handles) else
                         t.empty((0, self.storage.pitch), dtype=t.uint8, device=dev))
            size_parts.append(t.as_tensor(np.asarray(sizes, dtype=np.int64)).to(dev).view(-1, 1).view(t.uint8).view(-1, 8))
            send_rows[r], send_units[r] = len(handles), len(sizes)
        counts = t.tensor([send_rows, send_units], dtype=t.float64, device=dev).view(-1)
        matrix = g.all_gather_f64(counts).view(P, 2, P)                    # [member, rows|units, requester]
        recv_rows = [int(matrix[m, 0, me].item()) for m in range(P)]
        recv_units = [int(matrix[m, 1, me].item()) for m in range(P)]
        rows = g.all_to_all_bytes(t.cat(parts) if parts else t.empty((0, self.storage.pitch), dtype=t.uint8, device=dev),
                                  send_rows, recv_rows)
        sizes = g.all_to_all_bytes(t.cat(size_parts) if size_parts else t.empty((0, 8), dtype=t.uint8, device=dev),
                                   send_units, recv_units)
        if me not in self.requesters:
            return 0, None, None
        unit_sizes = sizes.contiguous().view(t.int64).view(-1).tolist() if sizes.shape[0] else []
        return sum(recv_units), rows.contiguous(), unit_sizes

    def sample_batch(self, batch_size: int, concatenate: bool = True, device: Union[str, t.device] = None,
                     sample_method: Union[Callable, str] = "random_unique", sample_attrs: List[str] = None,
                     additional_concat_custom_attrs: List[str] = None, *_, **__) -> Tuple[int, Union[None, tuple]]:
        units, rows, _ = self.collect(batch_size, sample_method)
        if rows is None or rows.shape[0] == 0:
            return 0, None
        return units, self.post_process_batch(None, device, concatenate, sample_attrs,
                                              additional_concat_custom_attrs or [], packed_rows=rows)


class IMPALABuffer(DistributedBuffer):
    """Samples ``batch_size`` whole EPISODES (impala.py:28-66)."""

    def sample_batch(self, batch_size: int, concatenate: bool = True, device=None, sample_attrs: List[str] = None,
                     additional_concat_custom_attrs: List[str] = None, *_, **__) -> Any:
        return super().sample_batch(batch_size=batch_size, concatenate=concatenate, device=device,
                                    sample_method="episode", sample_attrs=sample_attrs,
                                    additional_concat_custom_attrs=additional_concat_custom_attrs)

    def sample_method_episode(self, batch_size: int):
        batch_size = min(len(self.episode_transition_handles), batch_size)
        live = self.book.live_episode_range()
        if live is not None:
            # random.choices(population, k) picks population[floor(random() * n)]: choosing positions consumes the
            # stream exactly like choosing from list(keys) (impala.py:58), and position i is episode lo + i
            picks = random.choices(range(live[1] - live[0]), k=batch_size)
            handles, lens = self.book.episode_slots(live[0] + np.asarray(picks, dtype=np.int64))
            return batch_size, handles, lens.tolist()
        episodes = random.choices(list(self.episode_transition_handles.keys()), k=batch_size)   # impala.py:58
        parts = [np.asarray(self.episode_transition_handles[episode], dtype=np.int64) for episode in episodes]
        handles = np.concatenate(parts) if parts else np.empty(0, dtype=np.int64)
        return batch_size, handles, [len(p) for p in parts]

    def sample_episodes_device(self, batch_size: int, names=None):
        """Collective device path: (n_episodes, {field: tensor}, lengths list) on requesters."""
        units, rows, sizes = self.collect(batch_size, "episode")
        if rows is None or rows.shape[0] == 0:
            return 0, None, None
        return units, self.storage.unpack(rows, names=names), sizes
