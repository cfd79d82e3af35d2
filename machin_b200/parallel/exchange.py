"""Sharded prioritized sampling: the exchange step of DistributedPrioritizedBuffer, backend-agnostic.

Follows machin/frame/buffers/prioritized_buffer_d.py:184-303:

  sample:  gather every shard's weight sum           (:221-229, was one RPC per member)
           all_weight_sum = sum + 1e-6               (:229)
           ssize_m = ceil(w_m * batch / all_sum)     (:233-234)
           shard m draws ssize_m rows with the GLOBAL sum in its IS weights, max-normalised per shard,
           N = len(local storage)                    (:236-245, :293-303 + prioritized_buffer.py:397-428)
           rows + (index, version, is_weight) travel to the requester, merged in member order (:247-266)
  update:  priorities are split by shard lengths and routed back with (index, version); a shard keeps
           only entries whose version still matches   (:184-202, :282-291)

The transport is one variable all-to-all of fixed-width byte rows
``[packed transition row | index i64 | version i64 | is_weight f64]`` per direction.
A *shard* is any object with ``weight_sum()``, ``draw(n, all_weight_sum) -> (rows u8[n,w], index,
version, is_weight)`` and ``apply(index, version, priority)``; the CUDA shard lives in
frame/buffers/prioritized_buffer_d.py, a NumPy one backs the CPU (gloo) tests.
"""
from collections import OrderedDict
from typing import List, Sequence, Tuple

import numpy as np
import torch as t

META = 24  # bytes appended to every row: index i64, version i64, is_weight f64


def shard_sample_sizes(sums: Sequence[float], batch_size: int) -> Tuple[List[int], float]:
    total = float(sum(float(s) for s in sums)) + 1e-6                                # :229  prevent all zero
    return [int(x) for x in np.ceil(np.array(sums, dtype=np.float64) * batch_size / total)], total   # :233-234


def _append_meta(rows: t.Tensor, index: t.Tensor, version: t.Tensor, is_weight: t.Tensor) -> t.Tensor:
    n = rows.shape[0]
    meta = t.stack([index.view(t.uint8).view(n, 8), version.view(t.uint8).view(n, 8),
                    is_weight.view(t.uint8).view(n, 8)], dim=1).view(n, META) if n else \
        t.empty((0, META), dtype=t.uint8, device=rows.device)
    return t.cat([rows, meta], dim=1)


def _split_meta(buf: t.Tensor):
    n, w = buf.shape[0], buf.shape[1] - META
    meta = buf[:, w:].contiguous()
    index = meta[:, 0:8].contiguous().view(t.int64).view(n)
    version = meta[:, 8:16].contiguous().view(t.int64).view(n)
    is_weight = meta[:, 16:24].contiguous().view(t.float64).view(n)
    return buf[:, :w], index, version, is_weight


class ShardedSampler:
    def __init__(self, group, shard, row_width: int, requesters: Sequence[int] = None):
        """``requesters``: group positions that receive a batch (default: every member, i.e. every
        process is a learner that also holds a shard, BASELINE config C4)."""
        self.group, self.shard, self.row_width = group, shard, int(row_width)
        self.requesters = list(range(group.size())) if requesters is None else list(requesters)
        self._last_served = 0

    # ------------------------------------------------------------------ sample
    def sample(self, batch_size: int):
        """Collective.  Returns (rows u8[n, row_width (+ META on the CUDA path)], index
        OrderedDict{member: (index, version)}, is_weight f64[n], counts per shard) on requesters;
        (None, None, None, None) elsewhere."""
        if hasattr(self.shard, "serve_into"):
            return self._sample_wire(batch_size)
        g = self.group
        dev = g.device
        me = g.index
        P = g.size()
        local_sum = t.tensor([float(self.shard.weight_sum())], dtype=t.float64, device=dev)
        sums = g.all_gather_f64(local_sum).view(-1).tolist()
        ssize, total = shard_sample_sizes(sums, batch_size)
        if batch_size <= 0:
            ssize = [0] * P
        # this shard serves every requester in member order, one independent draw each (one RPC each, :236-245)
        parts, send_counts = [], [0] * P
        for r in self.requesters:
            n = ssize[me]
            rows, index, version, isw = self.shard.draw(n, total)
            parts.append(_append_meta(rows, index, version, isw))
            send_counts[r] = n
        width = self.row_width + META
        send = t.cat(parts, dim=0) if parts else t.empty((0, width), dtype=t.uint8, device=dev)
        recv_counts = [ssize[m] if me in self.requesters else 0 for m in range(P)]
        self._last_served = ssize[me]
        recv = g.all_to_all_bytes(send, send_counts, recv_counts)
        if me not in self.requesters or recv.shape[0] == 0:
            return None, None, None, None
        rows, index, version, isw = _split_meta(recv)
        all_index, o = OrderedDict(), 0
        for m, name in enumerate(g.get_group_members()):
            if recv_counts[m]:
                all_index[name] = (index[o:o + recv_counts[m]], version[o:o + recv_counts[m]])
                o += recv_counts[m]
        return rows, all_index, isw, recv_counts

    # CUDA shards: payload and metadata are written straight into one wire buffer by two kernels per
    # requester (gather with a wide output pitch + pack_meta); nothing is concatenated or re-copied.
    def _sample_wire(self, batch_size: int):
        g, dev = self.group, self.group.device
        me, P = g.index, g.size()
        # the root of the local tree is gathered straight from device memory: the only host
        # synchronisation of the exchange is reading the P sums back for the ceil apportioning
        sums = g.all_gather_f64(self.shard.weight_sum_tensor()).view(-1).tolist()
        ssize, total = shard_sample_sizes(sums, batch_size)
        if batch_size <= 0:
            ssize = [0] * P
        W = self.row_width + META
        n_mine = ssize[me]
        wire = t.empty((n_mine * len(self.requesters), W), dtype=t.uint8, device=dev)
        send_counts, o = [0] * P, 0
        for r in self.requesters:                      # one independent draw per requester, member order
            self.shard.serve_into(n_mine, total, wire[o:o + n_mine])
            send_counts[r] = n_mine
            o += n_mine
        recv_counts = [ssize[m] if me in self.requesters else 0 for m in range(P)]
        self._last_served = n_mine
        recv = g.all_to_all_bytes(wire, send_counts, recv_counts)
        if me not in self.requesters or recv.shape[0] == 0:
            self._last = None
            return None, None, None, None
        index, version, isw = self.shard.unpack_meta(recv, self.row_width)
        all_index, o = OrderedDict(), 0
        for m, name in enumerate(g.get_group_members()):
            if recv_counts[m]:
                all_index[name] = (index[o:o + recv_counts[m]], version[o:o + recv_counts[m]])
                o += recv_counts[m]
        self._last = (all_index, index, version, recv_counts)
        return recv, all_index, isw, recv_counts

    def _update_wire(self, priorities, indexes):
        g, dev = self.group, self.group.device
        P, me = g.size(), g.index
        send_counts = [0] * P
        if indexes is not None:
            last = getattr(self, "_last", None)
            if last is not None and last[0] is indexes:          # the dict handed out by sample(): no re-packing
                index, version, send_counts = last[1], last[2], list(last[3])
            else:
                names = g.get_group_members()
                as_i64 = lambda x: (x if t.is_tensor(x) else t.as_tensor(np.asarray(x).astype(np.int64))).to(dev).to(t.int64)
                index = t.cat([as_i64(indexes[nm][0]) for nm in names if nm in indexes])
                version = t.cat([as_i64(indexes[nm][1]) for nm in names if nm in indexes])
                send_counts = [len(indexes[nm][0]) if nm in indexes else 0 for nm in names]
            pri = (priorities if t.is_tensor(priorities) else t.as_tensor(np.asarray(priorities))).to(dev).reshape(-1)
            send = self.shard.pack_update(index, version, pri.to(t.float32).contiguous())
        else:
            send = t.empty((0, META), dtype=t.uint8, device=dev)
        served = self._last_served
        recv_counts = [served if r in self.requesters else 0 for r in range(P)]
        recv = g.all_to_all_bytes(send, send_counts, recv_counts)
        if recv.shape[0]:
            self.shard.apply_wire(recv)

    # ------------------------------------------------------------------ priority write-back
    def update_priority(self, priorities, indexes: "OrderedDict"):
        """Collective.  ``priorities``: tensor/array [n] in merged order; ``indexes``: the OrderedDict
        returned by :meth:`sample` (None on non-requesters)."""
        if hasattr(self.shard, "serve_into"):
            return self._update_wire(priorities, indexes)
        g = self.group
        dev = g.device
        P = g.size()
        me = g.index
        names = g.get_group_members()
        send_counts = [0] * P
        parts = []
        if indexes is not None:
            pri = priorities if t.is_tensor(priorities) else t.as_tensor(np.asarray(priorities))
            pri = pri.to(dev).reshape(-1).to(t.float32)
            o = 0
            for m, name in enumerate(names):
                if name in indexes:
                    idx, ver = indexes[name]
                    idx = (idx if t.is_tensor(idx) else t.as_tensor(np.asarray(idx))).to(dev).to(t.int64)
                    ver = (ver if t.is_tensor(ver) else t.as_tensor(np.asarray(ver).astype(np.int64))).to(dev).to(t.int64)
                    n = idx.shape[0]
                    p64 = pri[o:o + n].to(t.float64)      # carried as f64 on the wire, exact for f32 values
                    parts.append(_append_meta(t.empty((n, 0), dtype=t.uint8, device=dev), idx, ver, p64))
                    send_counts[m] = n
                    o += n
        send = t.cat(parts, dim=0) if parts else t.empty((0, META), dtype=t.uint8, device=dev)
        # how many entries each requester sends me = how many rows I served it in the matching sample()
        # (every member computed the same apportioning, so no extra collective / host sync is needed)
        served = self._last_served
        recv_counts = [served if r in self.requesters else 0 for r in range(P)]
        recv = g.all_to_all_bytes(send, send_counts, recv_counts)
        if recv.shape[0]:
            _, idx, ver, p64 = _split_meta(recv)
            self.shard.apply(idx, ver, p64.to(t.float32))
