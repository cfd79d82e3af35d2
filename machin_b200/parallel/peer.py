"""Peer-mapped (symmetric) device buffers over NVLink / NVSwitch for the shard exchange.

Replaces, for the fixed-shape APEX update, the NCCL collectives that stood in for the reference's per-member RPCs
(prioritized_buffer_d.py:221-245 sample, :184-202 priority write-back; apex.py:141-150 weights / gradients): every
member allocates the same buffers, each buffer is mapped into every process ONCE at set-up, and from then on this
library's kernels (csrc/peer.cu) store straight into the other GPUs' memory; one small barrier kernel per exchange
orders producers and consumers.  Nothing here is on the per-update path.

The mapping itself is torch's: ``torch.distributed._symmetric_memory`` (CUDA virtual-memory allocations whose handles
travel once through the process group's store) -- plumbing, like ``torch.distributed`` is for NCCL.  If it is not
available (old torch, no peer access between the devices) :func:`PeerMemory.create` raises and the caller keeps the NCCL
path.
"""
import ctypes as C
from typing import List, Tuple

import torch as t

from .. import _lib
from .._lib import lib

MAX_SITES = 8
BARRIER_TIMEOUT_NS = 20_000_000_000      # a member that does not arrive within 20 s is reported, not waited for


class PeerBuffer:
    """One symmetric allocation: ``local`` is this member's tensor, ``ptrs`` a C array of every member's device
    address of it as mapped into this process (member order; ``ptrs[self]`` is ``local`` itself)."""

    def __init__(self, local: t.Tensor, addresses: List[int], handle):
        self.local, self.handle = local, handle
        self.ptrs = (C.c_void_p * len(addresses))(*addresses)


class PeerMemory:
    def __init__(self, group, symm_mem):
        self.group, self._symm = group, symm_mem
        self.P, self.me, self.device = group.size(), group.index, group.device
        self._buffers = []
        # One flag row + one epoch counter per barrier SITE (a place in the update where the members meet): sites on
        # parallel branches of a captured graph may run in either order on different members, so they must not share
        # flags.  Flag [site, m]: "member m's pushes before its site-th barrier of the current epoch are done".
        self.flags = self.alloc((MAX_SITES, _lib.MB_MAX_PEERS), t.int64)
        self._site_flags = [(C.c_void_p * self.P)(*[int(a) + site * _lib.MB_MAX_PEERS * 8 for a in self.flags.ptrs])
                            for site in range(MAX_SITES)]
        self.state = t.zeros((MAX_SITES, 2), dtype=t.int64, device=self.device)  # per site: [epoch, timed-out]
        self.tickets = t.zeros(MAX_SITES, dtype=t.int32, device=self.device)     # per site: CTAs of the producer that are done
        t.cuda.synchronize(self.device)
        group.barrier()                                             # every member's flags are zero before anyone signals

    @classmethod
    def create(cls, group) -> "PeerMemory":
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        if group.size() > _lib.MB_MAX_PEERS:
            raise RuntimeError(f"peer exchange supports up to {_lib.MB_MAX_PEERS} members")
        pg = group.pg if group.pg is not None else dist.group.WORLD
        enable = getattr(symm_mem, "enable_symm_mem_for_group", None)
        if enable is not None:
            try:
                enable(pg.group_name)
            except Exception:       # noqa: BLE001 -- newer torch enables groups implicitly and may refuse a second call
                pass
        self = cls.__new__(cls)
        self._pg = pg
        cls.__init__(self, group, symm_mem)
        return self

    def alloc(self, shape: Tuple[int, ...], dtype) -> PeerBuffer:
        """Collective: every member allocates the same shape; returns the local tensor (zeroed) + all peer addresses."""
        local = self._symm.empty(*shape, dtype=dtype, device=self.device)
        handle = self._symm.rendezvous(local, self._pg)
        local.zero_()
        offset = int(local.data_ptr()) - int(handle.buffer_ptrs[handle.rank])
        addresses = [int(p) + offset for p in handle.buffer_ptrs]
        if len(addresses) != self.P or handle.rank != self.me:
            raise RuntimeError("symmetric memory group does not match the shard group")
        buf = PeerBuffer(local, addresses, handle)
        self._buffers.append(buf)
        return buf

    def barrier(self, site: int):
        """Capturable: one 32-thread kernel on the current stream (csrc/peer.cu: peer_barrier_kernel).  Every member
        calls the barriers of one ``site`` in the same order."""
        _lib.check(lib.mb_peer_barrier(self._site_flags[site], self.P, self.me, _lib.ptr(self.state[site]),
                                       BARRIER_TIMEOUT_NS, _lib.stream_ptr()))

    def site(self, site: int):
        """(flag pointer table, ticket, state) of one site: the arguments of the entry points that fold the barrier into
        the producing kernel (signal from its last CTA) and the consuming kernel (wait at its top)."""
        return self._site_flags[site], _lib.ptr(self.tickets[site:site + 1]), _lib.ptr(self.state[site])

    def check(self):
        """Host-side (synchronises): raise if a barrier gave up waiting for a member."""
        if int(self.state[:, 1].sum().item()) != 0:
            raise RuntimeError("peer exchange: a member did not reach a barrier within "
                               f"{BARRIER_TIMEOUT_NS / 1e9:.0f} s (crashed process?)")
