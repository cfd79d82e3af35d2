"""Process group of replay shards / learners: one process per GPU, NCCL over NVLink.

Replaces, on the data path only, the reference's ``RpcGroup`` (machin/parallel/distributed/
_world.py:598-977): named members, membership queries and a barrier keep their surface; the
registered-service RPC calls (pickled Python objects over TensorPipe/TCP, ``_world.py:950-965``)
are replaced by collectives and grouped point-to-point transfers of flat tensors.  Execution is
SPMD: every member enters ``sample_batch`` / ``update_priority`` / ``push`` / ``pull`` together.
The backend is whatever the default torch.distributed group was initialised with (``nccl`` on the
GPU box, ``gloo`` in the CPU tests of the host-side logic).
"""
from typing import List, Optional, Sequence

import torch as t
import torch.distributed as dist


class ShardGroup:
    def __init__(self, name: str = "group", ranks: Optional[Sequence[int]] = None, device: Optional[t.device] = None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun / init_process_group)")
        self.group_name = name
        world = dist.get_world_size()
        self.ranks = list(range(world)) if ranks is None else list(ranks)
        self.pg = None if ranks is None else dist.new_group(self.ranks)
        self.rank = dist.get_rank()
        backend = dist.get_backend()           # (of the default group: a rank outside ``ranks`` holds no handle to ask)
        self.device = device if device is not None else (
            t.device("cuda", t.cuda.current_device()) if backend == "nccl" else t.device("cpu"))

    # --- RpcGroup-like surface (_world.py:905-948)
    def size(self) -> int:
        return len(self.ranks)

    def is_member(self, target: str = None) -> bool:
        return (self.rank if target is None else int(target)) in self.ranks

    def get_group_name(self) -> str:
        return self.group_name

    def get_group_members(self) -> List[str]:
        return [str(r) for r in self.ranks]

    def get_cur_name(self) -> str:
        return str(self.rank)

    def barrier(self):
        dist.barrier(group=self.pg)

    def destroy(self):
        if self.pg is not None:
            dist.destroy_process_group(self.pg)

    # --- collectives used by the shard exchange
    @property
    def index(self) -> int:
        """Position of this process inside the group."""
        return self.ranks.index(self.rank)

    def all_gather_f64(self, value: t.Tensor) -> t.Tensor:
        """value: f64[k] on ``self.device`` -> f64[size, k]."""
        out = t.empty((self.size(),) + tuple(value.shape), dtype=value.dtype, device=self.device)
        dist.all_gather_into_tensor(out, value.contiguous(), group=self.pg) if value.is_cuda else \
            dist.all_gather(list(out.unbind(0)), value.contiguous(), group=self.pg)
        return out

    def all_to_all_bytes(self, send: t.Tensor, send_counts: List[int], recv_counts: List[int]) -> t.Tensor:
        """Variable all-to-all of a uint8 row buffer: ``send`` holds the rows for member 0, then member
        1, ... (``send_counts`` rows of ``send.shape[1]`` bytes each).  NCCL: one all_to_all_single;
        gloo (CPU tests): grouped isend/irecv."""
        width = send.shape[1]
        recv = t.empty((sum(recv_counts), width), dtype=send.dtype, device=send.device)
        if send.is_cuda:
            dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=recv_counts,
                                   input_split_sizes=send_counts, group=self.pg)
            return recv
        ops, so, ro = [], 0, 0
        me = self.index
        for m, peer in enumerate(self.ranks):
            s, r = send_counts[m], recv_counts[m]
            if m == me:
                recv[ro:ro + r] = send[so:so + s]
            else:
                if s:
                    ops.append(dist.P2POp(dist.isend, send[so:so + s].contiguous(), peer, group=self.pg))
                if r:
                    ops.append(dist.P2POp(dist.irecv, recv[ro:ro + r], peer, group=self.pg))
            so += s
            ro += r
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return recv

    def broadcast(self, tensor: t.Tensor, src_index: int):
        dist.broadcast(tensor, src=self.ranks[src_index], group=self.pg)

    def all_reduce_sum(self, tensor: t.Tensor):
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.pg)


class LocalGroup:
    """A group of one: the same interface without torch.distributed (single-GPU runs and tests)."""

    def __init__(self, name: str = "group", device=None):
        self.group_name, self.ranks, self.rank, self.pg = name, [0], 0, None
        self.device = t.device(device) if device is not None else t.device("cuda", t.cuda.current_device())

    def size(self):
        return 1

    def is_member(self, target=None):
        return True

    def get_group_name(self):
        return self.group_name

    def get_group_members(self):
        return ["0"]

    def get_cur_name(self):
        return "0"

    def barrier(self):
        pass

    def destroy(self):
        pass

    index = 0

    def all_gather_f64(self, value):
        return value.unsqueeze(0)

    def all_to_all_bytes(self, send, send_counts, recv_counts):
        return send

    def broadcast(self, tensor, src_index):
        pass

    def all_reduce_sum(self, tensor):
        pass
