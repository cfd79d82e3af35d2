"""CPU, world_size 2-4 over gloo: IMPALA's episode sampling across replay shards (frame/buffers/buffer_d.py:
DistributedBuffer.collect / IMPALABuffer -- impala.py:28-66 + buffer_d.py:131-198) with the product's own Python over
a host-memory test double of the device ring (tests/fake_ring.py).  Every member draws ceil(B / P) whole episodes
with Python's random (same stream as the reference) and ships rows + episode lengths to the learner (member 0);
the learner's batch is compared with a replay of every member's draw.  The same scenario runs over NCCL in
tests/test_gpu_zz_impala_exchange.py."""
import random
from threading import RLock

import numpy as np
import pytest
import torch as t

from test_distributed_cpu import run_multi


def episodes_of(rank):
    rng = np.random.RandomState(50 + rank)
    eps = []
    for e in range(9 + rank):
        n = int(rng.randint(1, 8))
        eps.append([dict(state={"s": t.full((1, 3), float(1000 * rank + 10 * e + i))}, action={"a": t.tensor([[rank]])},
                         next_state={"s": t.zeros(1, 3)}, reward=float(e), terminal=(i == n - 1),
                         tag=1000 * rank + 10 * e + i) for i in range(n)])
    return eps


def _impala_member(rank, world):
    from fake_ring import make_cpu_buffer
    from machin_b200.frame.buffers import IMPALABuffer
    from machin_b200.parallel import ShardGroup
    group = ShardGroup("group")
    buf = make_cpu_buffer(IMPALABuffer, 500, buffer_name="impala", group=group, requesters=[0], wr_lock=RLock(),
                          _schema_agreed=False)
    if rank > 0:
        for ep in episodes_of(rank):
            buf.store_episode(ep)
    B = 10
    local_b = int(np.ceil(B / world))
    for round_ in range(3):
        random.seed(7 * rank + round_)
        units, fields, lengths = buf.sample_episodes_device(B)
        if rank != 0:
            assert units == 0 and fields is None
            continue
        want_tags, want_len = [], []
        for m in range(world):
            eps = episodes_of(m) if m > 0 else []
            random.seed(7 * m + round_)
            chosen = random.choices(list(range(len(eps))), k=min(len(eps), local_b))
            for e in chosen:
                want_len.append(len(eps[e]))
                want_tags += [1000 * m + 10 * e + i for i in range(len(eps[e]))]
        assert units == len(want_len) and lengths == want_len, (units, lengths, want_len)
        assert fields["tag"].view(-1).tolist() == want_tags
        assert fields["state/s"][:, 0].tolist() == [float(x) for x in want_tags]
        ends = np.cumsum(want_len) - 1
        assert [i for i, x in enumerate(fields["terminal"].view(-1).tolist()) if x] == ends.tolist()
    random.seed(99 + rank)
    n, batch = buf.sample_batch(B, sample_attrs=["state", "reward", "tag"], additional_concat_custom_attrs=["tag"])
    if rank == 0:
        assert n > 0 and batch[0]["s"].shape[1] == 3 and batch[1].shape[0] == batch[2].shape[0] == batch[0]["s"].shape[0]
    else:
        assert n == 0 and batch is None
    assert buf.all_size() == sum(sum(len(e) for e in episodes_of(m)) for m in range(1, world))
    return "ok"


@pytest.mark.parametrize("world", [2, 4])
def test_impala_episode_exchange_over_gloo(world):
    assert run_multi(_impala_member, world) == ["ok"] * world


def _episode5():
    return [{"state": {"state_1": t.zeros([1, 2])}, "action": {"action_1": t.ones([1, 3])},
             "next_state": {"next_state_1": t.zeros([1, 2])}, "reward": 1.5, "terminal": True}] * 5


def _dist_buffer(group, requesters=None):
    from fake_ring import make_cpu_buffer
    from machin_b200.frame.buffers import DistributedBuffer
    return make_cpu_buffer(DistributedBuffer, 5, buffer_name="buffer", group=group, wr_lock=RLock(), _schema_agreed=False,
                           requesters=list(range(group.size())) if requesters is None else requesters)


def _distributed_buffer_cases(rank, world):
    """The three cases of the reference's test_buffer_d.py:16-110 (store + sample, size / all_size, clear / all_clear),
    SPMD: members 0 and 1 hold episodes, member 2 samples; collectives are entered by every member."""
    from machin_b200.parallel import ShardGroup
    group = ShardGroup("group")
    # ---- store_episode and sample_batch
    buffer = _dist_buffer(group, requesters=[2])
    if rank in (0, 1):
        buffer.store_episode(_episode5())
    random.seed(rank)
    batch_size, sample = buffer.sample_batch(7)
    if rank == 2:
        assert batch_size > 0 and len(sample) == 5
        assert t.all(sample[0]["state_1"] == 0) and list(sample[0]["state_1"].shape) == [batch_size, 2]
        assert t.all(sample[1]["action_1"] == 1) and list(sample[1]["action_1"].shape) == [batch_size, 3]
        assert t.all(sample[3] == 1.5) and list(sample[3].shape) == [batch_size, 1]
        assert t.all(sample[4]) and list(sample[4].shape) == [batch_size, 1]
        assert batch_size == 2 * 3                       # ceil(7 / 3) from each of the two members that hold data
    else:
        assert batch_size == 0 and sample is None
    # ---- size and all_size
    buffer = _dist_buffer(group)
    if rank == 0:
        buffer.store_episode(_episode5())
    assert buffer.size() == (5 if rank == 0 else 0)
    assert buffer.all_size() == 5
    # ---- clear and all_clear
    buffer = _dist_buffer(group)
    if rank in (0, 1):
        buffer.store_episode(_episode5())
    if rank == 0:
        buffer.clear()
        assert buffer.size() == 0
    assert buffer.all_size() == 5
    buffer.all_clear()
    assert buffer.all_size() == 0
    return "ok"


def test_distributed_buffer_reference_cases_over_gloo():
    assert run_multi(_distributed_buffer_cases, 3) == ["ok"] * 3
