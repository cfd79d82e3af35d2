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
