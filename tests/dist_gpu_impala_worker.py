"""torchrun worker for tests/test_gpu_zz_impala_exchange.py: IMPALA's episode sampling across replay shards over
NCCL (impala.py:28-66 + buffer_d.py:131-198): every member draws ceil(B / P) whole episodes with Python's random
(same stream as the reference) and ships rows + episode lengths to the learner (member 0)."""
import os
import random
import sys

import numpy as np
import torch as t
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def episodes_of(rank):
    """Deterministic episodes of a member: lengths 1..7, every step tagged (rank, episode, step)."""
    rng = np.random.RandomState(50 + rank)
    eps = []
    for e in range(9 + rank):
        n = int(rng.randint(1, 8))
        eps.append([dict(state={"s": t.full((1, 3), float(1000 * rank + 10 * e + i))}, action={"a": t.tensor([[rank]])},
                         next_state={"s": t.zeros(1, 3)}, reward=float(e), terminal=(i == n - 1),
                         tag=1000 * rank + 10 * e + i) for i in range(n)])
    return eps


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    t.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=t.device("cuda", local))
    from machin_b200.frame.buffers import IMPALABuffer
    from machin_b200.parallel import ShardGroup
    group = ShardGroup("group")
    buf = IMPALABuffer("impala", group, buffer_size=500, requesters=[0])
    if rank > 0:                                             # actors hold the trajectories, the learner holds none
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
        # replay every member's draw: random.choices over its live episodes, in member order
        want_tags, want_len = [], []
        for m in range(world):
            eps = episodes_of(m) if m > 0 else []
            random.seed(7 * m + round_)
            k = min(len(eps), local_b)
            chosen = random.choices(list(range(len(eps))), k=k)
            for e in chosen:
                want_len.append(len(eps[e]))
                want_tags += [1000 * m + 10 * e + i for i in range(len(eps[e]))]
        assert units == len(want_len) and lengths == want_len, (units, lengths, want_len)
        assert fields["tag"].view(-1).tolist() == want_tags
        assert fields["state/s"][:, 0].tolist() == [float(x) for x in want_tags]
        term = fields["terminal"].view(-1).tolist()
        ends = np.cumsum(want_len) - 1
        assert [i for i, x in enumerate(term) if x] == ends.tolist()
    # the reference-style call: concatenated tuple on the requester, (0, None) elsewhere
    random.seed(99 + rank)
    n, batch = buf.sample_batch(B, sample_attrs=["state", "reward", "tag"], additional_concat_custom_attrs=["tag"])
    if rank == 0:
        assert n > 0 and batch[0]["s"].shape[1] == 3 and batch[1].shape[0] == batch[2].shape[0] == batch[0]["s"].shape[0]
    else:
        assert n == 0 and batch is None
    assert buf.all_size() == sum(sum(len(e) for e in episodes_of(m)) for m in range(1, world))
    dist.barrier()
    if rank == 0:
        print("impala ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
