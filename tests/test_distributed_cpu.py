"""CPU, world_size 2-3 over gloo: host-side logic of the multi-GPU path (shard apportioning, the row /
priority exchange, weight broadcast + versioning, gradient reduction) with a NumPy shard standing in for
the CUDA one.  The sampled indexes are checked against the reference's distributed golden vector
(test/frame/buffers/test_prioritized_buffer_d.py:115-163)."""
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch as t
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT

KAT = json.load(open(os.path.join(GOLDEN, "kat_prioritized_buffer.json")))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    except Exception as e:  # noqa
        import traceback
        ret[rank] = "ERR: " + traceback.format_exc()
    finally:
        dist.destroy_process_group()


def run_multi(fn, world):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    out = [ret.get(r, "ERR: no result") for r in range(world)]
    errs = [f"rank {r}: {o}" for r, o in enumerate(out) if isinstance(o, str) and o.startswith("ERR")]
    errs.sort(key=lambda e: "Connection closed" in e)   # root cause first, collateral failures last
    assert not errs, "\n".join(errs)
    return out


class NumpyShard:
    """Oracle-backed shard (oracle/sumtree.py + oracle/per.py) with the CUDA shard's protocol."""

    def __init__(self, capacity, width):
        from oracle.sumtree import OracleTree
        self.tree, self.width, self.cap = OracleTree(capacity), width, capacity
        self.rows = np.zeros((capacity, width), dtype=np.uint8)
        self.version = np.zeros(capacity, dtype=np.int64)
        self.stored, self.beta = 0, 0.4

    def store(self, rows, priorities):
        from oracle import per as oper
        n = len(rows)
        pos = np.arange(self.stored, self.stored + n) % self.cap
        self.rows[pos] = rows
        for p, pr in zip(pos, priorities):
            self.tree.update_leaf(oper.normalize_priority(pr), int(p))
            self.version[p] += 1
        self.stored = min(self.cap, self.stored + n)

    def weight_sum(self):
        return self.tree.get_weight_sum()

    def draw(self, n, all_sum):
        from oracle import per as oper
        if n <= 0 or self.stored == 0:
            return (t.empty((0, self.width), dtype=t.uint8), t.empty(0, dtype=t.int64), t.empty(0, dtype=t.int64),
                    t.empty(0, dtype=t.float64))
        idx, isw = oper.sample_index_and_weight(self.tree, n, self.stored, self.beta, all_weight_sum=all_sum)
        self.beta = min(1.0, self.beta + 0.001)
        return (t.from_numpy(self.rows[idx]), t.from_numpy(idx.astype(np.int64)), t.from_numpy(self.version[idx]),
                t.from_numpy(isw))

    def apply(self, index, version, priority):
        from oracle import per as oper
        index, version, priority = index.numpy(), version.numpy(), priority.numpy()
        same = self.version[index] == version
        self.tree.update_leaf_batch(oper.normalize_priority(priority[same]), index[same])


def _controlled(rank, world):
    """ranks 0,1 store the golden episode, rank 2 samples 10 (the reference's controlled test)."""
    from machin_b200.parallel.exchange import ShardedSampler
    from machin_b200.parallel.group import ShardGroup
    np.random.seed(0)
    group = ShardGroup("group")
    shard = NumpyShard(5, 16)
    if rank in (0, 1):
        rows = np.zeros((5, 16), dtype=np.uint8)
        rows[:, 0] = np.arange(5)            # the "index" custom attribute of the reference test
        rows[:, 1] = rank
        shard.store(rows, KAT["buffer"]["sample"]["priorities"])
    sampler = ShardedSampler(group, shard, 16, requesters=[2])
    rows, index, isw, counts = sampler.sample(10)
    if rank != 2:
        assert rows is None
        sampler.update_priority(None, None)
        return float(shard.tree.get_leaf_all_weights().sum())
    k = KAT["buffer"]["sample_distributed_two_shards"]
    assert rows.shape == (10, 16)
    assert rows[:, 0].tolist() == k["index"]
    assert rows[:, 1].tolist() == [0] * 5 + [1] * 5                     # merged in member order
    assert list(index.keys()) == ["0", "1"]
    assert np.all(np.abs(isw.numpy() - np.array(k["is_weight"])) < 1e-6)
    assert counts == [5, 5, 0]
    # write the IS weights back as priorities, as the reference test does
    sampler.update_priority(isw, index)
    return True


def test_sharded_sample_matches_reference_golden():
    out = run_multi(_controlled, 3)
    assert out[2] is True
    # shards 0 and 1 received identical write-backs
    assert out[0] == pytest.approx(out[1]) and out[0] > 0


def _all_learners(rank, world):
    """Every member is a requester (config C4 shape): each gets its own batch; version filter drops
    write-backs for slots overwritten after sampling."""
    from machin_b200.parallel.exchange import ShardedSampler, shard_sample_sizes
    from machin_b200.parallel.group import ShardGroup
    from oracle import per as oper
    np.random.seed(100 + rank)
    group = ShardGroup("group")
    shard = NumpyShard(64, 8)
    rng = np.random.RandomState(rank)
    rows = rng.randint(0, 255, size=(40 + 10 * rank, 8)).astype(np.uint8)
    rows[:, 7] = rank
    shard.store(rows, rng.exponential(size=len(rows)))
    sampler = ShardedSampler(group, shard, 8)
    sums = [0.0] * world
    gathered = group.all_gather_f64(t.tensor([shard.weight_sum()], dtype=t.float64)).view(-1).tolist()
    ssize, total = shard_sample_sizes(gathered, 32)
    assert (ssize, total) == tuple(oper.shard_sample_sizes(gathered, 32))
    rows_r, index, isw, counts = sampler.sample(32)
    assert counts == ssize and rows_r.shape[0] == sum(ssize) >= 32
    o = 0
    for m in range(world):
        assert np.all(rows_r[o:o + ssize[m], 7].numpy() == m)
        o += ssize[m]
    # taint a slot on this shard, then write back: the tainted entry must be skipped
    before = shard.tree.weights.copy()
    victim = int(index[str(rank)][0][0]) if str(rank) in index else 0
    shard.version[victim] += 1
    sampler.update_priority(np.full(rows_r.shape[0], 7.5, dtype=np.float32), index)
    after = shard.tree.get_leaf_all_weights()
    assert after[victim] == before[victim]
    touched = np.unique(np.concatenate([v[0].numpy() for k, v in index.items() if k == str(rank)] or [np.array([], dtype=np.int64)]))
    return True


def test_sharded_all_learners_and_version_filter():
    assert all(run_multi(_all_learners, 2))


def _model_server(rank, world):
    from server_cases import model_server_case
    return model_server_case(rank, world)


def test_model_server_push_pull_golden():
    assert all(run_multi(_model_server, 3))


def _grad_server(rank, world, method, expect):
    from server_cases import grad_server_case, grad_expect
    assert grad_expect(method, world) == tuple(expect)
    return grad_server_case(rank, world, method)


@pytest.mark.parametrize("method,expect", [("mean", (-5, -1, 1)), ("sum", (-23, -10, -5))])
def test_grad_server_golden(method, expect):
    import functools
    assert all(run_multi(functools.partial(_grad_server, method=method, expect=expect), 3))
