"""CPU: the ALGORITHM of the small sorted tree update kernels (csrc/sumtree.cu: tree_update_sorted,
tree_update_sorted_v2 -- barrier walk, polling walk with point-to-point hand-overs, 8-subtree cluster split),
stepped by oracle/sorted_walk_model.py under random lane / warp scheduling, against the reference restatement
(OracleTree.update_leaf_batch): bit-identical trees, and the polling walk always terminates."""
import numpy as np
import pytest

from oracle.sorted_walk_model import update_sorted
from oracle.sumtree import OracleTree


def _case(rng, size, n, kind):
    if kind == "uniform":
        idx = rng.randint(0, size, size=n)
    elif kind == "dense":                                   # neighbouring leaves: merges on the lowest levels
        start = int(rng.randint(0, max(1, size - n)))
        idx = start + np.arange(n)
    elif kind == "one_subtree":
        span = max(1, size // 16)
        lo = int(rng.choice([0, size - span]))
        idx = rng.randint(lo, lo + span, size=n)
    else:                                                   # clumps: runs of close keys far apart + duplicates
        centres = rng.randint(0, size, size=max(1, n // 16))
        idx = np.clip(rng.choice(centres, size=n) + rng.randint(0, 5, size=n), 0, size - 1)
    return np.sort(idx).astype(np.int64)


@pytest.mark.parametrize("walk", ["barrier", "poll"])
@pytest.mark.parametrize("cluster_bits", [0, 3])
@pytest.mark.parametrize("size", [1000, 1 << 16, 100_000])
def test_walk_model_is_bit_identical_to_reference_update(walk, cluster_bits, size):
    rng = np.random.RandomState(size % 1000 + cluster_bits + (7 if walk == "poll" else 0))
    a, b = OracleTree(size), OracleTree(size)
    leaves = rng.exponential(size=size)
    a.update_all_leaves(leaves)
    b.update_all_leaves(leaves)
    worst = 0
    for kind in ("uniform", "dense", "one_subtree", "clumps"):
        for n in (1, 2, 33, 65, 200, 512):
            idx = _case(rng, size, n, kind)
            w = rng.exponential(size=n).astype(np.float32 if n % 2 else np.float64)
            a.update_leaf_batch(w, idx)
            rounds = update_sorted(b, idx, w, cluster_bits=cluster_bits, walk=walk, rng=rng)
            worst = max(worst, rounds)
            assert np.array_equal(a.weights, b.weights), (kind, n)
            assert a.max_leaf == b.max_leaf
    if walk == "poll":
        # the critical path is the chain of hand-overs, not levels x entries: a few rounds per level at most
        assert worst <= 8 * a.depth, worst
