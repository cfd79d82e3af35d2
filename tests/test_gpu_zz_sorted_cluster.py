"""GPU: the opt-in second-generation small sorted tree update (csrc/sumtree.cu: tree_update_sorted_v2;
MB_SORTED_CLUSTER=1: 8 subtrees on a cluster of 8 CTAs, MB_SORTED_ASYNC=1: barrier-free walk with point-to-point
hand-overs, or both) against the oracle -- bit-exact, like the single-CTA kernel."""
import numpy as np
import pytest
import torch as t


from oracle.sumtree import OracleTree

pytestmark = [pytest.mark.gpu]


@pytest.fixture(params=["cluster", "async", "cluster+async"])
def cluster_tree(monkeypatch, request):
    for name, key in (("MB_SORTED_CLUSTER", "cluster"), ("MB_SORTED_ASYNC", "async")):   # read by mb_tree_create
        monkeypatch.setenv(name, "1" if key in request.param else "0")
    from machin_b200.frame.buffers import WeightTree
    return WeightTree


@pytest.mark.parametrize("size", [1000, 1 << 20, 10_000_000])
def test_cluster_sorted_update_vs_oracle(cluster_tree, size):
    rng = np.random.RandomState(size % 89)
    leaves = rng.exponential(size=size)
    tree, ora = cluster_tree(size), OracleTree(size)
    tree.update_all_leaves(leaves)
    ora.update_all_leaves(leaves)

    def both(w, idx):
        tree.update_leaf_batch(t.from_numpy(w).cuda(), t.from_numpy(idx).cuda())
        ora.update_leaf_batch(w, idx)
        assert np.array_equal(tree.weights, ora.weights)
        assert tree.max_leaf == ora.max_leaf

    for n in (65, 96, 200, 511, 512):
        for dtype in (np.float32, np.float64):
            idx = np.sort(rng.randint(0, size, size=n)).astype(np.int64)
            idx[n // 2] = idx[n // 2 - 1]                       # a run of equal keys: last position wins
            both(rng.exponential(size=n).astype(dtype), np.sort(idx))
    # all entries in ONE of the 8 subtrees (7 CTAs only publish their stored root), first and last subtree
    span = max(1, size // 16)
    for lo in (0, size - span):
        idx = np.sort(rng.randint(lo, lo + span, size=300)).astype(np.int64)
        both(rng.exponential(size=300), idx)
    # neighbouring leaves (almost all ancestors shared) across a subtree boundary when the tree is large enough
    mid = (1 << (tree.depth - 2)) if size > (1 << (tree.depth - 2)) + 300 else size // 2
    idx = np.arange(max(0, mid - 256), min(size, mid + 256), dtype=np.int64)
    both(rng.exponential(size=len(idx)), idx)
    # a new maximum, float32: max_leaf dtype follows the batch
    idx = np.sort(rng.choice(size, size=128, replace=False)).astype(np.int64)
    w = rng.exponential(size=128).astype(np.float32)
    w[17] = 1e3
    both(w, idx)
    # unsorted input falls back to the generic body inside the same launch
    idx = rng.randint(0, size, size=400).astype(np.int64)
    both(rng.exponential(size=400), idx)
    # priority write-back (normalisation on the device), as the learner calls it
    from machin_b200 import _lib
    from machin_b200._lib import lib
    idx = np.sort(rng.randint(0, size, size=512)).astype(np.int64)
    p = rng.exponential(size=512).astype(np.float32)
    d_i, d_p = t.from_numpy(idx).cuda(), t.from_numpy(p).cuda()
    _lib.check(lib.mb_tree_update_priority(tree.handle, _lib.ptr(d_i), _lib.ptr(d_p), _lib.MB_F32, None, 512, 1e-2, 0.6,
                                           _lib.stream_ptr()))
    tree.check()
    # device formula (csrc/sumtree.cu normalized_weight): |p| + eps in float32, pow in float64, rounded once
    x = (np.abs(p) + np.float32(1e-2)).astype(np.float32)
    leaf = np.power(x.astype(np.float64), np.float64(np.float32(0.6))).astype(np.float32)
    ora.update_leaf_batch(leaf, idx)
    assert np.array_equal(tree.weights, ora.weights)
    # out-of-range entry in a sorted batch: rejected, nothing written
    bad = np.concatenate([np.arange(100, dtype=np.int64), [size]])
    tree.update_leaf_batch(t.ones(101, dtype=t.float64).cuda(), t.from_numpy(bad).cuda())
    with pytest.raises(ValueError, match="out of boundary"):
        tree.check()
    assert np.array_equal(tree.weights, ora.weights)
