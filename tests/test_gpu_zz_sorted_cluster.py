"""GPU: the small-batch tree update kernels against the oracle, bit-exact: the default (tree_update_part: one CTA
per subtree, one launch, batches up to 4096 in any order) and the legacy variants kept for A/B timing
(MB_UPDATE_LEGACY=1: single-CTA sorted walk; MB_SORTED_CLUSTER=1: 8 subtrees on a cluster of 8 CTAs;
MB_SORTED_ASYNC=1: barrier-free walk with point-to-point hand-overs)."""
import numpy as np
import pytest
import torch as t


from oracle.sumtree import OracleTree

pytestmark = [pytest.mark.gpu]


@pytest.fixture(params=["part", "legacy", "cluster", "async", "cluster+async"])
def cluster_tree(monkeypatch, request):
    for name, key in (("MB_SORTED_CLUSTER", "cluster"), ("MB_SORTED_ASYNC", "async"), ("MB_UPDATE_LEGACY", "legacy")):   # read by mb_tree_create
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


@pytest.mark.parametrize("size", [3000, 1 << 20, 10_000_000])
def test_partitioned_update_large_unsorted_masked_batches(size):
    """tree_update_part beyond the sorted learner batch: the sharded write-back (P ascending blocks, a validity
    mask, duplicates across blocks), 4096 entries, heavy skew (one subtree gets more entries than a CTA walks out of
    shared memory), and back-to-back launches (the ticket is reset by the last CTA)."""
    from machin_b200 import _lib
    from machin_b200._lib import lib
    from machin_b200.frame.buffers import WeightTree
    rng = np.random.RandomState(size % 83)
    leaves = rng.exponential(size=size)
    tree, ora = WeightTree(size), OracleTree(size)
    tree.update_all_leaves(leaves)
    ora.update_all_leaves(leaves)

    def both(w, idx, valid=None):
        d_i, d_w = t.from_numpy(idx).cuda(), t.from_numpy(w).cuda()
        d_v = None if valid is None else t.from_numpy(valid.astype(np.uint8)).cuda()
        _lib.check(lib.mb_tree_update_batch(tree.handle, _lib.ptr(d_i), _lib.ptr(d_w),
                                            _lib.MB_F32 if w.dtype == np.float32 else _lib.MB_F64, _lib.ptr(d_v), len(idx),
                                            _lib.stream_ptr()))
        tree.check()
        if valid is None:
            ora.update_leaf_batch(w, idx)
        else:
            ora.update_leaf_batch(w[valid], idx[valid])
        assert np.array_equal(tree.weights, ora.weights)
        assert tree.max_leaf == ora.max_leaf

    for blocks in (2, 4, 8):
        n = 512 * blocks
        idx = np.concatenate([np.sort(rng.randint(0, size, size=512)) for _ in range(blocks)]).astype(np.int64)
        idx[700] = idx[3]
        valid = rng.rand(n) < 0.85
        idx[np.flatnonzero(~valid)[:4]] = size + 7                 # masked-out entries may be out of range
        both(rng.exponential(size=n).astype(np.float32), idx, valid)
    span = max(8, size >> 7)
    idx = np.concatenate([rng.randint(size - span, size, size=3500), rng.randint(0, size, size=596)]).astype(np.int64)
    both(rng.exponential(size=4096), idx)
    for _ in range(5):                                             # back to back, no host sync in between
        idx = rng.randint(0, size, size=1000).astype(np.int64)
        w = rng.exponential(size=1000)
        tree.update_leaf_batch(t.from_numpy(w).cuda(), t.from_numpy(idx).cuda())
        ora.update_leaf_batch(w, idx)
    assert np.array_equal(tree.weights, ora.weights)
