"""GPU parity: the CUDA sum tree / PER draw (through the C ABI) against the oracle, the reference's
known-answer vectors and the reference-generated fixtures.  Bit-exact for tree contents and indexes."""
import json
import os

import numpy as np
import pytest
import torch as t

from conftest import GOLDEN
from oracle.sumtree import OracleTree

pytestmark = pytest.mark.gpu

KAT = json.load(open(os.path.join(GOLDEN, "kat_prioritized_buffer.json")))
SIZES = (1, 2, 3, 5, 8, 1000, 1025, 4096)


@pytest.fixture(scope="module")
def WT():
    from machin_b200.frame.buffers import WeightTree
    return WeightTree


def _kat_tree(WT):
    tree = WT(KAT["tree"]["size"])
    tree.update_all_leaves(KAT["tree"]["base"])
    return tree


def test_kat_tree(WT):
    k = KAT["tree"]
    tree = _kat_tree(WT)
    assert np.all(tree.weights == k["weights"])
    assert tree.get_weight_sum() == k["weights"][-1]
    assert tree.get_leaf_max() == max(k["base"])
    assert np.all(tree.get_leaf_all_weights() == k["base"])
    for q, expect in k["find"]:
        got = tree.find_leaf_index(q)
        assert np.all(got == expect)
        assert isinstance(got, int) == (not isinstance(q, list))
    for idx, expect in k["leaf_weight"]:
        assert np.all(tree.get_leaf_weight(idx) == expect)
    for bad in k["leaf_weight_oob"]:
        with pytest.raises(ValueError, match=k["messages"]["oob"]):
            tree.get_leaf_weight(bad)
    for bad in k["update_leaf"]["oob"]:
        with pytest.raises(ValueError, match=k["messages"]["oob"]):
            tree.update_leaf(2.0, bad)
    tree.update_leaf(k["update_leaf"]["weight"], k["update_leaf"]["index"])
    assert np.all(tree.weights == k["update_leaf"]["weights"])
    tree = _kat_tree(WT)
    tree.update_leaf_batch(k["update_leaf_batch"]["weights_in"], k["update_leaf_batch"]["indexes"])
    assert np.all(tree.weights == k["update_leaf_batch"]["weights"])
    with pytest.raises(ValueError, match=k["messages"]["mismatch"]):
        tree.update_leaf_batch(*k["update_leaf_batch"]["mismatch"])
    with pytest.raises(ValueError, match=k["messages"]["oob"]):
        tree.update_leaf_batch(*k["update_leaf_batch"]["oob"])
    assert np.all(tree.weights == k["update_leaf_batch"]["weights"])  # failed call wrote nothing
    tree.update_leaf_batch([], [])
    with pytest.raises(ValueError, match=k["messages"]["size"]):
        tree.update_all_leaves([1] * 3)
    tree.print_weights()


@pytest.mark.parametrize("size", SIZES)
def test_tree_vs_reference_fixture(WT, size):
    g = np.load(os.path.join(GOLDEN, "ref_tree.npz"))
    p = f"n{size}_"
    tree = WT(size)
    assert tree.depth == OracleTree(size).depth and np.array_equal(tree.offsets, OracleTree(size).offsets)
    tree.update_all_leaves(g[p + "leaves"])
    assert np.array_equal(tree.weights, g[p + "built"])
    assert tree.max_leaf == g[p + "built_max"]
    assert np.array_equal(np.asarray(tree.find_leaf_index(g[p + "query"])).reshape(-1), g[p + "found"])
    # device-tensor path must agree with the host path
    q = t.from_numpy(g[p + "query"]).cuda()
    assert np.array_equal(tree.find_leaf_index(q).cpu().numpy(), g[p + "found"])
    tree.update_leaf_batch(g[p + "upd_w"], g[p + "upd_idx"])
    assert np.array_equal(tree.weights, g[p + "after_upd"])
    assert tree.max_leaf == g[p + "after_upd_max"]
    tree.update_leaf_batch(t.from_numpy(g[p + "upd2_w"]).cuda(), t.from_numpy(g[p + "upd2_idx"]).cuda())
    tree.check()
    assert np.array_equal(tree.weights, g[p + "after_upd2"])
    if size > 1:
        for k in range(3):
            i, v = g[p + f"single{k}"]
            tree.update_leaf(float(v), int(i))
    assert np.array_equal(tree.weights, g[p + "after_single"])
    assert np.array_equal(np.asarray(tree.find_leaf_index(g[p + "query"])).reshape(-1), g[p + "found2"])


@pytest.mark.parametrize("size,nb", [(1000, 5000), (100000, 20000), (1 << 20, 300000), (3000001, 1000000)])
def test_tree_large_batches_vs_oracle(WT, size, nb):
    """Exercises the 11-level build passes and the cooperative (multi-CTA) update kernel."""
    rng = np.random.RandomState(size % 1000)
    leaves = rng.exponential(size=size)
    tree, ora = WT(size), OracleTree(size)
    tree.update_all_leaves(t.from_numpy(leaves).cuda())
    ora.update_all_leaves(leaves)
    assert np.array_equal(tree.weights, ora.weights)
    assert tree.max_leaf == ora.max_leaf
    idx = rng.randint(0, size, size=nb).astype(np.int64)   # many duplicates: last write wins
    w = rng.exponential(size=nb).astype(np.float32)
    tree.update_leaf_batch(t.from_numpy(w).cuda(), t.from_numpy(idx).cuda())
    ora.update_leaf_batch(w, idx)
    tree.check()
    assert np.array_equal(tree.weights, ora.weights)
    assert tree.max_leaf == ora.max_leaf
    q = rng.uniform(0, ora.get_weight_sum(), size=200000)
    assert np.array_equal(tree.find_leaf_index(q), ora.find_leaf_index(q))
    bad = idx.copy()
    bad[nb // 2] = size
    tree.update_leaf_batch(t.from_numpy(w).cuda(), t.from_numpy(bad).cuda())
    with pytest.raises(ValueError, match="out of boundary"):
        tree.check()
    assert np.array_equal(tree.weights, ora.weights)   # rejected batch wrote nothing
    tree.update_leaf_batch(t.from_numpy(w[:7]).cuda(), t.from_numpy(idx[:7]).cuda())  # next launch is not poisoned
    ora.update_leaf_batch(w[:7], idx[:7])
    assert np.array_equal(tree.weights, ora.weights)


def test_full_size_tree_properties(WT):
    """BASELINE config C2: 10 M leaves.  Size-independent properties + oracle spot checks."""
    size = 10_000_000
    rng = np.random.RandomState(0)
    leaves = (rng.exponential(size=size) + 1e-2) ** 0.6
    tree = WT(size)
    assert tree.depth == 25 and tree.n_nodes == 33554431
    tree.update_all_leaves(t.from_numpy(leaves).cuda())
    ora = OracleTree(size)
    ora.update_all_leaves(leaves)
    assert tree.get_weight_sum() == ora.get_weight_sum()
    assert np.array_equal(tree.weights, ora.weights)
    # stratified queries are sorted -> the found indexes must be sorted; prefix-sum bracket holds
    q = np.sort(rng.uniform(0, ora.get_weight_sum(), size=1 << 20))
    got = tree.find_leaf_index(q)
    assert np.all(np.diff(got) >= 0)
    assert np.array_equal(got, ora.find_leaf_index(q))
    # update then inverse update restores every node bit-for-bit (parents are recomputed, not delta'd)
    idx = rng.choice(size, size=4096, replace=False).astype(np.int64)
    old = leaves[idx]
    tree.update_leaf_batch(rng.exponential(size=4096), idx)
    tree.update_leaf_batch(old, idx)
    assert np.array_equal(tree.weights, ora.weights)


def test_kat_per_buffer_sample():
    from machin_b200.frame.buffers import PrioritizedBuffer
    k = KAT["buffer"]
    ep = lambda: [dict(state={"state_1": t.zeros([1, 2])}, action={"action_1": t.zeros([1, 3])},
                       next_state={"next_state_1": t.zeros([1, 2])}, reward=1, terminal=True, index=i)
                  for i in range(5)]
    buf = PrioritizedBuffer(**k["kwargs"])
    with pytest.raises(ValueError, match="Episode must be non-empty"):
        buf.store_episode([], priorities=[])
    buf.store_episode(ep()[:1], priorities=[1])
    assert np.all(buf.wt_tree.get_leaf_all_weights() == ((1 + 1e-2) ** 0.6, 0, 0, 0, 0))
    buf.update_priority(np.array([2]), np.array([0]))
    assert np.all(buf.wt_tree.get_leaf_all_weights() == ((2 + 1e-2) ** 0.6, 0, 0, 0, 0))
    buf = PrioritizedBuffer(**k["kwargs"])
    buf.store_episode(ep()[:1])
    assert np.all(buf.wt_tree.get_leaf_all_weights() == (1e-2 ** 0.6, 0, 0, 0, 0))
    buf.clear()
    assert buf.size() == 0 and buf.sample_batch(5) == (0, None, None, None)
    # golden sample (test_prioritized_buffer.py:371-402,462-482)
    s = k["sample"]
    np.random.seed(s["np_seed"])
    buf = PrioritizedBuffer(**k["kwargs"])
    buf.store_episode(ep(), priorities=s["priorities"])
    assert buf.sample_batch(0, sample_attrs=["index"]) == (0, None, None, None)
    bsize, result, index, is_weight = buf.sample_batch(s["batch_size"], concatenate=True, sample_attrs=["index"],
                                                       additional_concat_custom_attrs=["index"])
    assert bsize == 5
    assert result[0].flatten().cpu().tolist() == s["index"]
    assert np.all(index == s["index"])
    assert np.all(np.abs(is_weight - s["is_weight"]) < s["atol"])


@pytest.mark.parametrize("tag", ["small", "wrap", "big"])
def test_per_vs_reference_fixture(tag):
    from machin_b200.frame.buffers import PrioritizedBuffer
    g = np.load(os.path.join(GOLDEN, "ref_per.npz"))
    cap = int(g[tag + "_cap"])
    buf = PrioritizedBuffer(buffer_size=cap)
    rng = np.random.RandomState(1)
    off = 0
    for n in g[tag + "_ep_len"]:
        ep = [dict(state={"state": t.from_numpy(rng.randn(1, 4).astype(np.float32))},
                   action={"action": t.tensor([[1]])}, next_state={"state": t.zeros(1, 4)},
                   reward=0.5, terminal=False) for _ in range(n)]
        buf.store_episode(ep, priorities=list(g[tag + "_ep_pri"][off:off + n]))
        off += n
    assert np.array_equal(buf.wt_tree.weights, g[tag + "_tree"])
    assert buf.wt_tree.max_leaf == g[tag + "_max_leaf"]
    np.random.seed(7)
    for step in range(4):
        p = f"{tag}_s{step}_"
        b = len(g[p + "uniform"])
        assert len(buf.storage) == int(g[p + "nstored"]) and buf.curr_beta == float(g[p + "beta"])
        index, isw = buf.sample_index_and_weight(b)
        assert np.array_equal(index, g[p + "index"])                      # bit-exact draw
        np.testing.assert_allclose(isw, g[p + "isw"], rtol=1e-12)          # pow(): <= 2 ulp vs glibc
        buf.update_priority(g[p + "newp"], index)
        assert np.array_equal(buf.wt_tree.weights, g[p + "tree_after"])
    buf.store_episode([dict(state={"state": t.zeros(1, 4)}, action={"action": t.tensor([[1]])},
                            next_state={"state": t.zeros(1, 4)}, reward=0.5, terminal=False) for _ in range(3)])
    assert np.array_equal(buf.wt_tree.weights, g[tag + "_tree_after_default_store"])


def test_device_priority_normalisation_matches_numpy():
    """(|p| + eps) ** alpha on the device vs NumPy, float32 and float64 inputs; reports mismatches."""
    from machin_b200.frame.buffers import PrioritizedBuffer
    n = 1 << 16
    rng = np.random.RandomState(3)
    buf = PrioritizedBuffer(buffer_size=n)
    idx = np.arange(n, dtype=np.int64)
    for dtype, tol in ((np.float32, 0), (np.float64, 4)):
        p = (rng.exponential(size=n) * rng.choice([-1, 1], size=n)).astype(dtype)
        buf.update_priority(t.from_numpy(p).cuda(), t.from_numpy(idx).cuda())
        buf.wt_tree.check()
        got = buf.wt_tree.get_leaf_all_weights()
        want = ((np.abs(p) + 1e-2) ** 0.6).astype(np.float64)
        ulp = np.abs(got.view(np.int64) - want.view(np.int64)) if dtype == np.float64 else \
            np.abs(got.astype(np.float32).view(np.int32).astype(np.int64) - want.astype(np.float32).view(np.int32))
        print(f"normalise {dtype.__name__}: leaves differing from this host's NumPy {np.count_nonzero(ulp)}/{n}, "
              f"max ulp {ulp.max()}")
        # NumPy's float32/float64 power is a SIMD (AVX-512 SVML) routine on this class of host and is
        # itself not correctly rounded (~21 % of float32 results are 1 ulp off the correctly rounded
        # value), so the device can only be required to stay within 1 float32 ulp / 4 float64 ulp of it
        assert ulp.max() <= max(tol, 1)
        if dtype == np.float32:
            # ... and to equal the correctly rounded result: float64 pow of the float32 operands, rounded once
            exact = ((np.abs(p) + np.float32(1e-2)).astype(np.float64) ** np.float64(np.float32(0.6))).astype(np.float32)
            assert np.count_nonzero(got.astype(np.float32) != exact) <= 2


@pytest.mark.parametrize("size", [7, 1000, 1 << 20, 10_000_000])
def test_sorted_small_batch_update_vs_oracle(WT, size):
    """The sorted fast path (csrc/sumtree.cu: tree_update_sorted) that the learner's priority write-back
    takes: ascending indexes with duplicates, all batch sizes up to 512, float32 and float64 weights."""
    rng = np.random.RandomState(size % 97)
    leaves = rng.exponential(size=size)
    tree, ora = WT(size), OracleTree(size)
    tree.update_all_leaves(leaves)
    ora.update_all_leaves(leaves)
    for n in (1, 2, 31, 32, 33, 200, 511, 512):
        for dtype in (np.float32, np.float64):
            idx = np.sort(rng.randint(0, size, size=n)).astype(np.int64)          # duplicates: last write wins
            if n > 3:
                idx[n // 2] = idx[n // 2 - 1]                                       # force at least one run
                idx = np.sort(idx)
            w = rng.exponential(size=n).astype(dtype)
            tree.update_leaf_batch(t.from_numpy(w).cuda(), t.from_numpy(idx).cuda())
            ora.update_leaf_batch(w, idx)
            assert np.array_equal(tree.weights, ora.weights), (n, dtype)
            assert tree.max_leaf == ora.max_leaf
    # clustered keys (neighbouring leaves share almost all ancestors)
    start = int(rng.randint(0, max(1, size - 600)))
    idx = np.arange(start, min(size, start + 512), dtype=np.int64)
    w = rng.exponential(size=len(idx))
    tree.update_leaf_batch(t.from_numpy(w).cuda(), t.from_numpy(idx).cuda())
    ora.update_leaf_batch(w, idx)
    assert np.array_equal(tree.weights, ora.weights)
    # out-of-range entry in a sorted batch: rejected, nothing written
    bad = np.array([0, 1, size], dtype=np.int64)
    tree.update_leaf_batch(t.ones(3, dtype=t.float64).cuda(), t.from_numpy(bad).cuda())
    with pytest.raises(ValueError, match="out of boundary"):
        tree.check()
    assert np.array_equal(tree.weights, ora.weights)


@pytest.mark.parametrize("n", [700, 2048, 4096, 20000])
def test_masked_update_multi_launch_and_graph_replay(WT, n):
    """Priority write-back with a validity mask (version filter of the sharded buffer) at the sizes the
    fixed-shape exchange produces (P * batch entries): single-CTA path, multi-launch sparse/dense path, and
    the same call replayed from a CUDA graph."""
    size = 1 << 20
    rng = np.random.RandomState(n)
    leaves = rng.exponential(size=size)
    tree, ora = WT(size), OracleTree(size)
    tree.update_all_leaves(leaves)
    ora.update_all_leaves(leaves)
    idx = rng.randint(0, size, size=n).astype(np.int64)
    idx[n // 3] = idx[n // 3 + 1]                      # duplicates, one of them possibly masked out
    valid = (rng.rand(n) < 0.7)
    w = rng.exponential(size=n).astype(np.float32)
    d_idx, d_w, d_v = t.from_numpy(idx).cuda(), t.from_numpy(w).cuda(), t.from_numpy(valid).cuda()
    tree.update_leaf_batch(d_w, d_idx, valid=d_v)
    ora.update_leaf_batch(w[valid], idx[valid])
    tree.check()
    assert np.array_equal(tree.weights, ora.weights)
    # same launch sequence captured once and replayed with new contents in the static buffers
    side = t.cuda.Stream()
    side.wait_stream(t.cuda.current_stream())
    with t.cuda.stream(side):
        tree.update_leaf_batch(d_w, d_idx, valid=d_v)
    t.cuda.current_stream().wait_stream(side)
    t.cuda.synchronize()
    g = t.cuda.CUDAGraph()
    with t.cuda.graph(g):
        tree.update_leaf_batch(d_w, d_idx, valid=d_v)
    for rep in range(3):
        idx2 = rng.randint(0, size, size=n).astype(np.int64)
        valid2 = rng.rand(n) < 0.5
        w2 = rng.exponential(size=n).astype(np.float32)
        d_idx.copy_(t.from_numpy(idx2)); d_w.copy_(t.from_numpy(w2)); d_v.copy_(t.from_numpy(valid2))
        g.replay()
        ora.update_leaf_batch(w2[valid2], idx2[valid2])
        assert np.array_equal(tree.weights, ora.weights), rep
