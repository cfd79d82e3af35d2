"""CPU: the DEVICE SOURCE of the small-batch sum-tree update kernels (machin_b200/csrc/sumtree_update.cuh)
compiled with g++ against tests/cpu_emul/cuda_shim_mp.h -- CUDA threads = OS threads, warps with ballot / shuffle,
CTA barriers, a cluster of 8 CTAs as 8 processes with distributed-shared-memory reads -- and compared bit for bit
with the reference restatement (OracleTree.update_leaf_batch, prioritized_buffer.py:152-186).

Variant 0 (tree_update_sorted) and 4 (the generic body) are the kernels already validated on a B200: they check the
shim.  Variants 1-3 (tree_update_sorted_v2: barrier-free walk, cluster of 8 subtrees, both) were written after the
GPU budget of round 1 was spent; this is their first execution anywhere.  Not a CPU path of the library."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle.sumtree import OracleTree

EMUL = os.path.join(ROOT, "tests", "cpu_emul")
VARIANTS = {"sorted": 0, "v2_async": 1, "v2_cluster": 2, "v2_cluster_async": 3, "generic": 4, "part": 5}


@pytest.fixture(scope="module")
def tree_emul(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("emul") / "tree_emul")
    # (tree_update_part with 128 instead of 512 threads per CTA: the same source, a quarter of the OS threads per emulated CTA)
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread", "-DMB_PART_THREADS=128",
           "-I" + os.path.join(ROOT, "machin_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(EMUL, "tree_harness.cpp"), "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    return exe


def run_kernel(exe, tmp, variant, tree, idx, w, valid=None, normalize=None):
    """Run one emulated launch on a copy of ``tree`` (an OracleTree) -> (weights, max_leaf, max_is_f32, err)."""
    src, dst = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    w = np.ascontiguousarray(w)
    assert w.dtype in (np.float32, np.float64)
    eps, alpha = normalize or (0.0, 0.0)
    with open(src, "wb") as f:
        f.write(struct.pack("<7q", VARIANTS[variant], tree.size, tree.depth, len(idx), int(w.dtype == np.float64),
                            int(normalize is not None), int(valid is not None)))
        f.write(struct.pack("<3d", eps, alpha, float(tree.max_leaf)))
        f.write(np.ascontiguousarray(tree.weights, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(idx, dtype=np.int64).tobytes())
        f.write(w.tobytes())
        if valid is not None:
            f.write(np.ascontiguousarray(valid, dtype=np.uint8).tobytes())
    res = subprocess.run([exe, src, dst], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, (res.returncode, res.stderr[-2000:])
    raw = open(dst, "rb").read()
    nodes = len(tree.weights)
    out = np.frombuffer(raw[:8 * nodes], dtype=np.float64).copy()
    max_leaf, = struct.unpack("<d", raw[8 * nodes:8 * nodes + 8])
    is_f32, err = struct.unpack("<2q", raw[8 * nodes + 8:])
    return out, max_leaf, is_f32, err


def _tree(size, seed):
    rng = np.random.RandomState(seed)
    tree = OracleTree(size)
    tree.update_all_leaves(rng.exponential(size=size))
    return tree, rng


CASES = [  # (tree size, batch, kind)
    (1000, 40, "uniform"), (1 << 14, 96, "uniform"), (1 << 14, 64, "dense"), (100_000, 33, "clumps"),
    (1 << 14, 70, "one_subtree"), (5000, 1, "uniform"), (1 << 16, 512, "uniform"), (1 << 20, 200, "clumps"),
]


def _indexes(rng, size, n, kind):
    if kind == "uniform":
        idx = rng.randint(0, size, size=n)
    elif kind == "dense":
        start = int(rng.randint(0, size - n))
        idx = start + np.arange(n)
    elif kind == "one_subtree":
        span = max(1, size // 16)
        idx = rng.randint(size - span, size, size=n)
    else:
        centres = rng.randint(0, size, size=max(1, n // 8))
        idx = np.clip(rng.choice(centres, size=n) + rng.randint(0, 4, size=n), 0, size - 1)
    idx = np.sort(idx).astype(np.int64)
    if n > 3:
        idx[n // 2] = idx[n // 2 - 1]                     # a run of equal keys: the last position wins
        idx = np.sort(idx)
    return idx


@pytest.mark.parametrize("variant", ["sorted", "v2_async", "v2_cluster", "v2_cluster_async", "part"])
@pytest.mark.parametrize("size,n,kind", CASES)
def test_sorted_update_kernels_under_emulation(tree_emul, tmp_path, variant, size, n, kind):
    if variant == "part" and (size, n, kind) not in ((1 << 14, 96, "uniform"), (1 << 16, 512, "uniform"), (1 << 14, 70, "one_subtree"),
                                                     (1000, 40, "uniform")):
        pytest.skip("one CTA per subtree = up to 128 emulated CTAs of 512 OS threads each: four representative cases in the CPU "
                    "suite (the GPU suite runs them all)")
    if n > 256 and "cluster" in variant:
        pytest.skip("8 CTAs x 512 OS threads: kept out of the CPU suite (the 200-entry case covers several warps per CTA)")
    tree, rng = _tree(size, size % 97 + n)
    idx = _indexes(rng, size, n, kind)
    w = rng.exponential(size=n).astype(np.float32 if n % 2 else np.float64)
    got, max_leaf, is_f32, err = run_kernel(tree_emul, str(tmp_path), variant, tree, idx, w)
    tree.update_leaf_batch(w, idx)
    assert err == 0
    assert np.array_equal(got, tree.weights)
    assert max_leaf == tree.max_leaf


@pytest.mark.parametrize("variant", ["sorted", "v2_async", "v2_cluster", "v2_cluster_async", "generic", "part"])
def test_unsorted_masked_and_out_of_range_under_emulation(tree_emul, tmp_path, variant):
    size, n = 3000, 64
    tree, rng = _tree(size, 5)
    before = tree.weights.copy()
    # unsorted input: every variant falls back to the generic body inside the same launch
    idx = rng.randint(0, size, size=n).astype(np.int64)
    idx[7] = idx[3]
    w = rng.exponential(size=n)
    got, max_leaf, _, err = run_kernel(tree_emul, str(tmp_path), variant, tree, idx, w)
    ref = OracleTree(size)
    ref.weights[:] = before
    ref.max_leaf = tree.max_leaf
    ref.update_leaf_batch(w, idx)
    assert err == 0 and np.array_equal(got, ref.weights) and max_leaf == ref.max_leaf
    # validity mask (version filter): masked entries are skipped, even when out of range
    valid = rng.rand(n) < 0.6
    idx2 = np.sort(rng.randint(0, size, size=n)).astype(np.int64)
    idx2[np.flatnonzero(~valid)[0]] = size + 5
    got, _, _, err = run_kernel(tree_emul, str(tmp_path), variant, tree, idx2, w, valid=valid)
    ref = OracleTree(size)
    ref.weights[:] = before
    ref.update_leaf_batch(w[valid], idx2[valid])
    assert err == 0 and np.array_equal(got, ref.weights)
    # an out-of-range index in a sorted batch: rejected, nothing written
    bad = np.sort(rng.randint(0, size, size=n)).astype(np.int64)
    bad[-1] = size
    got, _, _, err = run_kernel(tree_emul, str(tmp_path), variant, tree, bad, w)
    assert err == -2 and np.array_equal(got, before)       # MB_ERR_INDEX_OOB


def test_priority_normalisation_float32_under_emulation(tree_emul, tmp_path):
    """(|p| + eps) ** alpha on float32 priorities: float32 add, float64 pow, rounded once (csrc: normalized_weight)."""
    size, n = 2048, 96
    tree, rng = _tree(size, 9)
    idx = np.sort(rng.choice(size, size=n, replace=False)).astype(np.int64)
    p = (rng.exponential(size=n) * rng.choice([-1, 1], size=n)).astype(np.float32)
    got, max_leaf, is_f32, err = run_kernel(tree_emul, str(tmp_path), "v2_cluster_async", tree, idx, p, normalize=(1e-2, 0.6))
    x = (np.abs(p) + np.float32(1e-2)).astype(np.float32)
    leaf = np.power(x.astype(np.float64), np.float64(np.float32(0.6))).astype(np.float32)
    tree.update_leaf_batch(leaf, idx)
    assert err == 0 and is_f32 in (0, 1)
    assert np.array_equal(got, tree.weights)


@pytest.mark.parametrize("size,n,kind", [
    (1 << 12, 1500, "merged_blocks"),      # P ascending blocks concatenated (the sharded write-back), masked, duplicates across blocks
    (1 << 15, 4096, "skewed"),             # more distinct entries in ONE subtree (256 leaves) than a CTA walks out of shared memory
    (50, 9, "tiny"),                       # depth 7: a single CTA, no cut
    (1, 1, "one_leaf"),                    # depth 1: the leaf is the root
])
def test_partitioned_update_kernel_under_emulation(tree_emul, tmp_path, size, n, kind):
    """tree_update_part: one CTA per subtree, batches up to 4096 in any order, with a mask and duplicates."""
    tree, rng = _tree(size, 11 + n)
    valid = None
    if kind == "merged_blocks":
        idx = np.concatenate([np.sort(rng.randint(0, size, size=n // 4)) for _ in range(4)])
        idx = np.concatenate([idx, idx[:n - len(idx)]]).astype(np.int64)
        idx[5] = idx[n // 2 + 3]
        valid = rng.rand(n) < 0.8
        idx[np.flatnonzero(~valid)[:3]] = size + 100           # masked-out entries may be out of range
    elif kind == "skewed":
        span = size >> part_cut_py(size)
        idx = np.concatenate([rng.randint(3 * span, 4 * span, size=n - 200), rng.randint(0, size, size=200)]).astype(np.int64)
    elif kind == "skewed_sorted":
        span = size >> part_cut_py(size)
        idx = np.sort(rng.choice(np.arange(2 * span, 3 * span), size=n, replace=False)).astype(np.int64)
    else:
        idx = rng.randint(0, size, size=n).astype(np.int64)
    w = rng.exponential(size=n).astype(np.float32 if n % 2 else np.float64)
    got, max_leaf, is_f32, err = run_kernel(tree_emul, str(tmp_path), "part", tree, idx, w, valid=valid)
    if valid is not None:
        tree.update_leaf_batch(w[valid], idx[valid])
    else:
        tree.update_leaf_batch(w, idx)
    assert err == 0
    assert np.array_equal(got, tree.weights)
    assert max_leaf == tree.max_leaf
    # a second launch on the result: the ticket was reset by the last CTA of the first
    tree.weights[:] = got
    got2, *_ , err2 = run_kernel(tree_emul, str(tmp_path), "part", tree, idx[:1], w[:1],
                                 valid=None if valid is None else np.ones(1, dtype=bool))
    tree.update_leaf_batch(w[:1], idx[:1]) if idx[0] < size else None


def part_cut_py(size):
    depth = int(np.ceil(np.log2(size))) + 1 if size > 1 else 1
    return max(0, min(7, depth - 1 - 6))
