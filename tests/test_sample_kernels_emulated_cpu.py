"""CPU: the DEVICE SOURCE of the sum-tree descent and of the fused PER draw (machin_b200/csrc/sumtree_sample.cuh)
compiled with g++ against tests/cpu_emul/cuda_shim_mp.h and compared with the reference restatement
(oracle/sumtree.py find_leaf_index, oracle/per.py sample_index_and_weight -- prioritized_buffer.py:96-125,
:397-428): indexes bit for bit, IS weights to 1e-12.

tree_find_kernel, per_sample_small and per_sample_cluster run on a B200 every round: here they check the shim (the
cluster kernel exchanges its maxima over emulated distributed shared memory).  per_sample_cluster_multi (one launch
for all requesters of a fixed-shape shard exchange) has its first execution here.  Not a CPU path of the library."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import per as oper
from oracle.sumtree import OracleTree

EMUL = os.path.join(ROOT, "tests", "cpu_emul")


@pytest.fixture(scope="module")
def sample_emul(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("emul") / "sample_emul")
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread",
           "-I" + os.path.join(ROOT, "machin_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(EMUL, "sample_harness.cpp"), "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    return exe


def run(exe, tmp, variant, tree, u, n, R=1, n_stored=0, betas=(0.4,), all_sum=0.0, n_draw=0):
    src, dst = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    with open(src, "wb") as f:
        f.write(struct.pack("<7q", variant, tree.size, tree.depth, n, R, n_stored, n_draw))
        f.write(np.asarray(betas, dtype=np.float64).tobytes())
        f.write(struct.pack("<d", all_sum))
        f.write(np.ascontiguousarray(tree.weights, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(u, dtype=np.float64).tobytes())
    res = subprocess.run([exe, src, dst], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, (res.returncode, res.stderr[-2000:])
    raw = open(dst, "rb").read()
    idx = np.frombuffer(raw[:8 * R * n], dtype=np.int64).reshape(R, n)
    isw = np.frombuffer(raw[8 * R * n:], dtype=np.float64).reshape(R, n)
    return idx, isw


def _tree(size, seed):
    rng = np.random.RandomState(seed)
    leaves = rng.exponential(size=size)
    leaves[rng.rand(size) < 0.1] = 0.0
    tree = OracleTree(size)
    tree.update_all_leaves(leaves)
    return tree, rng


@pytest.mark.parametrize("size", [5, 1000, 1025, 1 << 16, 100_000])
def test_descent_under_emulation(sample_emul, tmp_path, size):
    tree, rng = _tree(size, size % 91)
    total = tree.get_weight_sum()
    q = np.concatenate((rng.uniform(-0.5, total * 1.1 + 0.5, size=400), [0.0, total, total - 1e-6, -1.0, 2 * total + 1],
                        np.cumsum(tree.get_leaf_all_weights())[:64]))
    idx, _ = run(sample_emul, str(tmp_path), 0, tree, q, len(q))
    assert np.array_equal(idx[0], np.asarray(tree.find_leaf_index(q)).reshape(-1))


@pytest.mark.parametrize("variant,n", [(1, 5), (1, 64), (1, 700), (2, 65), (2, 256), (2, 512), (2, 1000)])
def test_fused_draw_under_emulation(sample_emul, tmp_path, variant, n):
    tree, rng = _tree(50_000, n)
    u = rng.uniform(size=n)
    n_stored, beta = 43_210, 0.4 + 0.001 * n
    idx, isw = run(sample_emul, str(tmp_path), variant, tree, u, n, n_stored=n_stored, betas=(beta,))
    want_idx, want_isw = oper.sample_index_and_weight(tree, n, n_stored, beta, uniform=u)
    assert np.array_equal(idx[0], want_idx)
    np.testing.assert_allclose(isw[0], want_isw, rtol=1e-12)


@pytest.mark.parametrize("R,n,n_draw", [(1, 512, 0), (2, 512, 300), (4, 96, 17), (8, 512, 71), (3, 1000, 999)])
def test_multi_requester_draw_under_emulation(sample_emul, tmp_path, R, n, n_draw):
    """One launch serves R requesters: requester r draws n_draw (or n) samples from its own uniforms with its own
    beta, normalised by its own maximum, global weight sum given (prioritized_buffer_d.py:293-303)."""
    tree, rng = _tree(80_000, R * 1000 + n)
    u = rng.uniform(size=(R, n))
    betas = [min(1.0, 0.4 + 0.001 * r) for r in range(R)]
    all_sum = tree.get_weight_sum() * 3.7
    idx, isw = run(sample_emul, str(tmp_path), 3, tree, u, n, R=R, n_stored=60_000, betas=betas, all_sum=all_sum,
                   n_draw=n_draw)
    d = n_draw if 0 < n_draw < n else n
    for r in range(R):
        want_idx, want_isw = oper.sample_index_and_weight(tree, d, 60_000, betas[r], all_weight_sum=all_sum, uniform=u[r, :d])
        assert np.array_equal(idx[r, :d], want_idx), r
        np.testing.assert_allclose(isw[r, :d], want_isw, rtol=1e-12)
    # the single-requester kernel with a device-side parameter block takes the same route (n_draw < n)
    idx1, isw1 = run(sample_emul, str(tmp_path), 4, tree, u[0], n, n_stored=60_000, betas=betas[:1], all_sum=all_sum,
                     n_draw=n_draw)
    assert np.array_equal(idx1[0, :d], idx[0, :d]) and np.array_equal(isw1[0, :d], isw[0, :d])
