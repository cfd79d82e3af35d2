"""GPU: the *_host entry points of the C ABI (host buffers in, host buffers out -- what INTEGRATION.md's
reference-side stub binds), called with raw ctypes and checked against the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch as t

from oracle import gae as ogae
from oracle import per as oper
from oracle import vtrace as ovt
from oracle.sumtree import OracleTree

pytestmark = pytest.mark.gpu


def _p(a):
    return C.c_void_p(a.ctypes.data)


def test_tree_and_per_host_entry_points():
    from machin_b200._lib import lib
    size = 100_003
    rng = np.random.RandomState(0)
    leaves = rng.exponential(size=size)
    h = C.c_void_p()
    assert lib.mb_tree_create(size, 0, C.byref(h)) == 0
    ora = OracleTree(size)
    assert lib.mb_tree_build_host(h, _p(leaves), size) == 0
    ora.update_all_leaves(leaves)
    assert lib.mb_tree_build_host(h, _p(leaves), size - 1) == -4            # MB_ERR_SIZE_MISMATCH
    out = np.empty(ora.weights.shape[0])
    assert lib.mb_tree_read_host(h, 0, out.shape[0], _p(out)) == 0
    assert np.array_equal(out, ora.weights)
    q = rng.uniform(0, ora.get_weight_sum(), size=5000)
    idx = np.empty(5000, dtype=np.int64)
    assert lib.mb_tree_find_host(h, _p(q), 5000, _p(idx)) == 0
    assert np.array_equal(idx, ora.find_leaf_index(q))
    lw = np.empty(5000)
    assert lib.mb_tree_leaf_weight_host(h, _p(idx), 5000, _p(lw)) == 0
    assert np.array_equal(lw, ora.get_leaf_weight(idx))
    bad = idx.copy()
    bad[7] = size
    assert lib.mb_tree_leaf_weight_host(h, _p(bad), 5000, _p(lw)) == -2     # MB_ERR_INDEX_OOB
    w32 = rng.exponential(size=5000).astype(np.float32)
    assert lib.mb_tree_update_batch_host(h, _p(idx), _p(w32), 0, 5000) == 0
    ora.update_leaf_batch(w32, idx)
    assert lib.mb_tree_update_batch_host(h, _p(bad), _p(w32), 0, 5000) == -2
    assert lib.mb_tree_read_host(h, 0, out.shape[0], _p(out)) == 0
    assert np.array_equal(out, ora.weights)
    s, m, d = C.c_double(), C.c_double(), C.c_int()
    assert lib.mb_tree_stats(h, C.byref(s), C.byref(m), C.byref(d), None) == 0
    assert s.value == ora.get_weight_sum() and m.value == float(ora.max_leaf)
    assert d.value == (0 if isinstance(ora.max_leaf, np.float32) else 1)   # MB_F32 / MB_F64, like the reference's max_leaf
    for B, beta in ((512, 0.4), (4096, 0.7), (5, 1.0)):
        u = rng.uniform(size=B)
        oi, ow = np.empty(B, dtype=np.int64), np.empty(B)
        assert lib.mb_per_sample_host(h, _p(u), B, size, 0.0, beta, _p(oi), _p(ow)) == 0
        ei, ew = oper.sample_index_and_weight(ora, B, size, beta, uniform=u)
        assert np.array_equal(oi, ei)
        np.testing.assert_allclose(ow, ew, rtol=1e-12)
    assert lib.mb_tree_destroy(h) == 0


def test_vtrace_and_gae_host_entry_points():
    from machin_b200._lib import lib
    rng = np.random.RandomState(1)
    lengths = list(rng.randint(1, 41, size=300))
    S = int(sum(lengths))
    off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
    v, nv, r = (rng.randn(S).astype(np.float32) for _ in range(3))
    lp, lm = (-rng.rand(S)).astype(np.float32), (-rng.rand(S)).astype(np.float32)
    term = np.zeros(S, dtype=np.uint8)
    term[off[1:] - 1] = rng.rand(len(lengths)) < 0.5
    vs, vss, rho, adv = (np.empty(S, dtype=np.float32) for _ in range(4))
    assert lib.mb_vtrace_host(0, _p(off), len(lengths), S, _p(v), _p(nv), _p(r), _p(term), _p(lp), _p(lm), 0.99, 0.9, 1.1,
                              _p(vs), _p(vss), _p(rho), _p(adv)) == 0
    tt = lambda a: t.from_numpy(a.copy()).view(-1, 1)
    c, rh, delta = ovt.vtrace_prepare(tt(lp), tt(lm), tt(v), tt(nv), tt(r), tt(term).bool(), 0.99, 0.9, 1.1)
    evs, evss = ovt.vtrace_batched(tt(v), delta, c, lengths, 0.99)
    np.testing.assert_allclose(vs, evs.numpy().reshape(-1), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(vss, evss.numpy().reshape(-1), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rho, rh.numpy().reshape(-1), rtol=1e-6)
    r64 = r.astype(np.float64)
    for lam in (1.0, 0.0, 0.9):
        ret, gae = np.empty(S), np.empty(S)
        assert lib.mb_gae_host(0, _p(off), len(lengths), S, _p(r64), _p(term), _p(v), _p(nv), 0.97, lam, _p(ret), _p(gae)) == 0
        pos = 0
        for n in lengths:
            sl = slice(pos, pos + n)
            eret = ogae.discounted_returns([float(x) for x in r64[sl]], 0.97)
            egae = ogae.gae([float(x) for x in r64[sl]], [bool(x) for x in term[sl]], [float(x) for x in v[sl]],
                            [float(x) for x in nv[sl]], eret, 0.97, lam)
            assert np.array_equal(ret[sl], np.array(eret)) and np.array_equal(gae[sl], np.array(egae))
            pos += n


def test_host_entry_points_from_concurrent_threads():
    """ctypes releases the GIL inside the library: two Python threads (an actor storing, a learner sampling) may be in
    *_host entry points of the same GPU at once.  The per-device staging area is locked for the whole call, so each
    thread's results equal the oracle's for its own tree."""
    import threading
    from machin_b200._lib import lib
    errors = []

    def worker(seed, size, rounds):
        try:
            rng = np.random.RandomState(seed)
            h = C.c_void_p()
            assert lib.mb_tree_create(size, 0, C.byref(h)) == 0
            ora = OracleTree(size)
            leaves = rng.exponential(size=size)
            assert lib.mb_tree_build_host(h, _p(leaves), size) == 0
            ora.update_all_leaves(leaves)
            for r in range(rounds):
                n = int(rng.randint(1, 3000))
                idx = rng.randint(0, size, size=n).astype(np.int64)
                w = rng.exponential(size=n)
                assert lib.mb_tree_update_batch_host(h, _p(idx), _p(w), 1, n) == 0
                ora.update_leaf_batch(w, idx)
                q = rng.uniform(0, ora.get_weight_sum(), size=n)
                got = np.empty(n, dtype=np.int64)
                assert lib.mb_tree_find_host(h, _p(q), n, _p(got)) == 0
                assert np.array_equal(got, ora.find_leaf_index(q)), (seed, r)
                lw = np.empty(n)
                assert lib.mb_tree_leaf_weight_host(h, _p(got), n, _p(lw)) == 0
                assert np.array_equal(lw, ora.get_leaf_weight(got)), (seed, r)
            assert lib.mb_tree_destroy(h) == 0
        except BaseException as e:   # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(s, 20_000 + 977 * s, 40)) for s in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
