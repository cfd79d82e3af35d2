"""CPU: the WHOLE data path of one fixed-shape sharded update (frame/algorithms/apex.py body(): apportion -> draw for
every requester -> gather into wire rows + metadata -> all-to-all -> compaction -> [TD] -> priority scatter ->
all-to-all -> version mask -> masked tree update) for 4 shards, every step executed by the kernels' own device
source under emulation (tests/cpu_emul), against what the reference's distributed buffer does
(prioritized_buffer_d.py:184-303 through the oracle): every requester's merged batch is the per-shard draws in member
order with per-shard IS weights over the global sum, and after the write-back every shard's tree equals the oracle
tree updated requester by requester with stale versions dropped."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import per as oper
from oracle.sumtree import OracleTree
from test_exchange_emulated_cpu import META, MbField, MbPerParams, p
from test_sample_kernels_emulated_cpu import run as run_sample
from test_tree_kernels_emulated_cpu import run_kernel as run_tree

EMUL = os.path.join(ROOT, "tests", "cpu_emul")


def _build(tmp, name, shared):
    out = str(tmp / (name + (".so" if shared else "")))
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread"] + \
          (["-shared", "-fPIC"] if shared else []) + \
          ["-I" + os.path.join(ROOT, "machin_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(EMUL, name + ".cpp"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    return out


@pytest.fixture(scope="module")
def tools(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("emul")
    return {"wire": C.CDLL(_build(tmp, "wire_harness", True)), "sample": _build(tmp, "sample_harness", False),
            "tree": _build(tmp, "tree_harness", False)}


def test_fixed_shape_update_data_path_for_four_shards(tools, tmp_path):
    wire = tools["wire"]
    rng = np.random.RandomState(7)
    P, B, pitch, size = 4, 96, 48, 700                     # batch 96 -> cluster draw kernel (n > 64)
    cap, capB, W = B, B + P, pitch + META
    trees, rings, tables, stored = [], [], [], []
    for m in range(P):
        tr = OracleTree(size)
        n_m = 500 + 40 * m                                  # rows stored on shard m
        leaves = np.zeros(size)
        leaves[:n_m] = rng.exponential(size=n_m) * (m + 1)
        tr.update_all_leaves(leaves)
        trees.append(tr)
        stored.append(n_m)
        rings.append(rng.randint(0, 256, size=(size, pitch)).astype(np.uint8))
        tables.append(rng.randint(0, 4, size=size).astype(np.int64))
    sums = np.array([tr.get_weight_sum() for tr in trees])
    beta0, inc = [0.4 + 0.01 * m for m in range(P)], 1e-3
    # ---- apportioning on every shard (device kernel), identical counts everywhere
    counts_all, params_all = [], []
    for m in range(P):
        counts, scale = np.zeros(P, np.int64), np.zeros(1)
        params = (MbPerParams * P)()
        beta = np.array([beta0[m]])
        ns = np.array([stored[m]], dtype=np.int64)
        assert wire.emul_wire_apportion(p(sums), P, C.c_int64(B), m, C.c_int64(capB), p(ns), p(beta), C.c_double(inc),
                                        p(counts), p(scale), params) == 0
        counts_all.append(counts)
        params_all.append(params)
    ssize, total = oper.shard_sample_sizes(list(sums), B)
    assert all(c.tolist() == ssize for c in counts_all)
    counts, n = counts_all[0], int(sum(ssize))
    # ---- every shard draws for every requester in ONE launch (per_sample_cluster_multi), then gathers + packs
    uni = [rng.uniform(size=(P, cap)) for _ in range(P)]
    idx, isw, send = [], [], []
    for m in range(P):
        betas = [params_all[m][r].beta for r in range(P)]
        i_m, w_m = run_sample(tools["sample"], str(tmp_path), 3, trees[m], uni[m], cap, R=P, n_stored=stored[m], betas=betas,
                              all_sum=total, n_draw=ssize[m])
        for r in range(P):                                  # the draw itself against the reference restatement
            want_i, want_w = oper.sample_index_and_weight(trees[m], ssize[m], stored[m], betas[r], all_weight_sum=total,
                                                          uniform=uni[m][r, :ssize[m]])
            assert np.array_equal(i_m[r, :ssize[m]], want_i)
            np.testing.assert_allclose(w_m[r, :ssize[m]], want_w, rtol=1e-12)
        i_m, w_m = np.array(i_m), np.array(w_m)              # (frombuffer views are read-only)
        i_m[:, ssize[m]:] = 0                               # (stale slots beyond the apportioned count: any valid index)
        block = np.zeros((P * cap, W), dtype=np.uint8)
        f = (MbField * 1)(MbField(0, pitch, block.ctypes.data, W))
        flat = np.ascontiguousarray(i_m.reshape(-1))
        assert wire.emul_rows_gather(p(rings[m]), C.c_int64(pitch), C.c_int64(size), p(flat), C.c_int64(P * cap), f, 1) == 0
        assert wire.emul_pack_meta(p(block), C.c_int64(W), C.c_int64(pitch), C.c_int64(P * cap), p(flat), p(tables[m]),
                                   p(np.ascontiguousarray(w_m.reshape(-1)))) == 0
        idx.append(i_m), isw.append(w_m), send.append(block)
    version_at_sample = [tb.copy() for tb in tables]
    # ---- all-to-all + compaction on every requester
    scale = np.array([capB / n])
    merged_idx, merged_ver = [], []
    for r in range(P):
        recv = np.concatenate([send[m][r * cap:(r + 1) * cap] for m in range(P)])
        mg = np.zeros((capB, W), dtype=np.uint8)
        mi, mv, mw = np.zeros(capB, np.int64), np.zeros(capB, np.int64), np.zeros(capB)
        assert wire.emul_wire_compact(p(recv), P, C.c_int64(cap), C.c_int64(W), C.c_int64(pitch), p(counts), p(mg),
                                      C.c_int64(capB), p(mi), p(mv), p(mw), p(scale)) == 0
        j = 0
        for m in range(P):
            for k in range(ssize[m]):
                s = int(idx[m][r, k])
                assert np.array_equal(mg[j, :pitch], rings[m][s]) and mi[j] == s and mw[j] == isw[m][r, k] * scale[0]
                j += 1
        assert j == n and np.all(mw[n:] == 0)
        merged_idx.append(mi), merged_ver.append(mv)
    # ---- |delta| of every merged row goes back to its shard; some slots were overwritten meanwhile
    pri = [rng.rand(capB).astype(np.float32) for _ in range(P)]
    send2 = []
    for r in range(P):
        s2 = np.zeros((P * cap, META), dtype=np.uint8)
        assert wire.emul_wire_scatter_update(p(s2), P, C.c_int64(cap), p(counts), p(merged_idx[r]), p(merged_ver[r]), p(pri[r]),
                                             C.c_int64(capB)) == 0
        send2.append(s2)
    for m in range(P):
        tables[m][rng.rand(size) < 0.25] += 1
    offset = np.concatenate([[0], np.cumsum(ssize)])
    for m in range(P):
        recv2 = np.concatenate([send2[r][m * cap:(m + 1) * cap] for r in range(P)])
        u_idx, u_pri, u_val = np.zeros(P * cap, np.int64), np.zeros(P * cap, np.float32), np.zeros(P * cap, np.uint8)
        assert wire.emul_wire_update_mask(p(recv2), P, C.c_int64(cap), p(counts), m, p(tables[m]), C.c_int64(size), p(u_idx),
                                          p(u_pri), p(u_val)) == 0
        # masked tree update with on-device normalisation: the generic single-CTA body (P * cap <= 1024 entries)
        got, max_leaf, _, err = run_tree(tools["tree"], str(tmp_path), "generic", trees[m], u_idx, u_pri, valid=u_val,
                                         normalize=(1e-2, 0.6))
        assert err == 0
        # reference: _update_priority_service per requester, in member order, stale versions dropped (:282-291)
        want = OracleTree(size)
        want.weights[:] = trees[m].weights
        want.max_leaf = trees[m].max_leaf
        for r in range(P):
            ix = idx[m][r, :ssize[m]]
            pr = pri[r][offset[m]:offset[m] + ssize[m]]
            keep = tables[m][ix] == version_at_sample[m][ix]
            if keep.any():
                x = (np.abs(pr[keep]) + np.float32(1e-2)).astype(np.float32)
                leaf = np.power(x.astype(np.float64), np.float64(np.float32(0.6))).astype(np.float32)
                want.update_leaf_batch(leaf, ix[keep])
        assert np.array_equal(got, want.weights), m
