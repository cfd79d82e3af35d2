"""CPU: ONE fixed-shape shard exchange (DQNApex static exchange, frame/algorithms/apex.py: draw -> gather into wire
rows -> pack metadata -> all-to-all -> compaction -> TD -> priority scatter -> all-to-all -> version mask) replayed
for 1, 2, 4 and 8 shards with the DEVICE SOURCE of the kernels (machin_b200/csrc/rows_device.cuh) compiled with
g++ against tests/cpu_emul/cuda_shim.h; the all-to-all is a NumPy block transpose.  The exchange was measured on 2
GPUs only (round 1); this replays the 4- and 8-shard data path in the build container.  Expectations are written
directly from prioritized_buffer_d.py:204-303 (merged batch in member order, (index, version) per shard, stale
versions dropped).  Not a CPU path of the library."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

EMUL = os.path.join(ROOT, "tests", "cpu_emul")
META = 24


class MbField(C.Structure):
    _fields_ = [("row_offset", C.c_int64), ("nbytes", C.c_int64), ("d_dense", C.c_void_p), ("dense_pitch", C.c_int64)]


@pytest.fixture(scope="module")
def wire(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "wire_emul.so")
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread", "-shared", "-fPIC",
           "-I" + os.path.join(ROOT, "machin_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(EMUL, "wire_harness.cpp"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    return C.CDLL(out)


def p(a):
    return C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("P", [1, 2, 4, 8])
@pytest.mark.parametrize("pitch", [48, 688])
def test_fixed_shape_exchange_replayed_for_P_shards(wire, P, pitch):
    rng = np.random.RandomState(P * 1000 + pitch)
    B = 24                                   # requested batch; every (shard, requester) block holds `cap` = B wire rows
    cap, capB, W, rows_in_ring = B, B + P, pitch + META, 200
    # ceil apportioning gives sum(counts) in [B, B + P - 1]; include an empty shard when there are enough of them
    counts = rng.multinomial(B, np.ones(P) / P).astype(np.int64)
    if P >= 4:
        counts[1] += counts[2]
        counts[2] = 0
    counts[0] += min(P - 1, 2)
    assert counts.sum() <= capB - 1 and counts.max() <= cap
    ring = [rng.randint(0, 256, size=(rows_in_ring, pitch)).astype(np.uint8) for _ in range(P)]
    table = [rng.randint(0, 5, size=rows_in_ring).astype(np.int64) for _ in range(P)]
    # ---- every shard m serves every requester r: draw (given), gather into the wire block, pack metadata
    idx = [[np.sort(rng.randint(0, rows_in_ring, size=cap)).astype(np.int64) for _ in range(P)] for _ in range(P)]
    isw = [[rng.rand(cap) for _ in range(P)] for _ in range(P)]
    send = [np.zeros((P * cap, W), dtype=np.uint8) for _ in range(P)]
    for m in range(P):
        for r in range(P):
            block = send[m][r * cap:(r + 1) * cap]
            f = (MbField * 1)(MbField(0, pitch, block.ctypes.data, W))
            assert wire.emul_rows_gather(p(ring[m]), C.c_int64(pitch), C.c_int64(rows_in_ring), p(idx[m][r]), C.c_int64(cap), f, 1) == 0
            assert wire.emul_pack_meta(p(block), C.c_int64(W), C.c_int64(pitch), C.c_int64(cap), p(idx[m][r]), p(table[m]), p(isw[m][r])) == 0
    version_at_sample = [t.copy() for t in table]
    # ---- all-to-all: requester r receives block r of every shard, in member order
    recv = [np.concatenate([send[m][r * cap:(r + 1) * cap] for m in range(P)]) for r in range(P)]
    # ---- the peer-memory variant (csrc/peer.cu: mb_wire_push_rows): ONE kernel per shard gathers, adds the metadata
    #      and stores every row straight into its requester's receive buffer -- must equal gather + pack + all-to-all
    pushed = [np.zeros((P * cap, W), dtype=np.uint8) for _ in range(P)]
    peers = (C.c_void_p * P)(*[b.ctypes.data for b in pushed])
    for m in range(P):
        all_idx, all_isw = np.concatenate(idx[m]), np.concatenate(isw[m])
        assert wire.emul_wire_push_rows(p(ring[m]), C.c_int64(pitch), C.c_int64(rows_in_ring), p(all_idx), p(all_isw),
                                        p(table[m]), P, C.c_int64(cap), m, C.c_int64(W), peers) == 0
    for r in range(P):
        assert np.array_equal(pushed[r], recv[r]), r
    n = int(counts.sum())
    scale = np.array([capB / n])
    merged, m_idx, m_ver, m_isw = [], [], [], []
    for r in range(P):
        mg = np.full((capB, W), 0xAB, dtype=np.uint8)
        mi, mv, mw = np.full(capB, -7, np.int64), np.full(capB, -7, np.int64), np.full(capB, -7.0)
        assert wire.emul_wire_compact(p(recv[r]), P, C.c_int64(cap), C.c_int64(W), C.c_int64(pitch), p(counts), p(mg),
                                      C.c_int64(capB), p(mi), p(mv), p(mw), p(scale)) == 0
        j = 0
        for m in range(P):                                  # merged batch: shard 0's rows, then shard 1's, ...
            for k in range(int(counts[m])):
                s = int(idx[m][r][k])
                assert np.array_equal(mg[j, :pitch], ring[m][s]), (r, m, k)
                assert mi[j] == s and mv[j] == version_at_sample[m][s]
                assert mw[j] == isw[m][r][k] * scale[0]
                j += 1
        assert j == n
        assert np.all(mg[n:] == 0) and np.all(mi[n:] == 0) and np.all(mv[n:] == -1) and np.all(mw[n:] == 0.0)
        # compact + unpack in one launch (mb_wire_compact_fields): two fields cut out of every merged row land in dense
        # tensors, byte for byte what splitting `mg` gives; same metadata
        half = (pitch // 2) // 8 * 8
        fa, fb = np.full((capB, half), 0x55, np.uint8), np.full((capB, pitch - half), 0x55, np.uint8)
        fl = (MbField * 2)(MbField(0, half, fa.ctypes.data, half), MbField(half, pitch - half, fb.ctypes.data, pitch - half))
        mi2, mv2, mw2 = np.full(capB, -7, np.int64), np.full(capB, -7, np.int64), np.full(capB, -7.0)
        assert wire.emul_wire_compact_fields(p(recv[r]), P, C.c_int64(cap), C.c_int64(W), C.c_int64(pitch), p(counts), fl, 2,
                                             C.c_int64(capB), p(mi2), p(mv2), p(mw2), p(scale)) == 0
        assert np.array_equal(fa, mg[:, :half]) and np.array_equal(fb, mg[:, half:pitch])
        assert np.array_equal(mi2, mi) and np.array_equal(mv2, mv) and np.array_equal(mw2, mw)
        merged.append(mg), m_idx.append(mi), m_ver.append(mv), m_isw.append(mw)
    # ---- after TD: every requester sends (index, version, |delta|) back to the shard each row came from
    pri = [rng.rand(capB).astype(np.float32) for _ in range(P)]
    send2 = [np.full((P * cap, META), 0xCD, dtype=np.uint8) for _ in range(P)]
    for r in range(P):
        assert wire.emul_wire_scatter_update(p(send2[r]), P, C.c_int64(cap), p(counts), p(m_idx[r]), p(m_ver[r]), p(pri[r]),
                                             C.c_int64(capB)) == 0
    recv2 = [np.concatenate([send2[r][m * cap:(m + 1) * cap] for r in range(P)]) for m in range(P)]
    # peer-memory variant (mb_wire_push_update): the same (index, version, priority) triples land in the serving
    # shard's buffer directly; entries beyond counts[m] are never written (and never read: the mask checks counts)
    pushed2 = [np.full((P * cap, META), 0xCD, dtype=np.uint8) for _ in range(P)]
    peers2 = (C.c_void_p * P)(*[b.ctypes.data for b in pushed2])
    for r in range(P):
        assert wire.emul_wire_push_update(P, C.c_int64(cap), r, p(counts), p(m_idx[r]), p(m_ver[r]), p(pri[r]),
                                          C.c_int64(capB), peers2) == 0
    for m in range(P):
        assert np.array_equal(pushed2[m], recv2[m]), m
    # some slots were overwritten between sampling and write-back: their version moved on (prioritized_buffer_d.py:131)
    for m in range(P):
        table[m][rng.rand(rows_in_ring) < 0.2] += 1
    offset = np.concatenate([[0], np.cumsum(counts)])
    for m in range(P):
        u_idx = np.full(P * cap, -7, np.int64)
        u_pri = np.full(P * cap, -7, np.float32)
        u_val = np.full(P * cap, 9, np.uint8)
        assert wire.emul_wire_update_mask(p(recv2[m]), P, C.c_int64(cap), p(counts), m, p(table[m]), C.c_int64(rows_in_ring),
                                          p(u_idx), p(u_pri), p(u_val)) == 0
        for r in range(P):
            for k in range(cap):
                e = r * cap + k
                if k < counts[m]:
                    s = int(idx[m][r][k])
                    assert u_idx[e] == s and u_pri[e] == pri[r][offset[m] + k]
                    assert u_val[e] == int(table[m][s] == version_at_sample[m][s])
                else:
                    assert u_val[e] == 0 and u_idx[e] == 0


class MbPerParams(C.Structure):
    _fields_ = [("beta", C.c_double), ("all_weight_sum", C.c_double), ("n_stored", C.c_int64), ("step", C.c_int64),
                ("n_draw", C.c_int64)]


@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_device_apportioning_matches_host_apportioning(wire, P):
    """wire_apportion_kernel against the host formula the dynamic exchange uses (prioritized_buffer_d.py:229-234 via
    oracle.per.shard_sample_sizes) and the per-draw beta schedule (prioritized_buffer.py:425-427): bit-identical."""
    from oracle.per import shard_sample_sizes
    assert C.sizeof(MbPerParams) == 40
    rng = np.random.RandomState(P)
    for trial in range(20):
        B = int(rng.choice([24, 100, 512]))
        sums = rng.exponential(size=P) * rng.choice([1.0, 1e3, 1e-3], size=P)
        if P > 2 and trial % 3 == 0:
            sums[1] = 0.0                                   # an empty shard draws nothing
        me = int(rng.randint(P))
        beta0, inc = float(rng.choice([0.4, 0.9995, 1.0])), 1e-3
        n_stored = np.array([12345], dtype=np.int64)
        beta = np.array([beta0])
        counts, scale = np.zeros(P, np.int64), np.zeros(1)
        params = (MbPerParams * P)()
        capB = B + P
        assert wire.emul_wire_apportion(p(sums), P, C.c_int64(B), me, C.c_int64(capB), p(n_stored), p(beta), C.c_double(inc),
                                        p(counts), p(scale), params) == 0
        ssize, total = shard_sample_sizes(list(sums), B)
        assert counts.tolist() == ssize
        assert scale[0] == capB / sum(ssize)
        b = beta0
        for r in range(P):
            assert (params[r].beta, params[r].all_weight_sum, params[r].n_stored, params[r].step, params[r].n_draw) == \
                (b, total, 12345, 0, ssize[me])
            if ssize[me] > 0:
                b = float(np.min([1.0, b + inc]))
        assert beta[0] == b


@pytest.mark.parametrize("P", [1, 2, 3, 8])
def test_peer_push_and_ordered_sum(wire, P):
    """mb_peer_push + mb_peer_sum: every member's buffer lands in slot `self` of every member, and the sum is taken in
    member order with separately rounded float32 adds and one division -- all members get bit-identical averages."""
    rng = np.random.RandomState(P)
    n = 1000 + P                                             # not a multiple of 4: the byte-wise tail of the copy
    grads = [(rng.randn(n) * 10 ** rng.uniform(-3, 3, size=n)).astype(np.float32) for _ in range(P)]
    slots = [np.full((P, n), np.nan, dtype=np.float32) for _ in range(P)]
    peers = (C.c_void_p * P)(*[b.ctypes.data for b in slots])
    for m in range(P):
        assert wire.emul_peer_push(p(grads[m]), C.c_int64(4 * n), P, m, C.c_int64(4 * n), peers) == 0
    want = grads[0].copy()
    for m in range(1, P):
        want = (want + grads[m]).astype(np.float32)
    want = (want / np.float32(P)).astype(np.float32)
    for m in range(P):
        assert np.array_equal(slots[m], np.stack(grads))
        out = np.zeros(n, dtype=np.float32)
        assert wire.emul_peer_sum(p(slots[m]), P, C.c_int64(n), C.c_float(P), p(out)) == 0
        assert np.array_equal(out, want)
