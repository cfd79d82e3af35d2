"""CPU: the DEVICE SOURCE of the trajectory reverse scans (machin_b200/csrc/scan_device.cuh: vtrace_kernel,
gae_kernel) compiled with g++ against tests/cpu_emul/cuda_shim.h and compared with the reference restatement
(oracle/vtrace.py, oracle/gae.py -- impala.py:322-373, a2c.py:275-313).  Both kernels run on a B200 every round; here
they keep the CPU suite honest about the shim and give the long-episode work of the next round something to run
against.  V-trace: exp() aside every float32 operation is rounded like CPU torch -> 1e-6; returns / GAE: float64,
bit-identical.  Not a CPU path of the library."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch as t

from conftest import ROOT
from oracle import gae as ogae
from oracle import vtrace as ovt

EMUL = os.path.join(ROOT, "tests", "cpu_emul")


@pytest.fixture(scope="module")
def scan(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "scan_emul.so")
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread", "-shared", "-fPIC",
           "-I" + os.path.join(ROOT, "machin_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(EMUL, "scan_harness.cpp"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    return C.CDLL(out)


def p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


@pytest.mark.parametrize("know_len", [False, True])
@pytest.mark.parametrize("lengths,seed,use_offsets", [
    ([20] * 64, 0, True), ([20] * 300, 1, False), ([1, 7, 2, 33, 5, 1, 12, 40, 3], 2, True), ([1], 3, True),
    ([3500, 3, 2999], 4, True),                      # slices beyond the staging tile: the straight-from-memory branch
    ([100] * 200, 6, True), ([100] * 200, 7, False),  # long episodes: fewer episodes per CTA keep the staged path
    ([2976, 1, 2976, 50], 8, True), ([7000], 9, True), ([7000], 10, False), ([2977, 3, 9000, 1], 11, True),
    (list(np.random.RandomState(5).randint(1, 41, 400)), 5, True)])
def test_vtrace_kernel_source_under_emulation(scan, lengths, seed, use_offsets, know_len):
    g = t.Generator().manual_seed(seed)
    n = int(sum(lengths))
    value, next_value, reward = (t.randn(n, 1, generator=g) for _ in range(3))
    logp, logmu = -t.rand(n, 1, generator=g) * 2, -t.rand(n, 1, generator=g) * 2
    term = t.zeros(n, 1, dtype=t.bool)
    ends = np.cumsum(lengths) - 1
    term[ends[::2]] = True
    discount, clip_c, clip_rho = 0.99, 0.9, 1.2
    c, rho, delta = ovt.vtrace_prepare(logp, logmu, value, next_value, reward, term, discount, clip_c, clip_rho)
    vs, vss = ovt.vtrace_batched(value, delta, c, lengths, discount)
    f = lambda x: np.ascontiguousarray(x.numpy().reshape(-1), dtype=np.float32)
    off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64) if use_offsets else None
    out = {k: np.full(n, np.nan, np.float32) for k in ("vs", "vss", "rho", "adv")}
    term_u8 = np.ascontiguousarray(term.numpy().reshape(-1).astype(np.uint8))
    assert scan.emul_vtrace(p(off), C.c_int64(len(lengths)), C.c_int64(n), p(f(value)), p(f(next_value)), p(f(reward)),
                            p(term_u8), p(f(logp)), p(f(logmu)), C.c_float(discount), C.c_float(clip_c), C.c_float(clip_rho),
                            p(out["vs"]), p(out["vss"]), p(out["rho"]), p(out["adv"]),
                            C.c_int64(max(lengths) if know_len else 0)) == 0
    np.testing.assert_allclose(out["rho"], f(rho), rtol=1e-6)
    np.testing.assert_allclose(out["vs"], f(vs), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["vss"], f(vss), rtol=1e-5, atol=1e-6)
    adv = reward + discount * vss - value
    np.testing.assert_allclose(out["adv"], f(adv), rtol=1e-5, atol=1e-6)
    assert np.all(out["vs"][ends] == 0) and np.all(out["vss"][ends] == 0)


@pytest.mark.parametrize("lam", [1.0, 0.0, 0.95])
@pytest.mark.parametrize("know_len", [False, True])
@pytest.mark.parametrize("lengths", [[20] * 70, [1, 9, 2, 31, 4, 1, 17], [1700, 5, 40], [90] * 150, [1664, 2, 1664], [5000], [1665, 3, 4000, 1]])
def test_gae_kernel_source_under_emulation(scan, lam, lengths, know_len):
    rng = np.random.RandomState(int(lam * 100) + len(lengths))
    n = int(sum(lengths))
    reward = rng.randn(n)
    term = np.zeros(n, dtype=np.uint8)
    ends = np.cumsum(lengths) - 1
    term[ends[::2]] = 1
    cv, ncv = rng.randn(n).astype(np.float32), rng.randn(n).astype(np.float32)
    off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
    ret, adv = np.full(n, np.nan), np.full(n, np.nan)
    assert scan.emul_gae(p(off), C.c_int64(len(lengths)), C.c_int64(n), p(reward), p(term), p(cv), p(ncv),
                         C.c_double(0.97), C.c_double(lam), p(ret), p(adv), C.c_int64(max(lengths) if know_len else 0)) == 0
    pos = 0
    for L in lengths:
        sl = slice(pos, pos + L)
        want_ret = ogae.discounted_returns([float(x) for x in reward[sl]], 0.97)
        assert np.array_equal(ret[sl], np.array(want_ret))                       # float64, bit-identical
        want = ogae.gae([float(x) for x in reward[sl]], [bool(x) for x in term[sl]], [float(x) for x in cv[sl]],
                        [float(x) for x in ncv[sl]], want_ret, 0.97, lam)
        assert np.array_equal(adv[sl], np.array(want)), (lam, L)
        pos += L
