"""CPU: the DEVICE SOURCE of the row-tile DQN kernel (machin_b200/csrc/rows_program.cuh: dqn_rows_kernel --
forward of Q(s), Q(s'), Q_target(s'), TD row math, backward-input chain) compiled with g++ against
tests/cpu_emul/cuda_shim.h (CUDA threads = OS threads, __syncthreads = a barrier, one CTA at a time) and compared
with a NumPy restatement of dqn_per.py:105-139 + a Linear/ReLU stack.  This checks the kernel's indexing and
data flow in the build container, where no GPU exists; it is not a CPU path of the library."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

EMUL = os.path.join(ROOT, "tests", "cpu_emul")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "rows_emul.so")
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread", "-shared", "-fPIC",
           "-I" + os.path.join(ROOT, "machin_b200", "csrc"), os.path.join(EMUL, "rows_harness.cpp"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    lib = C.CDLL(out)
    lib.emul_dqn_rows.restype = C.c_int
    return lib


def _ptrs(arrays):
    return (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])


def _forward(Ws, bs, x):
    hs, h = [], x
    for l, (W, b) in enumerate(zip(Ws, bs)):
        z = h @ W.T + b
        h = np.maximum(z, 0) if l < len(Ws) - 1 else z
        hs.append(h)
    return hs


@pytest.mark.parametrize("sizes,batch", [([6, 16, 16, 4], 32), ([84, 256, 256, 6], 20), ([7, 33, 10, 3], 50),
                                          ([12, 64, 5], 9), ([20, 8, 8, 8, 2], 17), ([5, 3], 8)])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("cpt", [4, 2])
def test_rows_kernel_source_under_emulation(harness, sizes, batch, mode, cpt):
    rng = np.random.RandomState(len(sizes) * 100 + batch + mode)
    L, S, A = len(sizes) - 1, sizes[0], sizes[-1]
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    W = [f32(rng.randn(sizes[l + 1], sizes[l]) / np.sqrt(sizes[l])) for l in range(L)]
    b = [f32(rng.randn(sizes[l + 1]) * 0.1) for l in range(L)]
    tW = [f32(w + 0.05 * rng.randn(*w.shape)) for w in W]
    tb = [f32(x + 0.05 * rng.randn(*x.shape)) for x in b]
    Wt, tWt = [f32(w.T) for w in W], [f32(w.T) for w in tW]
    state, nstate = f32(rng.randn(batch, S)), f32(rng.randn(batch, S))
    action = rng.randint(0, A, size=batch).astype(np.int64)
    reward = f32(rng.randn(batch))
    terminal = (rng.rand(batch) < 0.3).astype(np.uint8)
    isw = rng.rand(batch).astype(np.float64) + 0.1
    gamma = np.float32(0.97)
    h = [np.full((batch, sizes[l + 1]), np.nan, dtype=np.float32) for l in range(L)]
    dz = [np.full((batch, sizes[l + 1]), np.nan, dtype=np.float32) for l in range(L)]
    qn, qtn = np.full((batch, A), np.nan, np.float32), np.full((batch, A), np.nan, np.float32)
    y, av, ab = (np.full(batch, np.nan, np.float32) for _ in range(3))
    loss = np.full(batch, np.nan, np.float64)
    rc = harness.emul_dqn_rows(
        L, (C.c_int * (L + 1))(*sizes), batch, _ptrs(W), _ptrs(Wt), _ptrs(tWt), _ptrs(b), _ptrs(tb),
        C.c_void_p(state.ctypes.data), C.c_void_p(nstate.ctypes.data), _ptrs(h), C.c_void_p(qn.ctypes.data),
        C.c_void_p(qtn.ctypes.data), _ptrs(dz), C.c_void_p(action.ctypes.data), C.c_void_p(reward.ctypes.data),
        C.c_void_p(terminal.ctypes.data), C.c_void_p(isw.ctypes.data), C.c_float(float(gamma)), mode,
        C.c_void_p(y.ctypes.data), C.c_void_p(av.ctypes.data), C.c_void_p(ab.ctypes.data), C.c_void_p(loss.ctypes.data), cpt)
    assert rc == 0
    # ---- restatement (float64 accumulation; the kernel is float32 FMA in k order: tolerance, not bits)
    hs = _forward(W, b, state.astype(np.float64))
    hn = _forward(W, b, nstate.astype(np.float64))
    ht = _forward(tW, tb, nstate.astype(np.float64))
    tol = dict(rtol=2e-4, atol=2e-5)
    for l in range(L):
        np.testing.assert_allclose(h[l], hs[l], err_msg=f"h[{l}]", **tol)
    np.testing.assert_allclose(qn, hn[-1], err_msg="qn", **tol)
    np.testing.assert_allclose(qtn, ht[-1], err_msg="qtn", **tol)
    # TD row math on the kernel's own q values (so that a near-tie in argmax cannot flip the comparison)
    sel = qn if mode == 0 else qtn
    best = np.argmax(sel, axis=1)                               # first maximum, like the kernel
    nxt = qtn[np.arange(batch), best]
    mask = np.where(terminal != 0, np.float32(0), np.float32(1))
    y_ref = reward + (gamma * mask) * nxt
    av_ref = h[-1][np.arange(batch), action]
    diff = av_ref - y_ref
    np.testing.assert_array_equal(y, y_ref.astype(np.float32))
    np.testing.assert_array_equal(av, av_ref)
    np.testing.assert_array_equal(ab, np.abs(diff))
    np.testing.assert_array_equal(loss, (diff * diff).astype(np.float32).astype(np.float64) * isw)
    g_ref = np.zeros((batch, A), np.float32)
    g_ref[np.arange(batch), action] = (2.0 * diff.astype(np.float64) * isw / batch).astype(np.float32)
    np.testing.assert_array_equal(dz[L - 1], g_ref)
    # backward-input chain from the kernel's own grad_q and activations
    cur = g_ref.astype(np.float64)
    for l in range(L - 1, 0, -1):
        cur = (cur @ W[l].astype(np.float64)) * (h[l - 1] > 0)
        np.testing.assert_allclose(dz[l - 1], cur, err_msg=f"dz[{l - 1}]", **tol)
    assert not any(np.isnan(x).any() for x in h + dz + [qn, qtn, y, av, ab, loss])
