"""CPU: the DEVICE SOURCE of the fused Adam step + soft target update (machin_b200/csrc/optim_device.cuh) compiled with
g++ against tests/cpu_emul/cuda_shim.h, against torch.optim.Adam on the CPU followed by the reference's soft update
(utils.py:8-27: target * (1 - rate) + param * rate)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch as t

from conftest import ROOT

EMUL = os.path.join(ROOT, "tests", "cpu_emul")


@pytest.fixture(scope="module")
def optim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emul") / "optim_emul.so")
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread", "-shared", "-fPIC",
           "-I" + os.path.join(ROOT, "machin_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(EMUL, "optim_harness.cpp"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-3000:]
    return C.CDLL(out)


@pytest.mark.parametrize("weight_decay", [0.0, 0.01])
def test_adam_with_soft_update_matches_torch(optim, weight_decay):
    t.manual_seed(0)
    shapes = [(7, 5), (5,), (300, 3), (1,)]
    params = [t.randn(s) for s in shapes]
    targets = [p + 0.1 * t.randn_like(p) for p in params]
    ref_p = [p.clone().requires_grad_(True) for p in params]
    ref_t = [x.clone() for x in targets]
    opt = t.optim.Adam(ref_p, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
    P = [np.ascontiguousarray(p.numpy().copy()) for p in params]
    T = [np.ascontiguousarray(x.numpy().copy()) for x in targets]
    M, V = [np.zeros_like(p) for p in P], [np.zeros_like(p) for p in P]
    step = np.zeros(1, dtype=np.float32)
    ptrs = lambda arrs: (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    sizes = (C.c_int64 * len(P))(*[p.size for p in P])
    rate = 0.05
    for it in range(5):
        grads = [t.randn(s) for s in shapes]
        for p, g in zip(ref_p, grads):
            p.grad = g.clone()
        opt.step()
        with t.no_grad():
            for x, p in zip(ref_t, ref_p):
                x.copy_(x * (1.0 - rate) + p * rate)
        G = [np.ascontiguousarray(g.numpy().copy()) for g in grads]
        assert optim.emul_adam_soft(ptrs(P), ptrs(G), ptrs(M), ptrs(V), ptrs(T), sizes, len(P), C.c_void_p(step.ctypes.data),
                                    C.c_float(1e-2), C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8), C.c_float(weight_decay),
                                    C.c_double(rate)) == 0
        assert step[0] == it + 1
        for a, b in zip(P, ref_p):
            np.testing.assert_allclose(a, b.detach().numpy(), rtol=2e-6, atol=1e-7)
        for a, b in zip(T, ref_t):
            np.testing.assert_allclose(a, b.numpy(), rtol=2e-6, atol=1e-7)
