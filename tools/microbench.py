#!/usr/bin/env python
"""Per-kernel timings on the GPU box: latency at the learner's batch sizes and saturated throughput at
the sizes the reference quotes in its own docstring (prioritized_buffer.py:32-40: build 10 M, look up
10 M, batched update 1 M).  Writes one JSON object to stdout / gpurun_out/microbench.json.

Every kernel is timed with CUDA events on the launching stream.  Small launches are captured into a CUDA
graph (16 launches with different, really sampled inputs) so that the figure is device time per launch,
not Python launch overhead; inputs exceed L2 (268 MB tree, multi-GB ring) and are addressed randomly.
"""
import json
import os
import sys
import time

import numpy as np
import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from machin_b200 import _lib, ops  # noqa: E402
from machin_b200._lib import lib  # noqa: E402
from machin_b200.frame.buffers import PrioritizedBuffer, WeightTree  # noqa: E402

dev = t.device("cuda", 0)
PEAK = 6575.1
pp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pp):
    PEAK = float(json.load(open(pp))["hbm_gbs"])


def timed(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    t.cuda.synchronize()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    t.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3  # seconds per call


def graphed(fns, reps=20):
    """fns: list of thunks (one launch each, different inputs) -> seconds per launch."""
    side = t.cuda.Stream()
    side.wait_stream(t.cuda.current_stream())
    with t.cuda.stream(side):
        fns[0]()
    t.cuda.current_stream().wait_stream(side)
    t.cuda.synchronize()
    g = t.cuda.CUDAGraph()
    with t.cuda.graph(g):
        for f in fns:
            f()
    return timed(g.replay, reps=reps) / len(fns)


def main():  # noqa: C901
    from bench import ClockSampler          # nvidia-smi clocks / throttle reasons sampled during the whole run
    sampler = ClockSampler(0)
    sampler.start()
    out = {"peak_hbm_gbs": PEAK, "results": {}}
    R = out["results"]
    N = 10_000_000
    rng = np.random.RandomState(0)
    tree = WeightTree(N)
    D = tree.depth
    leaves = t.from_numpy((rng.exponential(size=N) + 1e-2) ** 0.6).to(dev)
    sp = lambda: _lib.stream_ptr()

    # ---- build 10 M (reference docstring: 90 ms on i7-6700HQ)
    s = timed(lambda: tree.update_all_leaves(leaves), reps=5)
    byt = 16 * (1 << (D - 1))
    R["tree_build_10M"] = {"ms": s * 1e3, "alg_bytes": byt, "GBps": byt / s / 1e9, "frac": byt / s / 1e9 / PEAK}
    total = tree.get_weight_sum()

    # ---- find: 10 M queries (reference docstring: 230 ms), 1 M, 4096, 512
    for nq in (10_000_000, 1_000_000, 4096, 512):
        qs = [t.rand(nq, dtype=t.float64, device=dev) * total for _ in range(4 if nq > 100000 else 16)]
        o = t.empty(nq, dtype=t.int64, device=dev)
        mk = lambda q: (lambda: _lib.check(lib.mb_tree_find(tree.handle, _lib.ptr(q), nq, _lib.ptr(o), sp())))
        s = timed(mk(qs[0]), reps=5) if nq > 100000 else graphed([mk(q) for q in qs])
        byt = (8 * (D - 1) + 16) * nq
        R[f"tree_find_{nq}"] = {"us": s * 1e6, "Mlookups_per_s": nq / s / 1e6, "alg_bytes": byt,
                                "GBps": byt / s / 1e9, "frac": byt / s / 1e9 / PEAK}

    # ---- batched update: 1 M (reference docstring: 20 ms), 65536, 4096, 512 (small kernel)
    for nu in (1_000_000, 65536, 4096, 512):
        sets = [(t.randint(0, N, (nu,), device=dev), t.rand(nu, device=dev)) for _ in range(4 if nu > 5000 else 16)]
        mk = lambda iw: (lambda: _lib.check(lib.mb_tree_update_priority(
            tree.handle, _lib.ptr(iw[0]), _lib.ptr(iw[1]), _lib.MB_F32, None, nu, 1e-2, 0.6, sp())))
        s = timed(mk(sets[0]), reps=5) if nu > 5000 else graphed([mk(x) for x in sets])
        byt = (20 + 24 * (D - 1)) * nu
        R[f"tree_update_{nu}"] = {"us": s * 1e6, "Mleaves_per_s": nu / s / 1e6, "alg_bytes_upper": byt,
                                  "GBps": byt / s / 1e9, "frac": byt / s / 1e9 / PEAK}
    # the learner's write-back: ascending indexes (what the stratified draw produces) -> tree_update_sorted
    # (or its cluster variant under MB_SORTED_CLUSTER=1)
    for nu in (512, 256, 128):
        sets = [(t.sort(t.randint(0, N, (nu,), device=dev))[0], t.rand(nu, device=dev)) for _ in range(16)]
        mk = lambda iw: (lambda: _lib.check(lib.mb_tree_update_priority(
            tree.handle, _lib.ptr(iw[0]), _lib.ptr(iw[1]), _lib.MB_F32, None, nu, 1e-2, 0.6, sp())))
        s = graphed([mk(x) for x in sets])
        byt = (20 + 24 * (D - 1)) * nu
        R[f"tree_update_sorted_{nu}"] = {"us": s * 1e6, "alg_bytes_upper": byt, "GBps": byt / s / 1e9,
                                         "frac": byt / s / 1e9 / PEAK,
                                         "variant": "tree_update_part (one CTA per subtree)" if os.environ.get("MB_UPDATE_LEGACY") != "1"
                                         else "legacy single-CTA sorted walk"}
    tree.check()

    # ---- fused PER draw
    for nb in (512, 1024, 4096, 1 << 20):
        us = [t.rand(nb, dtype=t.float64, device=dev) for _ in range(16 if nb <= 4096 else 2)]
        oi = t.empty(nb, dtype=t.int64, device=dev)
        ow = t.empty(nb, dtype=t.float64, device=dev)
        mk = lambda u: (lambda: _lib.check(lib.mb_per_sample(tree.handle, _lib.ptr(u), nb, N, 0.0, 0.4, None,
                                                              _lib.ptr(oi), _lib.ptr(ow), sp())))
        s = graphed([mk(u) for u in us]) if nb <= 4096 else timed(mk(us[0]), reps=5)
        byt = (8 * (D - 1) + 32) * nb
        R[f"per_sample_{nb}"] = {"us": s * 1e6, "Msamples_per_s": nb / s / 1e6, "alg_bytes": byt,
                                 "GBps": byt / s / 1e9, "frac": byt / s / 1e9 / PEAK}
    del tree, leaves
    t.cuda.empty_cache()

    # ---- row gather (C2 row 685 B, C4 row 3081 B)
    for label, sd, ad, adt, rows in (("C2", 84, 1, t.int64, 4_000_000), ("C4", 376, 17, t.float32, 2_000_000)):
        buf = PrioritizedBuffer(buffer_size=rows)
        gen = t.Generator(device=dev).manual_seed(0)
        act = t.randint(0, 6, (rows, ad), device=dev) if adt == t.int64 else t.rand(rows, ad, device=dev)
        buf.store_episode_tensors(
            {"state": {"state": t.randn(rows, sd, device=dev)}, "action": {"action": act},
             "next_state": {"state": t.randn(rows, sd, device=dev)}},
            {"reward": t.randn(rows, 1, device=dev), "terminal": t.rand(rows, 1, device=dev) < 0.01},
            priorities=t.rand(rows, device=dev, dtype=t.float64))
        payload = sum(f.nbytes for f in buf.storage.fields.values())
        for nb in (512, 4096, 1 << 18):
            idxs = [t.randint(0, rows, (nb,), device=dev) for _ in range(16 if nb <= 4096 else 2)]
            names = list(buf.storage.fields)
            dense = buf.storage.gather(idxs[0])
            arr = buf.storage._field_array(names, dense)
            mk = lambda ix: (lambda: _lib.check(lib.mb_rows_gather(_lib.ptr(buf.storage.data), buf.storage.pitch,
                                                                    rows, _lib.ptr(ix), nb, arr, len(names), sp())))
            s = graphed([mk(ix) for ix in idxs]) if nb <= 4096 else timed(mk(idxs[0]), reps=5)
            byt = 2 * payload * nb
            R[f"rows_gather_{label}_{nb}"] = {"us": s * 1e6, "Mrows_per_s": nb / s / 1e6, "row_payload": payload,
                                              "pitch": buf.storage.pitch, "alg_bytes": byt,
                                              "GBps": byt / s / 1e9, "frac": byt / s / 1e9 / PEAK}
        del buf
        t.cuda.empty_cache()

    # ---- V-trace (C3: 1024 x 20; C5: 4096 x 20; saturated: 2^22 episodes x 20)
    for n_ep, T in ((1024, 20), (4096, 20), (1 << 22, 20)):
        S = n_ep * T
        v, nv, r = (t.randn(S, device=dev) for _ in range(3))
        lp, lm = -t.rand(S, device=dev), -t.rand(S, device=dev)
        te = t.zeros(S, dtype=t.bool, device=dev)
        te[T - 1::T] = True
        fn = lambda: ops.vtrace(v, nv, r, te, lp, lm, 0.99, 1.0, 1.0, n_episodes=n_ep)
        s = graphed([fn] * 8) if S < 1e6 else timed(fn, reps=5)
        byt = 37 * S
        R[f"vtrace_{n_ep}x{T}"] = {"us": s * 1e6, "Msteps_per_s": S / s / 1e6, "alg_bytes": byt,
                                   "GBps": byt / s / 1e9, "frac": byt / s / 1e9 / PEAK}
        rd = r.double()
        fn = lambda: ops.gae(rd, te, v, nv, 0.99, 0.95, n_episodes=n_ep)
        s = graphed([fn] * 8) if S < 1e6 else timed(fn, reps=5)
        byt = 33 * S
        R[f"gae_{n_ep}x{T}"] = {"us": s * 1e6, "Msteps_per_s": S / s / 1e6, "alg_bytes": byt,
                                "GBps": byt / s / 1e9, "frac": byt / s / 1e9 / PEAK}

    # ---- TD
    for nb in (512, 1 << 20):
        q = [t.randn(nb, 6, device=dev) for _ in range(3)]
        a = t.randint(0, 6, (nb, 1), device=dev)
        rw = t.randn(nb, 1, device=dev)
        te = t.rand(nb, 1, device=dev) < 0.1
        w = t.rand(nb, dtype=t.float64, device=dev)
        fn = lambda: ops.td_dqn(q[0], q[1], q[2], a, rw, te, 0.99, is_weight=w, want_grad=True)
        s = graphed([fn] * 8) if nb < 1e5 else timed(fn, reps=5)
        byt = 137 * nb
        R[f"td_dqn_{nb}"] = {"us": s * 1e6, "alg_bytes": byt, "GBps": byt / s / 1e9, "frac": byt / s / 1e9 / PEAK}
    # ---- RAINBOW value block (written after round 1's GPU budget: first numbers come from the next round)
    for nb in (512, 1 << 18):
        A, N = 6, 51
        qd = [t.softmax(t.randn(nb, A, N, device=dev), dim=-1) for _ in range(3)]
        a = t.randint(0, A, (nb, 1), device=dev)
        val = t.randn(nb, 1, device=dev) * 5
        te = t.rand(nb, 1, device=dev) < 0.1
        w = t.rand(nb, dtype=t.float64, device=dev)
        fn = lambda: ops.td_rainbow(qd[0], qd[1], qd[2], a, val, te, -10.0, 10.0, 0.99 ** 3, is_weight=w, want_grad=True)
        s = graphed([fn] * 8) if nb < 1e5 else timed(fn, reps=5)
        byt = (4 * A * N * 4 + N * 4 + 30) * nb
        R[f"td_rainbow_{nb}"] = {"us": s * 1e6, "alg_bytes": byt, "GBps": byt / s / 1e9, "frac": byt / s / 1e9 / PEAK}
    out["clocks"] = sampler.stop()
    out["env"] = {"force_generic": os.environ.get("MB_FORCE_GENERIC_UPDATE", "0"), "gpu": t.cuda.get_device_name(0),
                  "rows_legacy": os.environ.get("MB_ROWS_LEGACY", "0"), "update_legacy": os.environ.get("MB_UPDATE_LEGACY", "0")}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
