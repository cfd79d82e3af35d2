#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- randomised runs of the kernels' device source under CPU emulation (tests/cpu_emul), beyond the
fixed cases of the CPU suite.  Takes minutes; not collected by pytest.

    python tools/fuzz_emulated.py rainbow [n]    # td_dist_kernel vs oracle/rainbow.py, and the oracle vs eager CPU torch
    python tools/fuzz_emulated.py tree [n]       # the sorted-update variants vs the generic update body
    python tools/fuzz_emulated.py rows [n]       # dqn_rows_kernel vs its NumPy restatement, random network shapes
    python tools/fuzz_emulated.py scan [n]       # V-trace / GAE, episode-length mixes around the staging-tile boundaries
    python tools/fuzz_emulated.py exchange [n]   # the fixed-shape shard exchange for random shard counts / pitches / counts

DESIGN.md 8b quotes the first runs made in the build container (300 + 400, 240, 160, 284 launches); later runs of this
tool: rainbow 1500, tree 150 (x 4 variants), rows 100 (x 4 mode / layout combinations), scan 150, exchange 300 -- no
mismatch in any of them."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
GXX = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread",
       "-I" + os.path.join(ROOT, "machin_b200", "csrc"), "-I" + os.path.join(ROOT, "include")]


def build(harness, shared=True):
    out = os.path.join(tempfile.mkdtemp(prefix="mb_fuzz_"), harness + (".so" if shared else ""))
    subprocess.run(GXX + (["-shared", "-fPIC"] if shared else []) +
                   [os.path.join(ROOT, "tests", "cpu_emul", harness + "_harness.cpp"), "-o", out], check=True)
    return C.CDLL(out) if shared else out


def plain(fn):
    return getattr(fn, "__wrapped__", fn)


def fuzz_rainbow(n):
    import torch as t
    import test_rainbow_cpu as T
    from oracle import rainbow as orb
    lib, bad = build("td"), 0
    soft = lambda x: np.exp(x) / np.exp(x).sum(-1, keepdims=True)
    for seed in range(n):
        rng = np.random.RandomState(seed)
        B, A, N = int(rng.randint(1, 12)), int(rng.randint(1, 7)), int(rng.choice([2, 3, 5, 11, 21, 51, 64]))
        v_min = float(rng.choice([-10, -2, 0, -1.5, -100]))
        v_max = v_min + float(rng.choice([1, 5, 20, 3.5, 200]))
        qd, nq, tq = (soft(rng.choice([0.1, 1, 5]) * rng.randn(B, A, N)).astype(np.float32) for _ in range(3))
        action = rng.randint(0, A, size=B)
        delta = (v_max - v_min) / (N - 1)
        value = np.where(rng.rand(B) < 0.5, v_min + delta * rng.randint(-2, N + 2, size=B), rng.randn(B) * (v_max - v_min))
        terminal, isw = rng.rand(B) < 0.5, rng.rand(B) + 0.1
        disc = float(rng.choice([0.99 ** 3, 0.9, 1.0, 0.5 ** 20]))
        got = T.run_kernel(lib, qd, nq, tq, action, value, terminal, isw, v_min, v_max, disc)["target_dist"]
        nd = tq[np.arange(B), orb.act_discrete(nq, orb.support(v_min, v_max, N))]
        want = orb.project(value.astype(np.float32), terminal, nd, v_min, v_max, disc)
        # the same projection as eager CPU torch tensor ops (what the oracle restates scalar by scalar)
        tv, tt, tn = t.from_numpy(value.astype(np.float32)).view(B, 1), t.from_numpy(terminal).view(B, 1), t.from_numpy(nd)
        b = ((tv + disc * ~tt * t.linspace(v_min, v_max, N).unsqueeze(0)).clamp(v_min + 1e-6, v_max - 1e-6) - v_min) / (
            (v_max - v_min) / (N - 1.0))
        lo, up = b.floor(), b.ceil()
        off = t.arange(0, B * N, N).view(-1, 1).expand(B, N).flatten()
        eager = t.zeros(B * N)
        eager.index_add_(0, lo.long().view(-1) + off, ((up - b) * tn).view(-1))
        eager.index_add_(0, up.long().view(-1) + off, ((b - lo) * tn).view(-1))
        if not (np.array_equal(got, want) and np.array_equal(want, eager.view(B, N).numpy())):
            bad += 1
            print("MISMATCH rainbow seed", seed, B, A, N, v_min, v_max, disc)
    return bad


def fuzz_tree(n):
    import test_tree_kernels_emulated_cpu as T
    exe, tmp, bad = build("tree", shared=False), tempfile.mkdtemp(), 0
    for seed in range(n):
        rng = np.random.RandomState(7000 + seed)
        size, m = int(rng.choice([130, 1000, 4097, 1 << 14, 70000, 1 << 17])), int(rng.randint(1, 257))
        tree, _ = T._tree(size, seed)
        kind = rng.choice(["uniform", "dups", "narrow"])
        if kind == "uniform":
            idx = rng.randint(0, size, size=m)
        elif kind == "dups":
            idx = rng.randint(0, max(2, size // 64), size=m) * int(rng.randint(1, 64)) % size
        else:
            idx = np.clip(int(rng.randint(0, size)) + rng.randint(0, max(2, m // 2), size=m), 0, size - 1)
        idx = np.sort(idx).astype(np.int64)
        w = rng.exponential(size=m).astype(np.float32 if rng.rand() < 0.5 else np.float64)
        valid = (rng.rand(m) < 0.8).astype(np.uint8) if rng.rand() < 0.4 else None
        norm = (0.01, 0.6) if rng.rand() < 0.5 else None
        want = T.run_kernel(exe, tmp, "generic", tree, idx, w, valid, norm)
        for variant in ("v2_async", "v2_cluster", "v2_cluster_async", "sorted"):
            got = T.run_kernel(exe, tmp, variant, tree, idx, w, valid, norm)
            if not (np.array_equal(got[0], want[0]) and got[1:] == want[1:]):
                bad += 1
                print("MISMATCH tree seed", seed, variant, size, m, kind)
    return bad


def fuzz_rows(n):
    import test_rows_kernel_emulated_cpu as T
    lib, bad, fn = build("rows"), 0, plain(T.test_rows_kernel_source_under_emulation)
    for seed in range(n):
        rng = np.random.RandomState(9000 + seed)
        layers = int(rng.randint(1, 5))
        sizes = ([int(rng.randint(1, 130))] + [int(rng.choice([1, 2, 3, 5, 8, 16, 31, 64, 100, 128, 200, 256])) for _ in range(layers - 1)]
                 + [int(rng.randint(1, 20))])
        batch = int(rng.randint(1, 41))
        for mode in (0, 1):
            for cpt in (4, 2):
                try:
                    fn(lib, sizes, batch, mode, cpt)
                except AssertionError as e:
                    bad += 1
                    print("MISMATCH rows", sizes, batch, mode, cpt, str(e)[:160])
    return bad


def fuzz_scan(n):
    import test_scan_kernels_emulated_cpu as T
    lib, bad = build("scan"), 0
    vtrace, gae = plain(T.test_vtrace_kernel_source_under_emulation), plain(T.test_gae_kernel_source_under_emulation)
    for seed in range(n):
        rng = np.random.RandomState(4000 + seed)
        kind = rng.choice(["short", "mixed", "long", "fixed"])
        if kind == "short":
            lengths = list(rng.randint(1, 45, size=int(rng.randint(1, 500))))
        elif kind == "mixed":
            lengths = list(rng.choice([1, 2, 20, 100, 700, 1663, 1664, 1665, 2975, 2976, 2977, 5000], size=int(rng.randint(1, 12))))
        elif kind == "long":
            lengths = [int(rng.randint(2000, 12000)) for _ in range(int(rng.randint(1, 4)))]
        else:
            lengths = [int(rng.choice([1, 5, 20, 64, 200, 1000, 3000]))] * int(rng.randint(1, 200))
        if sum(lengths) > 60000:
            lengths = lengths[:max(1, len(lengths) // 4)]
        lengths = [int(x) for x in lengths]
        for know in (False, True):
            for use_offsets in ((True, False) if len(set(lengths)) == 1 else (True,)):
                try:
                    vtrace(lib, lengths, seed, use_offsets, know)
                except AssertionError as e:
                    bad += 1
                    print("MISMATCH vtrace", kind, lengths[:6], len(lengths), know, use_offsets, str(e)[:120])
            try:
                gae(lib, float(rng.choice([1.0, 0.0, 0.95])), lengths, know)
            except AssertionError as e:
                bad += 1
                print("MISMATCH gae", kind, lengths[:6], len(lengths), know, str(e)[:120])
    return bad


def fuzz_exchange(n):
    """The fixed-shape shard exchange (gather into wire rows -> metadata -> all-to-all -> compaction -> priority scatter ->
    all-to-all -> version mask) replayed for random shard counts, row pitches and apportioned counts: the test function of
    the suite with its random stream shifted."""
    import test_exchange_emulated_cpu as T
    lib, bad = build("wire"), 0
    exchange = plain(T.test_fixed_shape_exchange_replayed_for_P_shards)
    apportion = plain(T.test_device_apportioning_matches_host_apportioning)
    real = np.random.RandomState
    for seed in range(n):
        rng = real(12000 + seed)
        P, pitch = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 8])), 16 * int(rng.randint(1, 60))
        T.np.random.RandomState = lambda s, _k=seed: real(s + 7919 * (_k + 1))
        try:
            exchange(lib, P, pitch)
            apportion(lib, P)
        except AssertionError as e:
            bad += 1
            print("MISMATCH exchange seed", seed, P, pitch, str(e)[:160])
        finally:
            T.np.random.RandomState = real
    return bad


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "rainbow"
    count = int(sys.argv[2]) if len(sys.argv) > 2 else {"rainbow": 300, "tree": 60, "rows": 40, "scan": 60, "exchange": 100}[which]
    mismatches = {"rainbow": fuzz_rainbow, "tree": fuzz_tree, "rows": fuzz_rows, "scan": fuzz_scan,
                  "exchange": fuzz_exchange}[which](count)
    print(f"{which}: {count} random configurations, {mismatches} mismatches")
    sys.exit(1 if mismatches else 0)
