#!/usr/bin/env python
"""Device time of the learner's priority write-back (sorted batch into a 10 M-leaf tree) for the kernel variant
selected by MB_SORTED_CLUSTER / MB_SORTED_ASYNC.  CUDA events around graph replays of 16 launches with different,
really sorted inputs; prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from machin_b200 import _lib  # noqa: E402
from machin_b200._lib import lib  # noqa: E402
from machin_b200.frame.buffers import WeightTree  # noqa: E402


def main():
    dev = t.device("cuda", 0)
    N = 10_000_000
    rng = np.random.RandomState(0)
    tree = WeightTree(N)
    tree.update_all_leaves(t.from_numpy((rng.exponential(size=N) + 1e-2) ** 0.6).to(dev))
    legacy = os.environ.get("MB_UPDATE_LEGACY") == "1"
    out = {"variant": "one CTA per subtree (tree_update_part)" if not legacy else
           ("legacy " + ("cluster" if os.environ.get("MB_SORTED_CLUSTER") == "1" else "single-CTA") + "/" +
            ("async" if os.environ.get("MB_SORTED_ASYNC") == "1" else "barrier"))}
    # sorted batches (what the stratified draw returns), then merged shard blocks with a validity mask (the
    # sharded write-back: P ascending blocks of 512, ~1/8 of the rows masked out)
    for nu in (512, 256, 128, 65, 2048, 4096):
        merged = nu > 512
        if merged:
            sets = [(t.cat([t.sort(t.randint(0, N, (512,), device=dev))[0] for _ in range(nu // 512)]),
                     t.rand(nu, device=dev), (t.rand(nu, device=dev) < 0.875).to(t.uint8)) for _ in range(16)]
        else:
            sets = [(t.sort(t.randint(0, N, (nu,), device=dev))[0], t.rand(nu, device=dev), None) for _ in range(16)]

        def launch(iw):
            _lib.check(lib.mb_tree_update_priority(tree.handle, _lib.ptr(iw[0]), _lib.ptr(iw[1]), _lib.MB_F32,
                                                   _lib.ptr(iw[2]), nu, 1e-2, 0.6, _lib.stream_ptr()))
        side = t.cuda.Stream()
        side.wait_stream(t.cuda.current_stream())
        with t.cuda.stream(side):
            launch(sets[0])
        t.cuda.current_stream().wait_stream(side)
        t.cuda.synchronize()
        g = t.cuda.CUDAGraph()
        with t.cuda.graph(g):
            for x in sets:
                launch(x)
        for _ in range(3):
            g.replay()
        t.cuda.synchronize()
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            g.replay()
        e1.record()
        t.cuda.synchronize()
        out[f"us_{nu}" + ("_merged_masked" if merged else "")] = round(e0.elapsed_time(e1) * 1e3 / (50 * 16), 3)
    tree.check()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
