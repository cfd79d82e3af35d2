#!/usr/bin/env python
"""Launches each hot-path kernel a few times at the C2 shapes (10 M-leaf tree, 685 B rows, batch 512) so
that `ncu -k regex:<kernel>` can capture it.  Used only under the profiler; prints nothing measured."""
import os
import sys

import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from machin_b200 import _lib, ops  # noqa: E402
from machin_b200._lib import lib  # noqa: E402
from machin_b200.frame.buffers import PrioritizedBuffer  # noqa: E402

dev = t.device("cuda", 0)
N, B = 10_000_000, 512
rows = int(os.environ.get("NCU_ROWS", "2000000"))
buf = PrioritizedBuffer(buffer_size=N)
gen = t.Generator(device=dev).manual_seed(0)
buf.store_episode_tensors(
    {"state": {"state": t.randn(rows, 84, device=dev)}, "action": {"action": t.randint(0, 6, (rows, 1), device=dev)},
     "next_state": {"state": t.randn(rows, 84, device=dev)}},
    {"reward": t.randn(rows, 1, device=dev), "terminal": t.rand(rows, 1, device=dev) < 0.01},
    priorities=t.rand(rows, device=dev, dtype=t.float64))
buf.wt_tree.update_all_leaves(t.rand(N, device=dev, dtype=t.float64))
t.cuda.synchronize()
sp = lambda: _lib.stream_ptr()
for rep in range(4):
    u = t.rand(B, dtype=t.float64, device=dev)
    n, fields, idx, isw = buf.sample_device(B, uniform=u)
    idx = t.sort(idx % rows)[0]          # stays ascending: the write-back must take the sorted kernel, as in the learner
    fields = buf.storage.gather(idx)
    q = [t.randn(B, 6, device=dev) for _ in range(3)]
    out = ops.td_dqn(q[0], q[1], q[2], fields["action/action"], fields["reward"], fields["terminal"], 0.99,
                     is_weight=isw, want_grad=True)
    buf.update_priority(out["abs_error"], idx)
    big_i = t.randint(0, N, (1_000_000,), device=dev)
    buf.update_priority(t.rand(1_000_000, device=dev), big_i)
    qs = t.rand(1 << 20, dtype=t.float64, device=dev) * buf.wt_tree.get_weight_sum()
    buf.wt_tree.find_leaf_index(qs)
    S = 1024 * 20
    v, nv, r = (t.randn(S, device=dev) for _ in range(3))
    te = t.zeros(S, dtype=t.bool, device=dev)
    te[19::20] = True
    ops.vtrace(v, nv, r, te, -t.rand(S, device=dev), -t.rand(S, device=dev), 0.99, n_episodes=1024)
    t.cuda.synchronize()
print("done")
