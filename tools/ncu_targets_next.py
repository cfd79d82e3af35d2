#!/usr/bin/env python
"""Launches the kernels written after round 1's GPU budget was spent, at the C2 shapes, a few times each and without
CUDA graphs, so that `ncu -k regex:...` can capture them by name.  Used only under the profiler.

    MB_SORTED_CLUSTER=1 MB_SORTED_ASYNC=1 ncu --set full --clock-control none --import-source on \\
        -k regex:"tree_update_sorted_v2|dqn_rows_kernel|mlp_program_kernel|transpose_multi|per_sample_cluster_multi|wire_apportion" \\
        -c 24 -o gpurun_out/prof_next python tools/ncu_targets_next.py
"""
import os
import sys

import numpy as np
import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from machin_b200 import _lib  # noqa: E402
from machin_b200._lib import lib  # noqa: E402
from machin_b200.frame.algorithms import DQNPer  # noqa: E402
from machin_b200.frame.buffers import PrioritizedBuffer  # noqa: E402
from machin_b200.model import FusedAdam, FusedMLP  # noqa: E402

dev = t.device("cuda", 0)
N, B, S, A, H = 10_000_000, 512, 84, 6, 256
rows = int(os.environ.get("NCU_ROWS", "2000000"))
buf = PrioritizedBuffer(buffer_size=N)
buf.store_episode_tensors(
    {"state": {"state": t.randn(rows, S, device=dev)}, "action": {"action": t.randint(0, A, (rows, 1), device=dev)},
     "next_state": {"state": t.randn(rows, S, device=dev)}},
    {"reward": t.randn(rows, 1, device=dev), "terminal": t.rand(rows, 1, device=dev) < 0.01},
    priorities=t.rand(rows, device=dev, dtype=t.float64))
q, qt = FusedMLP([S, H, H, A]).to(dev), FusedMLP([S, H, H, A]).to(dev)
learner = DQNPer(q, qt, FusedAdam, t.nn.MSELoss(reduction="none"), batch_size=B, learning_rate=1e-3, discount=0.99,
                 update_rate=0.005, replay_buffer=buf)
learner.use_fused_program, learner.fused_kind = True, "rows"
np.random.seed(0)
for _ in range(4):                       # eager: per_sample, gather, transposes, rows kernel, dW program, tree update, adam, soft
    learner.update_device()
t.cuda.synchronize()
# the shard-exchange helpers at 8 requesters
P = 8
tree = buf.wt_tree.handle
uni = t.rand((P, B), dtype=t.float64, device=dev)
sums = t.full((P,), float(buf.wt_tree.get_weight_sum()), dtype=t.float64, device=dev)
counts, scale = t.zeros(P, dtype=t.int64, device=dev), t.zeros(1, dtype=t.float64, device=dev)
params = t.zeros((P, 40), dtype=t.uint8, device=dev)
beta, nst = t.full((1,), 0.4, dtype=t.float64, device=dev), t.full((1,), rows, dtype=t.int64, device=dev)
idx, isw = t.zeros((P, B), dtype=t.int64, device=dev), t.zeros((P, B), dtype=t.float64, device=dev)
for _ in range(3):
    _lib.check(lib.mb_wire_apportion(_lib.ptr(sums), P, B, 0, B + P, _lib.ptr(nst), _lib.ptr(beta), 1e-3, _lib.ptr(counts),
                                     _lib.ptr(scale), _lib.ptr(params), _lib.stream_ptr()))
    _lib.check(lib.mb_per_sample_multi(tree, _lib.ptr(uni), B, P, _lib.ptr(params), _lib.ptr(idx), _lib.ptr(isw),
                                       _lib.stream_ptr()))
# kernels that had no capture in round 1: returns / GAE, DDPG target, the scatter direction of the row mover
from machin_b200 import ops  # noqa: E402
T = 1024 * 20
te = t.zeros(T, dtype=t.bool, device=dev)
te[19::20] = True
for _ in range(2):
    ops.gae(t.randn(T, dtype=t.float64, device=dev), te, t.randn(T, device=dev), t.randn(T, device=dev), 0.99, 0.95, n_episodes=1024)
    ops.td_target(t.randn(B, 1, device=dev), t.randn(B, 1, device=dev), te[:B].view(B, 1), 0.99, value=t.randn(B, 1, device=dev))
    buf.store_episode_tensors(
        {"state": {"state": t.randn(4096, S, device=dev)}, "action": {"action": t.randint(0, A, (4096, 1), device=dev)},
         "next_state": {"state": t.randn(4096, S, device=dev)}},
        {"reward": t.randn(4096, 1, device=dev), "terminal": t.rand(4096, 1, device=dev) < 0.01},
        priorities=t.rand(4096, device=dev, dtype=t.float64))
# the RAINBOW value block at the learner batch (6 actions x 51 atoms)
AT = 51
dist = [t.softmax(t.randn(B, A, AT, device=dev), dim=-1) for _ in range(3)]
for _ in range(2):
    ops.td_rainbow(dist[0], dist[1], dist[2], t.randint(0, A, (B, 1), device=dev), t.randn(B, 1, device=dev) * 5, te[:B].view(B, 1),
                   -10.0, 10.0, 0.99 ** 3, is_weight=t.rand(B, dtype=t.float64, device=dev), want_grad=True)
t.cuda.synchronize()
buf.wt_tree.check()
print("done", counts.tolist())
