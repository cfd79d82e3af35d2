#!/usr/bin/env python
"""Launches EVERY kernel of the product paths eagerly (no CUDA graphs), at the BASELINE shapes, so that
`ncu -k regex:...` can capture each by name.  Used only under the profiler (tools/r2_ncu_call.sh):

  C2 step   per_sample_cluster, rows_gather_bulk_kernel, gemm_small_kernel (forward / input grad / weight grad),
            td_dqn_kernel, tree_update_part (sorted 512), adam_soft_kernel, adam_kernel, soft_update_multi_kernel
  exchange  per_sample_cluster_multi, rows_move_kernel (gather into wire rows, scatter), rows_pack_meta / unpack_meta,
            wire_compact, wire_scatter_update, wire_update_mask, wire_apportion, tree_update_part (4096 masked)
  scans     vtrace_kernel (C3 1024 x 20, C5 28 672 x 20... 4096 x 20), gae_kernel, td_target_kernel, td_dist_kernel
  tree      tree_build_pass / tree_build_top, tree_find_kernel (1 M queries), tree_leaf_weight_kernel
"""
import os
import sys

import numpy as np
import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from machin_b200 import _lib, ops  # noqa: E402
from machin_b200._lib import lib  # noqa: E402
from machin_b200.frame.algorithms import DQNApex, DQNPer  # noqa: E402
from machin_b200.frame.buffers import PrioritizedBuffer  # noqa: E402
from machin_b200.model import FusedAdam, FusedMLP  # noqa: E402
from machin_b200.parallel import LocalGroup  # noqa: E402

dev = t.device("cuda", 0)
N, B, S, A, H = 10_000_000, 512, 84, 6, 256
rows = int(os.environ.get("NCU_ROWS", "10000000"))
sp = _lib.stream_ptr


def fill(buf, n):
    done = 0
    while done < n:
        m = min(1_000_000, n - done)
        buf.store_episode_tensors(
            {"state": {"state": t.randn(m, S, device=dev)}, "action": {"action": t.randint(0, A, (m, 1), device=dev)},
             "next_state": {"state": t.randn(m, S, device=dev)}},
            {"reward": t.randn(m, 1, device=dev), "terminal": t.rand(m, 1, device=dev) < 0.01},
            priorities=t.empty(m, device=dev, dtype=t.float64).exponential_(1.0))
        done += m


# ---- C2 learner step, eager, FusedMLP nets (one kernel per layer per direction) + fused optimizer/target
buf = PrioritizedBuffer(buffer_size=N)
fill(buf, rows)
q, qt = FusedMLP([S, H, H, A]).to(dev), FusedMLP([S, H, H, A]).to(dev)
learner = DQNPer(q, qt, FusedAdam, t.nn.MSELoss(reduction="none"), batch_size=B, learning_rate=1e-3, discount=0.99,
                 update_rate=0.005, replay_buffer=buf)
np.random.seed(0)
for i in range(4):
    learner.fuse_optimizer_and_target = i >= 2          # both the two-launch and the one-launch optimizer + target update
    learner.update_device()
t.cuda.synchronize()
tree = buf.wt_tree.handle
# ---- tree kernels at the reference's docstring sizes
leaves = t.rand(N, dtype=t.float64, device=dev)
buf.wt_tree.update_all_leaves(leaves)
qs = t.rand(1_000_000, dtype=t.float64, device=dev) * buf.wt_tree.get_weight_sum()
oi = t.empty(1_000_000, dtype=t.int64, device=dev)
ow = t.empty(1_000_000, dtype=t.float64, device=dev)
_lib.check(lib.mb_tree_find(tree, _lib.ptr(qs), 1_000_000, _lib.ptr(oi), sp()))
_lib.check(lib.mb_tree_leaf_weight(tree, _lib.ptr(oi), 1_000_000, _lib.ptr(ow), sp()))
# ---- the fixed-shape shard exchange on a group of one at P = 8 requester blocks (same kernels, same shapes as N = 8)
P = 8
uni = t.rand((P, B), dtype=t.float64, device=dev)
sums = t.full((P,), float(buf.wt_tree.get_weight_sum()), dtype=t.float64, device=dev)
counts, scale = t.zeros(P, dtype=t.int64, device=dev), t.zeros(1, dtype=t.float64, device=dev)
params = t.zeros((P, 40), dtype=t.uint8, device=dev)
beta, nst = t.full((1,), 0.4, dtype=t.float64, device=dev), t.full((1,), rows, dtype=t.int64, device=dev)
idx, isw = t.zeros((P, B), dtype=t.int64, device=dev), t.zeros((P, B), dtype=t.float64, device=dev)
pitch = buf.storage.pitch
W = pitch + 24
send, recv = t.zeros((P * B, W), dtype=t.uint8, device=dev), t.zeros((P * B, W), dtype=t.uint8, device=dev)
merged = t.zeros((B + P, W), dtype=t.uint8, device=dev)
m_idx, m_ver, m_isw = (t.zeros(B + P, dtype=dt, device=dev) for dt in (t.int64, t.int64, t.float64))
send2, recv2 = t.zeros((P * B, 24), dtype=t.uint8, device=dev), t.zeros((P * B, 24), dtype=t.uint8, device=dev)
u_idx, u_pri, u_valid = t.zeros(P * B, dtype=t.int64, device=dev), t.zeros(P * B, dtype=t.float32, device=dev), t.zeros(P * B, dtype=t.uint8, device=dev)
table = t.zeros(N, dtype=t.int64, device=dev)
pri = t.rand(B + P, device=dev)
for _ in range(2):
    _lib.check(lib.mb_wire_apportion(_lib.ptr(sums), P, B, 0, B + P, _lib.ptr(nst), _lib.ptr(beta), 1e-3, _lib.ptr(counts),
                                     _lib.ptr(scale), _lib.ptr(params), sp()))
    _lib.check(lib.mb_per_sample_multi(tree, _lib.ptr(uni), B, P, _lib.ptr(params), _lib.ptr(idx), _lib.ptr(isw), sp()))
    buf.storage.gather_into_wire(idx.view(-1), send)
    _lib.check(lib.mb_rows_pack_meta(_lib.ptr(send), W, pitch, P * B, _lib.ptr(idx), None, _lib.ptr(table), _lib.ptr(isw), None, sp()))
    recv.copy_(send)
    _lib.check(lib.mb_wire_compact(_lib.ptr(recv), P, B, W, pitch, _lib.ptr(counts), _lib.ptr(merged), B + P, _lib.ptr(m_idx),
                                   _lib.ptr(m_ver), _lib.ptr(m_isw), _lib.ptr(scale), sp()))
    buf.storage.unpack(merged)
    _lib.check(lib.mb_wire_scatter_update(_lib.ptr(send2), P, B, _lib.ptr(counts), _lib.ptr(m_idx), _lib.ptr(m_ver), _lib.ptr(pri),
                                          B + P, sp()))
    recv2.copy_(send2)
    _lib.check(lib.mb_wire_update_mask(_lib.ptr(recv2), P, B, _lib.ptr(counts), 0, _lib.ptr(table), N, _lib.ptr(u_idx),
                                       _lib.ptr(u_pri), _lib.ptr(u_valid), sp()))
    PrioritizedBuffer.update_priority(buf, u_pri, u_idx, valid=u_valid)      # tree_update_part, 4096 masked entries
    v_idx, v_ver, v_isw = (t.empty(P * B, dtype=dt, device=dev) for dt in (t.int64, t.int64, t.float64))
    _lib.check(lib.mb_rows_unpack_meta(_lib.ptr(recv), W, pitch, P * B, _lib.ptr(v_idx), _lib.ptr(v_ver), _lib.ptr(v_isw), None, sp()))
t.cuda.synchronize()
buf.wt_tree.check()
# ---- the peer-memory exchange kernels on ONE GPU: every "peer" is this GPU's own buffer (the stores go to local HBM instead of
#      over NVLink; same code, same shapes as N = 8), the barrier with a group of one
import ctypes as C  # noqa: E402
W2 = pitch + 32
peer_recv = t.zeros((P * B, W2), dtype=t.uint8, device=dev)
peer_recv2 = t.zeros((P * B, 24), dtype=t.uint8, device=dev)
n_grad = sum(p.numel() for p in q.parameters())
flat = t.randn(n_grad, device=dev)
slots = t.zeros((P, n_grad), dtype=t.float32, device=dev)
flags, state = t.zeros(16, dtype=t.int64, device=dev), t.zeros(2, dtype=t.int64, device=dev)
same = lambda x: (C.c_void_p * P)(*[x.data_ptr()] * P)
one = lambda x: (C.c_void_p * 1)(x.data_ptr())
dense = {k: t.zeros((B + P,) + f.shape, dtype=f.dtype, device=dev) for k, f in buf.storage.fields.items()}
dense_arr = buf.storage._field_array(list(buf.storage.fields), dense)
recv_w2 = t.zeros((P * B, W2), dtype=t.uint8, device=dev)
for _ in range(2):
    _lib.check(lib.mb_wire_push_rows(_lib.ptr(buf.storage.data), pitch, buf.storage.max_size, _lib.ptr(idx), _lib.ptr(isw), _lib.ptr(table),
                                     P, B, 0, W2, same(peer_recv), None, None, None, sp()))
    _lib.check(lib.mb_peer_barrier(one(flags), 1, 0, _lib.ptr(state), 10 ** 9, sp()))
    _lib.check(lib.mb_wire_compact_fields(_lib.ptr(peer_recv), P, B, W2, pitch, _lib.ptr(counts), dense_arr, len(dense), B + P,
                                          _lib.ptr(m_idx), _lib.ptr(m_ver), _lib.ptr(m_isw), _lib.ptr(scale), None, 0, None, sp()))
    _lib.check(lib.mb_wire_push_update(P, B, 0, _lib.ptr(counts), _lib.ptr(m_idx), _lib.ptr(m_ver), _lib.ptr(pri), B + P, same(peer_recv2), sp()))
    _lib.check(lib.mb_peer_push(_lib.ptr(flat), 4 * n_grad, P, 0, 4 * n_grad, same(slots), None, None, None, sp()))
    _lib.check(lib.mb_peer_sum(_lib.ptr(slots), P, n_grad, n_grad, float(P), _lib.ptr(flat), None, 0, None, sp()))
    sums1, flags1, state1 = t.zeros(1, dtype=t.float64, device=dev), t.zeros(16, dtype=t.int64, device=dev), t.zeros(2, dtype=t.int64, device=dev)
    _lib.check(lib.mb_peer_sums_apportion(one(flags1), one(sums1), lib.mb_tree_root_ptr(tree), 1, 0, _lib.ptr(state1), 10 ** 9, B, B + 1,
                                          _lib.ptr(nst), _lib.ptr(beta), 1e-3, _lib.ptr(counts[:1]), _lib.ptr(scale), _lib.ptr(params[:1]), sp()))
t.cuda.synchronize()
# ---- scans and the other target kernels
for n_ep in (1024, 4096):
    T = n_ep * 20
    te = t.zeros(T, dtype=t.bool, device=dev)
    te[19::20] = True
    v, nv, r = (t.randn(T, device=dev) for _ in range(3))
    lp, lm = -t.rand(T, device=dev), -t.rand(T, device=dev)
    for _ in range(2):
        ops.vtrace(v, nv, r, te, lp, lm, 0.99, 1.0, 1.0, n_episodes=n_ep, max_len=20)
        ops.gae(r.double(), te, v, nv, 0.99, 0.95, n_episodes=n_ep, max_len=20)
te = t.rand(B, 1, device=dev) < 0.05
for _ in range(2):
    ops.td_target(t.randn(B, 1, device=dev), t.randn(B, 1, device=dev), te, 0.99, value=t.randn(B, 1, device=dev))
    d3 = [t.softmax(t.randn(B, A, 51, device=dev), dim=-1) for _ in range(3)]
    ops.td_rainbow(d3[0], d3[1], d3[2], t.randint(0, A, (B, 1), device=dev), t.randn(B, 1, device=dev) * 5, te, -10.0, 10.0,
                   0.99 ** 3, is_weight=t.rand(B, dtype=t.float64, device=dev), want_grad=True)
    ops.soft_update(qt, q, 0.005)
t.cuda.synchronize()
print("done")
