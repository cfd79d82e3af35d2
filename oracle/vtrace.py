"""TEST INFRASTRUCTURE ONLY -- restatement of the reference IMPALA V-trace block.

Follows machin/frame/algorithms/impala.py:322-373.  ``vtrace_loop`` is the literal
per-step scalar recurrence (what the reference executes, ~tens of microseconds per
step); ``vtrace_batched`` evaluates the same float32 expressions in the same order but
vectorised over episodes, for shapes where the loop takes too long.  Both leave vs = 0
on the last step of every episode (impala.py:341-357).
"""
import numpy as np
import torch as t


def vtrace_prepare(cur_log_prob, action_log_prob, value, next_value, reward, terminal,
                   discount, clip_c, clip_rho):
    # impala.py:322-332 ; all [sumL,1] float32 CPU tensors, terminal bool
    sim = t.exp(cur_log_prob - action_log_prob)
    c = t.min(t.full(sim.shape, clip_c, dtype=sim.dtype), sim)
    rho = t.min(t.full(sim.shape, clip_rho, dtype=sim.dtype), sim)
    next_value = next_value.clone()
    next_value[terminal] = 0
    delta_v = rho * (reward + discount * next_value - value)
    return c, rho, delta_v


def vtrace_loop(value, delta_v, c, lengths, discount):
    # impala.py:340-362
    vs = t.zeros_like(value)
    vss = t.zeros_like(value)
    start = 0
    for n in lengths:
        for k in range(n - 2, -1, -1):
            i = start + k
            vs[i] = value[i] + delta_v[i] + discount * c[i] * (vs[i + 1] - value[i + 1])
        vss[start : start + n - 1] = vs[start + 1 : start + n]
        start += n
    return vs, vss


def vtrace_batched(value, delta_v, c, lengths, discount):
    """Same arithmetic as vtrace_loop, vectorised across episodes (float32)."""
    v = value.detach().reshape(-1).numpy().astype(np.float32)
    d = delta_v.detach().reshape(-1).numpy().astype(np.float32)
    cc = c.detach().reshape(-1).numpy().astype(np.float32)
    lengths = np.asarray(lengths, dtype=np.int64)
    starts = np.concatenate(([0], np.cumsum(lengths)[:-1]))
    vs = np.zeros_like(v)
    g = np.float32(discount)
    for k in range(int(lengths.max()) - 2, -1, -1):
        live = lengths - 2 >= k
        i = starts[live] + k
        vs[i] = (v[i] + d[i]) + (g * cc[i]) * (vs[i + 1] - v[i + 1])
    vss = np.zeros_like(v)
    for s, n in zip(starts, lengths):
        vss[s : s + n - 1] = vs[s + 1 : s + n]
    return t.from_numpy(vs).view_as(value), t.from_numpy(vss).view_as(value)


def impala_losses(rho, cur_log_prob, reward, vss, value, vs, discount, entropy=None,
                  entropy_weight=None):
    # impala.py:364-373 (criterion = MSELoss(reduction="sum") in the reference tests)
    pol = -(rho.detach() * cur_log_prob * (reward + discount * vss - value).detach())
    if entropy_weight is not None:
        pol = pol + entropy_weight * entropy
    return pol.sum(), ((value - vs.type_as(value)) ** 2).sum()
