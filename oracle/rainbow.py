"""TEST INFRASTRUCTURE ONLY -- restatement of the reference RAINBOW learner math.

Follows machin/frame/algorithms/rainbow.py: n-step values at store time (:183-196), acting by the expectation
over the support (:110-133), and the categorical projection + cross-entropy of ``update`` (:212-306).  Written
as explicit float32 scalar arithmetic (NumPy), one IEEE operation per reference tensor op, so that the order of
the roundings is visible -- the projected distribution is compared BIT FOR BIT with fixtures of the unmodified
reference (tests/golden/ref_rainbow.npz, made by oracle/make_golden.py::gen_rainbow).

Reference behaviour reproduced, not repaired:
* a Bellman target that lands exactly on an atom has floor == ceil, so BOTH neighbour weights are 0 and its
  probability mass is dropped (:257-262; only the two ends are protected by the 1e-6 clamp);
* ``value_loss`` has shape [B] when it is multiplied by ``is_weight.view([B, 1])`` (:303), so the "weighted" loss
  is the mean of a [B, B] outer product: mean(row loss) x mean(is_weight).
"""
import numpy as np

F = np.float32


def nstep_values(rewards, discount, n_steps):
    """rainbow.py:186-196: truncated n-step discounted sums, Python float arithmetic in the reference's order."""
    out, length = [], len(rewards)
    for i in range(length):
        value_sum = 0
        for j in reversed(range(min(n_steps, length - i))):
            value_sum = value_sum * discount + rewards[i + j]
        out.append(value_sum)
    return out


def support(v_min, v_max, atoms):
    """t.linspace(v_min, v_max, atoms) in float32 (:227).  Taken from torch itself: its CPU kernel evaluates
    start + i * step (first half) / end - (steps - 1 - i) * step (second half) with fused multiply-adds, which
    NumPy cannot express; the learner under test obtains its support from the same call."""
    import torch as t
    return t.linspace(v_min, v_max, atoms).numpy()


def expected_q(q_dist, z):
    """sum(support * q_dist, dim=-1) (:121), accumulated atom by atom in float32."""
    acc = np.zeros(q_dist.shape[:-1], dtype=F)
    for k in range(q_dist.shape[-1]):
        acc = (acc + (z[k] * q_dist[..., k]).astype(F)).astype(F)
    return acc


def act_discrete(q_dist, z):
    return np.argmax(expected_q(q_dist, z), axis=1)            # first maximum, like t.argmax


def project(value, terminal, next_dist, v_min, v_max, discount_n):
    """:236-283.  value f32[B], terminal bool[B], next_dist f32[B, N] (target net, greedy online action) ->
    target_dist f32[B, N]."""
    B, N = next_dist.shape
    z = support(v_min, v_max, N)
    lo, hi = F(v_min + 1e-6), F(v_max - 1e-6)                   # clamp bounds become float32 scalars
    delta_z = F((v_max - v_min) / (N - 1.0))
    vmin = F(v_min)
    gamma = F(discount_n)
    out = np.zeros((B, N), dtype=F)
    for i in range(B):
        cont = F(gamma * F(0.0 if terminal[i] else 1.0))        # discount * ~terminal (dqn.py:501)
        lw, uw, li, ui = [], [], [], []
        for j in range(N):
            tz = F(F(value[i]) + F(cont * z[j]))
            tz = min(max(tz, lo), hi)
            b = F(F(tz - vmin) / delta_z)
            lower, upper = F(np.floor(b)), F(np.ceil(b))
            li.append(int(lower)), ui.append(int(upper))
            lw.append(F(F(upper - b) * next_dist[i, j]))        # l_weight = u_dist * p
            uw.append(F(F(b - lower) * next_dist[i, j]))        # u_weight = l_dist * p
        for j in range(N):                                      # index_add_ #1, element order
            out[i, li[j]] = F(out[i, li[j]] + lw[j])
        for j in range(N):                                      # index_add_ #2
            out[i, ui[j]] = F(out[i, ui[j]] + uw[j])
    return out


def loss_rows(target_dist, q_sel):
    """-(target_dist * (q_dist + 1e-6).log()).sum(dim=1) (:291-293), float32."""
    terms = (target_dist * np.log((q_sel + F(1e-6)).astype(F)).astype(F)).astype(F)
    acc = np.zeros(terms.shape[0], dtype=F)
    for k in range(terms.shape[1]):
        acc = (acc + terms[:, k]).astype(F)
    return (-acc).astype(F)


def abs_error(rows):
    return (np.abs(rows) + F(1e-6)).astype(F)                   # :295


def weighted_loss(rows, is_weight):
    """:303 -- [B] * [B, 1] broadcasts to [B, B]; float64 because is_weight is float64."""
    return float((rows.astype(np.float64)[None, :] * np.asarray(is_weight, dtype=np.float64)[:, None]).mean())


def grad_q_sel(target_dist, q_sel, is_weight):
    """d weighted_loss / d q_dist[i, action_i, :] as autograd produces it: the float64 mean over the [B, B] product
    sends sum_j(is_weight_j) / B^2 back to every row, cast to float32 where the float32 row loss was promoted."""
    B = target_dist.shape[0]
    g = F(np.sum(np.asarray(is_weight, dtype=np.float64) * (1.0 / (B * B))))
    return ((-g * target_dist).astype(F) / (q_sel + F(1e-6)).astype(F)).astype(F)
