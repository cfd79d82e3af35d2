"""Functional CUDA-tensor front end of the C ABI (TD target math, soft update, V-trace, GAE).

Each function takes contiguous CUDA tensors, launches the corresponding kernel of
libmachin_b200.so on torch's current stream and returns CUDA tensors.  No CPU path.
"""
from typing import Dict, Optional

import torch as t

from . import _lib
from ._lib import lib


def _c(x: Optional[t.Tensor], dtype=None):
    if x is None:
        return None
    if not x.is_cuda:
        raise RuntimeError("machin_b200.ops works on CUDA tensors only (no CPU fallback)")
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    return x.contiguous()


def _u8(x: t.Tensor):
    x = _c(x)
    return x.view(t.uint8) if x.dtype == t.bool else x.to(t.uint8)


def td_dqn(q, q_next, q_tgt_next, action, reward, terminal, discount: float, is_weight=None,
           mode: str = "double", want_grad: bool = False) -> Dict[str, t.Tensor]:
    """Double-DQN target + priority math (dqn_per.py:105-139, dqn.py:497-501).

    q, q_next, q_tgt_next: f32[B, A]; action: int[B,1]; reward f32[B,1]; terminal bool[B,1];
    is_weight f64[B].  Returns y, action_value, abs_error (f32[B,1] / f32[B]), loss_elem f64[B,1]
    (MSE * is_weight) and, if ``want_grad``, grad_q = d mean(loss_elem) / d q."""
    q, q_next, q_tgt_next = _c(q, t.float32), _c(q_next, t.float32), _c(q_tgt_next, t.float32)
    B, A = q_tgt_next.shape
    dev = q_tgt_next.device
    action, reward = _c(action, t.int64), _c(reward, t.float32)
    term = _u8(terminal)
    isw = _c(is_weight, t.float64)
    y = t.empty((B, 1), dtype=t.float32, device=dev)
    av = t.empty((B, 1), dtype=t.float32, device=dev) if q is not None else None
    ae = t.empty((B,), dtype=t.float32, device=dev) if q is not None else None
    le = t.empty((B, 1), dtype=t.float64, device=dev) if q is not None else None
    gq = t.empty((B, A), dtype=t.float32, device=dev) if (want_grad and q is not None) else None
    with t.cuda.device(dev):
        _lib.check(lib.mb_td_dqn(_lib.ptr(q), _lib.ptr(q_next), _lib.ptr(q_tgt_next), B, A, _lib.ptr(action),
                                 _lib.ptr(reward), _lib.ptr(term), _lib.ptr(isw), float(discount),
                                 0 if mode == "double" else 1, _lib.ptr(y), _lib.ptr(av), _lib.ptr(ae),
                                 _lib.ptr(le), _lib.ptr(gq), _lib.stream_ptr()))
    return {"y": y, "action_value": av, "abs_error": ae, "loss_elem": le, "grad_q": gq}


_SUPPORT_CACHE: Dict[tuple, t.Tensor] = {}


def value_support(v_min: float, v_max: float, atoms: int, device) -> t.Tensor:
    """t.linspace(v_min, v_max, atoms) as the reference evaluates it -- float32 on the CPU (rainbow.py:119,227) --
    uploaded once per (range, size, device)."""
    key = (float(v_min), float(v_max), int(atoms), str(device))
    sup = _SUPPORT_CACHE.get(key)
    if sup is None:
        sup = _SUPPORT_CACHE[key] = t.linspace(v_min, v_max, atoms).to(device)
    return sup


def td_rainbow(q_dist, next_q_dist, target_next_q_dist, action, value, terminal, v_min: float, v_max: float,
               discount_n: float, is_weight=None, want_grad: bool = False) -> Dict[str, t.Tensor]:
    """Categorical TD target, row loss and priority of RAINBOW (rainbow.py:212-306) in one launch.

    q_dist / next_q_dist / target_next_q_dist: f32[B, A, N] soft-maxed distributions of the online net on s, the
    online net on s' and the target net on s'; action int[B,1]; value f32[B,1] (stored n-step value); terminal
    bool[B,1]; discount_n = discount ** reward_future_steps; is_weight f64[B].  Returns target_dist f32[B, N],
    loss_row f32[B], abs_error f32[B] (= |loss_row| + 1e-6), loss_elem f64[B] (its mean is the reference's scalar loss)
    and, if ``want_grad``, grad_q = d loss / d q_dist."""
    q_dist, next_q_dist, target_next_q_dist = (_c(x, t.float32) for x in (q_dist, next_q_dist, target_next_q_dist))
    B, A, N = q_dist.shape
    if next_q_dist.shape != q_dist.shape or target_next_q_dist.shape != q_dist.shape:
        raise ValueError("the three value distributions must have the same [batch, action_num, atom_num] shape")
    dev = q_dist.device
    action, value, term = _c(action, t.int64), _c(value, t.float32), _u8(terminal)
    isw = _c(is_weight, t.float64)
    sup = value_support(v_min, v_max, N, dev)
    # every row's CTA would walk the whole is_weight vector: beyond a few thousand rows, reduce it once up front
    scale = (isw.sum() * (1.0 / (B * B))).view(1) if (isw is not None and B > 4096) else None
    out = {"target_dist": t.empty((B, N), dtype=t.float32, device=dev), "loss_row": t.empty(B, dtype=t.float32, device=dev),
           "abs_error": t.empty(B, dtype=t.float32, device=dev), "loss_elem": t.empty(B, dtype=t.float64, device=dev),
           "grad_q": t.empty((B, A, N), dtype=t.float32, device=dev) if want_grad else None}
    with t.cuda.device(dev):
        _lib.check(lib.mb_td_rainbow(_lib.ptr(q_dist), _lib.ptr(next_q_dist), _lib.ptr(target_next_q_dist), B, A, N,
                                     _lib.ptr(sup), _lib.ptr(action), _lib.ptr(value), _lib.ptr(term), _lib.ptr(isw),
                                     _lib.ptr(scale), float(v_min), float(v_max), float(discount_n), _lib.ptr(out["target_dist"]),
                                     _lib.ptr(out["loss_row"]), _lib.ptr(out["abs_error"]), _lib.ptr(out["loss_elem"]),
                                     _lib.ptr(out["grad_q"]), _lib.stream_ptr()))
    return out


def td_target(next_value, reward, terminal, discount: float, value=None) -> Dict[str, t.Tensor]:
    """y = reward + discount * ~terminal * next_value (ddpg.py:506-510); abs_error = |value - y|."""
    next_value, reward = _c(next_value, t.float32), _c(reward, t.float32)
    B = reward.shape[0]
    dev = reward.device
    term = _u8(terminal)
    value = _c(value, t.float32)
    y = t.empty((B, 1), dtype=t.float32, device=dev)
    ae = t.empty((B,), dtype=t.float32, device=dev) if value is not None else None
    with t.cuda.device(dev):
        _lib.check(lib.mb_td_target(_lib.ptr(next_value), _lib.ptr(reward), _lib.ptr(term), B, float(discount),
                                    _lib.ptr(value), _lib.ptr(y), _lib.ptr(ae), _lib.stream_ptr()))
    return {"y": y, "abs_error": ae}


def soft_update_flat(target: t.Tensor, source: t.Tensor, rate: float):
    """In place: target = target * (1 - rate) + source * rate over a flat float32 block (utils.py:8-27)."""
    assert target.is_cuda and target.is_contiguous() and target.dtype == t.float32
    source = _c(source, t.float32)
    with t.cuda.device(target.device):
        _lib.check(lib.mb_soft_update(_lib.ptr(target), _lib.ptr(source), target.numel(), float(rate),
                                      _lib.stream_ptr()))


def soft_update(target_net: t.nn.Module, source_net: t.nn.Module, rate: float):
    """soft_update over all parameters of a network: one launch (mb_soft_update_multi)."""
    import ctypes as C
    tps, sps = list(target_net.parameters()), list(source_net.parameters())
    n = len(tps)
    if n == 0:
        return
    srcs = []
    for tp, sp in zip(tps, sps):
        if not (tp.is_cuda and tp.dtype == t.float32 and tp.is_contiguous()):
            raise RuntimeError("soft_update: parameters must be contiguous float32 CUDA tensors")
        sp = sp.data
        if sp.device != tp.device or sp.dtype != t.float32 or not sp.is_contiguous():
            sp = sp.to(device=tp.device, dtype=t.float32).contiguous()
        srcs.append(sp)
    tgt = (C.c_void_p * n)(*[tp.data_ptr() for tp in tps])
    src = (C.c_void_p * n)(*[sp.data_ptr() for sp in srcs])
    sizes = (C.c_int64 * n)(*[tp.numel() for tp in tps])
    with t.cuda.device(tps[0].device):
        _lib.check(lib.mb_soft_update_multi(tgt, src, sizes, n, float(rate), _lib.stream_ptr()))


def vtrace(value, next_value, reward, terminal, log_pi, log_mu, discount: float, clip_c: float = 1.0,
           clip_rho: float = 1.0, offsets: Optional[t.Tensor] = None, n_episodes: Optional[int] = None,
           max_len: int = 0) -> Dict[str, t.Tensor]:
    """IMPALA V-trace (impala.py:322-362) over flattened episode-major steps.  ``offsets``: int64
    [n_episodes + 1] CSR episode boundaries on the device, or None for equal-length episodes.  ``max_len``: length
    of the longest episode if the caller knows it (keeps the staged path for long episodes), 0 = unknown."""
    value = _c(value, t.float32)
    shape, dev = value.shape, value.device
    S = value.numel()
    args = [_c(x, t.float32) for x in (next_value, reward)]
    term = _u8(terminal)
    lp, lm = _c(log_pi, t.float32), _c(log_mu, t.float32)
    if offsets is not None:
        offsets = _c(offsets, t.int64)
        n_episodes = offsets.numel() - 1
    out = {k: t.empty(shape, dtype=t.float32, device=dev) for k in ("vs", "vss", "rho", "adv")}
    with t.cuda.device(dev):
        _lib.check(lib.mb_vtrace_len(_lib.ptr(offsets), int(n_episodes), S, int(max_len), _lib.ptr(value), _lib.ptr(args[0]),
                                 _lib.ptr(args[1]), _lib.ptr(term), _lib.ptr(lp), _lib.ptr(lm), float(discount),
                                 float(clip_c), float(clip_rho), _lib.ptr(out["vs"]), _lib.ptr(out["vss"]),
                                 _lib.ptr(out["rho"]), _lib.ptr(out["adv"]), _lib.stream_ptr()))
    return out


def gae(reward, terminal, critic_value, next_critic_value, discount: float, gae_lambda: float,
        offsets: Optional[t.Tensor] = None, n_episodes: Optional[int] = None, max_len: int = 0) -> Dict[str, t.Tensor]:
    """Discounted returns + GAE (a2c.py:275-313), float64 like the reference's Python floats.  ``max_len``: see
    :func:`vtrace`."""
    reward = _c(reward, t.float64)
    dev, S = reward.device, reward.numel()
    term = _u8(terminal)
    cv, ncv = _c(critic_value, t.float32), _c(next_critic_value, t.float32)
    if offsets is not None:
        offsets = _c(offsets, t.int64)
        n_episodes = offsets.numel() - 1
    out = {k: t.empty(reward.shape, dtype=t.float64, device=dev) for k in ("returns", "gae")}
    with t.cuda.device(dev):
        _lib.check(lib.mb_gae_len(_lib.ptr(offsets), int(n_episodes), S, int(max_len), _lib.ptr(reward), _lib.ptr(term), _lib.ptr(cv),
                              _lib.ptr(ncv), float(discount), float(gae_lambda), _lib.ptr(out["returns"]),
                              _lib.ptr(out["gae"]), _lib.stream_ptr()))
    return out
