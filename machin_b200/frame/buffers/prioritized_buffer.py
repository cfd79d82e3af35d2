"""Prioritized replay on the device: sum tree + packed rows, reference ``PrioritizedBuffer`` surface.

Mirror of machin/frame/buffers/prioritized_buffer.py:234-434.  Two call styles:

* the reference's own (NumPy in / NumPy out): ``sample_batch`` returns ``index`` and
  ``is_weight`` as ``np.ndarray`` and ``update_priority`` takes ``np.ndarray`` -- uniforms are
  drawn from the global NumPy MT19937 stream exactly like :397-416, so a seeded run samples the
  same slots as the reference;
* the device path used by machin_b200's own algorithms (:meth:`sample_device`,
  :meth:`update_priority` with CUDA tensors): everything stays in HBM, nothing synchronises.
"""
from typing import Dict, List, Tuple, Union

import numpy as np
import torch as t

from ... import _lib
from ..._lib import lib, PER_PARAMS_BYTES
from ..transition import TransitionBase
from .buffer import Buffer
from .weight_tree import WeightTree


class PrioritizedBuffer(Buffer):
    def __init__(self, buffer_size: int = 1000000, buffer_device: Union[str, t.device] = "cuda",
                 epsilon: float = 1e-2, alpha: float = 0.6, beta: float = 0.4,
                 beta_increment_per_sampling: float = 0.001, **kwargs):
        super().__init__(buffer_size=buffer_size, buffer_device=buffer_device, storage=None, **kwargs)
        self.epsilon = epsilon
        self.alpha = alpha
        self.beta = beta
        self.beta_increment_per_sampling = beta_increment_per_sampling
        self.curr_beta = beta
        self.wt_tree = WeightTree(int(buffer_size), self.buffer_device)

    # ------------------------------------------------------------------ store
    def _default_priority(self):
        # "the initialization method used in the original essay" (:301-303); note max_leaf is
        # already normalised, the reference normalises it again -- reproduced as is.
        return self._normalize_priority(self.wt_tree.get_leaf_max())

    def store_episode(self, episode: List[Union[TransitionBase, Dict]], priorities: Union[List[float], None] = None,
                      required_attrs=("state", "action", "next_state", "reward", "terminal")):
        positions = super().store_episode(episode, required_attrs)
        if priorities is None:
            self.wt_tree.update_leaf_batch([self._default_priority()] * len(positions), positions)
        else:
            self.wt_tree.update_leaf_batch(self._normalize_priority(priorities), positions)
        return positions

    def store_episode_tensors(self, major, sub, custom=None, priorities=None):
        """Tensor-native store.  ``priorities``: None, a NumPy array (normalised on the host like
        the reference) or a CUDA tensor (normalised on the device)."""
        positions = super().store_episode_tensors(major, sub, custom)
        if priorities is None:
            self.wt_tree.update_leaf_batch(np.full(len(positions), self._default_priority()), positions)
        elif t.is_tensor(priorities) and priorities.is_cuda:
            self.update_priority(priorities, t.as_tensor(positions).to(self.buffer_device))
        else:
            self.wt_tree.update_leaf_batch(self._normalize_priority(np.asarray(priorities)), positions)
        return positions

    def clear(self):
        super().clear()
        self.wt_tree.clear()          # the reference builds a fresh tree (:315-316)
        self.curr_beta = self.beta

    # ------------------------------------------------------------------ priorities
    def update_priority(self, priorities, indexes, valid=None):
        if t.is_tensor(priorities) and priorities.is_cuda:
            p = priorities.detach().contiguous().view(-1)
            if p.dtype not in (t.float32, t.float64):
                p = p.to(t.float64)
            i = (indexes if t.is_tensor(indexes) else t.as_tensor(np.asarray(indexes))).to(
                device=self.buffer_device, dtype=t.int64).contiguous().view(-1)
            if p.shape[0] != i.shape[0]:
                raise ValueError("Dimension of weights and indexes doesn't match!")
            v = None if valid is None else valid.to(dtype=t.uint8).contiguous().view(-1)
            _lib.check(lib.mb_tree_update_priority(
                self.wt_tree.handle, _lib.ptr(i), _lib.ptr(p), _lib.MB_F32 if p.dtype == t.float32 else _lib.MB_F64,
                _lib.ptr(v), i.shape[0], float(self.epsilon), float(self.alpha), _lib.stream_ptr()))
            return
        self.wt_tree.update_leaf_batch(self._normalize_priority(priorities), indexes)

    def _normalize_priority(self, priority):
        return (np.abs(priority) + self.epsilon) ** self.alpha   # host inputs: NumPy, like :430-434

    # ------------------------------------------------------------------ sample
    def sample_batch(self, batch_size: int, concatenate: bool = True, device: Union[str, t.device] = None,
                     sample_attrs: List[str] = None, additional_concat_custom_attrs: List[str] = None, *_, **__
                     ) -> Tuple[int, Union[None, tuple], Union[None, np.ndarray], Union[None, np.ndarray]]:
        if batch_size <= 0 or self.size() == 0:
            return 0, None, None, None
        d_index, d_isw = self._draw(batch_size)
        result = self.post_process_batch(None, device, concatenate, sample_attrs, additional_concat_custom_attrs,
                                         index_tensor=d_index)
        index = d_index.cpu().numpy()
        is_weight = d_isw.cpu().numpy()
        return len(index), result, index, is_weight

    def sample_index_and_weight(self, batch_size: int, all_weight_sum: float = None):
        d_index, d_isw = self._draw(batch_size, all_weight_sum)
        return d_index.cpu().numpy(), d_isw.cpu().numpy()

    # ------------------------------------------------------------------ CUDA-graph support
    def enable_graph_params(self):
        """Keep (beta, all_weight_sum, n_stored) in a device-side ``mb_per_params`` block so that the
        draw can be replayed inside a CUDA graph while they change.  ``curr_beta`` is advanced on the
        device by the captured step (:meth:`advance_beta_device`) and mirrored on the host."""
        if getattr(self, "_params_raw", None) is None:
            self._params_raw = t.zeros(PER_PARAMS_BYTES, dtype=t.uint8, device=self.buffer_device)
            self._params_f64 = self._params_raw.view(t.float64)
            self._params_i64 = self._params_raw.view(t.int64)
        self.sync_graph_params()

    def sync_graph_params(self):
        # every field of mb_per_params: beta, all_weight_sum (0 = read the root), n_stored, step, n_draw
        host = t.tensor([float(self.curr_beta), 0.0], dtype=t.float64).view(t.uint8)
        ints = t.tensor([len(self.storage), 0, 0], dtype=t.int64).view(t.uint8)
        self._params_raw.copy_(t.cat([host, ints]))

    def advance_beta_device(self):
        self._params_f64[0:1].add_(self.beta_increment_per_sampling).clamp_(max=1.0)

    def _stage_uniform(self, values: np.ndarray) -> t.Tensor:
        n = values.shape[0]
        ring = getattr(self, "_u_ring", None)
        if ring is None or ring["host"][0].shape[0] < n:
            cap = max(n, 1024)
            ring = {"host": [t.empty(cap, dtype=t.float64).pin_memory() for _ in range(8)],
                    "dev": [t.empty(cap, dtype=t.float64, device=self.buffer_device) for _ in range(8)],
                    "ev": [t.cuda.Event() for _ in range(8)], "i": 0}
            self._u_ring = ring
        i = ring["i"] = (ring["i"] + 1) % 8
        ring["ev"][i].synchronize()                       # the copy that last used this slot has finished
        ring["host"][i].numpy()[:n] = values
        dev = ring["dev"][i][:n]
        dev.copy_(ring["host"][i][:n], non_blocking=True)
        ring["ev"][i].record()
        return dev

    def _draw(self, batch_size: int, all_weight_sum: float = None, uniform: t.Tensor = None,
              use_device_params: bool = False, out=None):
        """Fused stratified draw on the device (csrc/sumtree.cu: per_sample_*).  ``out``: optional
        pre-allocated (index i64[n], is_weight f64[n]) tensors."""
        if use_device_params:
            d_index, d_isw = out if out is not None else (
                t.empty(batch_size, dtype=t.int64, device=self.buffer_device),
                t.empty(batch_size, dtype=t.float64, device=self.buffer_device))
            _lib.check(lib.mb_per_sample(self.wt_tree.handle, _lib.ptr(uniform), batch_size, 0, 0.0, 0.0,
                                         _lib.ptr(self._params_raw), _lib.ptr(d_index), _lib.ptr(d_isw),
                                         _lib.stream_ptr()))
            if not getattr(self, "defer_beta_advance", False):
                self.advance_beta_device()             # (a learner that forks a side branch advances beta there: it only
            return d_index, d_isw                       #  feeds the NEXT draw -- two small launches off the critical path)
        if uniform is None:
            # same call and same stream as the reference (:411): B doubles from np.random, staged through a
            # small ring of pinned buffers so that the copy is a true asynchronous H2D
            uniform = self._stage_uniform(np.random.uniform(size=batch_size))
        d_index, d_isw = out if out is not None else (
            t.empty(batch_size, dtype=t.int64, device=self.buffer_device),
            t.empty(batch_size, dtype=t.float64, device=self.buffer_device))
        _lib.check(lib.mb_per_sample(self.wt_tree.handle, _lib.ptr(uniform), batch_size, len(self.storage),
                                     float(all_weight_sum or 0.0), float(self.curr_beta), None,
                                     _lib.ptr(d_index), _lib.ptr(d_isw), _lib.stream_ptr()))
        self.curr_beta = np.float64(min(1.0, self.curr_beta + self.beta_increment_per_sampling))   # == np.min([1.0, ...]) (:425-427), 25x cheaper
        return d_index, d_isw

    def sample_device(self, batch_size: int, uniform: t.Tensor = None, names=None, use_device_params=False
                      ) -> Tuple[int, Dict[str, t.Tensor], t.Tensor, t.Tensor]:
        """Device path: returns (n, {field: tensor}, index i64[n], is_weight f64[n]) without any
        host synchronisation.  ``uniform``: optional f64[n] CUDA tensor of U[0,1) draws."""
        d_index, d_isw = self._draw(batch_size, uniform=uniform, use_device_params=use_device_params)
        return batch_size, self.storage.gather(d_index, names=names), d_index, d_isw
