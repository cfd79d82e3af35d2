"""Device-resident sum tree with the reference's WeightTree surface.

Mirror of machin/frame/buffers/prioritized_buffer.py:8-231.  Storage and every operation
live in libmachin_b200.so (CUDA, sm_100a); this class only routes arguments:

* NumPy arrays / lists / scalars  -> the ``*_host`` C entry points (copy in, kernel, copy out,
  synchronous) and NumPy / Python results, exactly the reference's return conventions;
* CUDA tensors                    -> the asynchronous device entry points on torch's current
  stream, CUDA tensors out, no host synchronisation (index validation is deferred, see
  :meth:`check`).
"""
import ctypes as C
from typing import Any, List, Union

import numpy as np
import torch as t

from ... import _lib
from ..._lib import lib


def _is_cuda(x) -> bool:
    return t.is_tensor(x) and x.is_cuda


class WeightTree:
    def __init__(self, size: int, device: Union[str, t.device, int] = "cuda"):
        dev = t.device(device) if not isinstance(device, int) else t.device("cuda", device)
        if dev.type != "cuda":
            raise ValueError("machin_b200.WeightTree lives in GPU memory; device must be a CUDA device")
        self.device = t.device("cuda", dev.index if dev.index is not None else 0)
        _lib.require_device(self.device.index)
        self.size = int(size)
        # geometry exactly as the reference computes it (prioritized_buffer.py:45-51)
        self.depth = int(np.ceil(np.log2(self.size))) + 1
        self.sizes = np.power(2, np.arange(self.depth - 1, -1, -1))
        self.offsets = np.concatenate(([0], np.cumsum(self.sizes)))
        handle = C.c_void_p()
        _lib.check(lib.mb_tree_create(self.size, self.device.index, C.byref(handle)))
        self._h = handle
        depth = C.c_int()
        n_nodes = C.c_int64()
        _lib.check(lib.mb_tree_geometry(self._h, None, C.byref(depth), None, C.byref(n_nodes)))
        assert depth.value == self.depth and n_nodes.value == int(self.offsets[-1])
        self.n_nodes = n_nodes.value

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and lib is not None:
            lib.mb_tree_destroy(h)
            self._h = None

    # ------------------------------------------------------------------ helpers
    def _sync(self):
        t.cuda.current_stream(self.device).synchronize()

    def _stream(self):
        return _lib.stream_ptr(t.cuda.current_stream(self.device))

    def _stats(self):
        s, m, d = C.c_double(), C.c_double(), C.c_int()
        _lib.check(lib.mb_tree_stats(self._h, C.byref(s), C.byref(m), C.byref(d), self._stream()))
        # the reference keeps max_leaf in the dtype of the batch that set it (np.float32 after a
        # float32 priority write-back), and normalize(max_leaf) is evaluated in that dtype
        return s.value, (np.float32(m.value) if d.value == _lib.MB_F32 else m.value)

    def check(self):
        """Raise the deferred ``ValueError`` of an asynchronous (CUDA tensor) call, if any."""
        _lib.check(lib.mb_tree_check(self._h, self._stream()))

    @property
    def handle(self):
        return self._h

    @property
    def weights(self) -> np.ndarray:
        """Snapshot (copy) of the flattened float64 node array, reference layout (:52)."""
        self._sync()
        out = np.empty(self.n_nodes, dtype=np.float64)
        _lib.check(lib.mb_tree_read_host(self._h, 0, self.n_nodes, _lib.ptr(out)))
        return out

    @property
    def max_leaf(self) -> float:
        return self._stats()[1]

    def _view(self, address: int, shape, typestr: str) -> t.Tensor:
        class _Iface:
            pass

        obj = _Iface()
        obj.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "version": 3, "data": (address, False)}
        with t.cuda.device(self.device):
            return t.as_tensor(obj, device=self.device)

    def weights_tensor(self) -> t.Tensor:
        """Zero-copy CUDA view of the node array (read-only by convention)."""
        return self._view(lib.mb_tree_weights(self._h), (self.n_nodes,), "<f8")

    def snapshot(self):
        """Copies of (node array, scalar state) for :meth:`restore` (used around CUDA-graph warm-up)."""
        n = C.c_int64()
        addr = lib.mb_tree_state_blob(self._h, C.byref(n))
        return self.weights_tensor().clone(), self._view(addr, (n.value,), "|u1").clone()

    def restore(self, snap):
        n = C.c_int64()
        addr = lib.mb_tree_state_blob(self._h, C.byref(n))
        self.weights_tensor().copy_(snap[0])
        self._view(addr, (n.value,), "|u1").copy_(snap[1])

    # ------------------------------------------------------------------ reference surface
    def get_weight_sum(self) -> float:
        return self._stats()[0]

    def get_leaf_max(self) -> float:
        return self._stats()[1]

    def get_leaf_all_weights(self) -> np.ndarray:
        self._sync()
        out = np.empty(self.size, dtype=np.float64)
        _lib.check(lib.mb_tree_read_host(self._h, 0, self.size, _lib.ptr(out)))
        return out

    def get_leaf_weight(self, index: Union[int, List[int], np.ndarray, t.Tensor]) -> Any:
        if _is_cuda(index):
            idx = index.to(dtype=t.int64).contiguous().view(-1)
            out = t.empty(idx.shape[0], dtype=t.float64, device=self.device)
            _lib.check(lib.mb_tree_leaf_weight(self._h, _lib.ptr(idx), idx.shape[0], _lib.ptr(out), self._stream()))
            return out
        idx = np.ascontiguousarray(np.array(index).reshape(-1), dtype=np.int64)
        out = np.empty(idx.shape[0], dtype=np.float64)
        self._sync()
        _lib.check(lib.mb_tree_leaf_weight_host(self._h, _lib.ptr(idx), idx.shape[0], _lib.ptr(out)))
        return float(out[0]) if idx.shape[0] == 1 else out

    def find_leaf_index(self, weight: Union[float, List[float], np.ndarray, t.Tensor]):
        if _is_cuda(weight):
            w = weight.to(dtype=t.float64).contiguous().view(-1)
            out = t.empty(w.shape[0], dtype=t.int64, device=self.device)
            _lib.check(lib.mb_tree_find(self._h, _lib.ptr(w), w.shape[0], _lib.ptr(out), self._stream()))
            return out
        w = np.ascontiguousarray(np.array(weight).reshape(-1), dtype=np.float64)
        out = np.empty(w.shape[0], dtype=np.int64)
        self._sync()
        _lib.check(lib.mb_tree_find_host(self._h, _lib.ptr(w), w.shape[0], _lib.ptr(out)))
        return int(out[0]) if w.shape[0] == 1 else out

    def update_leaf(self, weight: float, index: int):
        _lib.check(lib.mb_tree_update_leaf(self._h, float(weight), int(index), self._stream()))

    def update_leaf_batch(self, weights, indexes, valid=None):
        if _is_cuda(weights) or _is_cuda(indexes):
            w = weights if _is_cuda(weights) else t.as_tensor(np.array(weights), device=self.device)
            i = indexes if _is_cuda(indexes) else t.as_tensor(np.array(indexes), device=self.device)
            if w.numel() != i.numel():
                raise ValueError("Dimension of weights and indexes doesn't match!")
            if w.numel() == 0:
                return
            w = w.contiguous().view(-1)
            if w.dtype not in (t.float32, t.float64):
                w = w.to(t.float64)
            i = i.to(dtype=t.int64).contiguous().view(-1)
            v = None if valid is None else valid.to(dtype=t.uint8).contiguous().view(-1)
            _lib.check(lib.mb_tree_update_batch(
                self._h, _lib.ptr(i), _lib.ptr(w), _lib.MB_F32 if w.dtype == t.float32 else _lib.MB_F64,
                _lib.ptr(v), i.shape[0], self._stream()))
            return
        if len(weights) != len(indexes):
            raise ValueError("Dimension of weights and indexes doesn't match!")
        if len(weights) == 0:
            return
        w = np.array(weights)
        if w.dtype not in (np.float32, np.float64):
            w = w.astype(np.float64)
        w = np.ascontiguousarray(w.reshape(-1))
        i = np.ascontiguousarray(np.array(indexes).reshape(-1), dtype=np.int64)
        self._sync()
        _lib.check(lib.mb_tree_update_batch_host(
            self._h, _lib.ptr(i), _lib.ptr(w), _lib.MB_F32 if w.dtype == np.float32 else _lib.MB_F64, i.shape[0]))

    def update_all_leaves(self, weights):
        if len(weights) != self.size:
            raise ValueError("Weights size must match tree size!")
        if _is_cuda(weights):
            w = weights.to(dtype=t.float64).contiguous().view(-1)
            _lib.check(lib.mb_tree_build(self._h, _lib.ptr(w), w.shape[0], self._stream()))
            return
        w = np.ascontiguousarray(np.array(weights), dtype=np.float64)
        self._sync()
        _lib.check(lib.mb_tree_build_host(self._h, _lib.ptr(w), w.shape[0]))

    def clear(self):
        _lib.check(lib.mb_tree_clear(self._h, self._stream()))

    def print_weights(self, precision=2):
        fmt = f"{{:.{precision}f}}"
        w = self.weights
        for i in range(self.depth):
            off, n = int(self.offsets[i]), int(self.sizes[i])
            print([fmt.format(x) for x in w[off:off + n]])
