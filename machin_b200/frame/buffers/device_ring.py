"""Device-resident replay ring: packed transition rows in HBM.

Replaces TransitionStorageBasic (machin/frame/buffers/storage.py:66-123), which keeps one
deep-copied Python ``Transition`` per slot.  Here a slot is one packed row of ``pitch`` bytes
in a single ``uint8[capacity, pitch]`` CUDA tensor; a *field* is one tensor of the transition
(``state["k"]``, ``action["k"]``, ``reward``, ``terminal``, numeric custom attributes).
Rows are written by ``mb_rows_scatter`` and sampled by ``mb_rows_gather`` (csrc/rows.cu).

Slot assignment follows storage.py:97-114: append until full, then overwrite at the ring
cursor; an episode longer than the ring raises ``ValueError("Failed to store episode.")``.
"""
from collections import OrderedDict
from typing import Dict, List, Optional

import torch as t

from ... import _lib
from ..._lib import MbField, lib


class Field:
    __slots__ = ("name", "attr", "key", "dtype", "shape", "nbytes", "offset", "kind", "scalar")

    def __init__(self, name, attr, key, dtype, shape, kind, scalar=False):
        self.name, self.attr, self.key, self.dtype, self.shape, self.kind = name, attr, key, dtype, tuple(shape), kind
        self.scalar = scalar     # stored from a Python scalar (reward=1.0, terminal=True): unconcatenated samples give those back
        n = 1
        for s in self.shape:
            n *= int(s)
        self.nbytes = n * t.empty((), dtype=dtype).element_size()
        self.offset = 0

    def __repr__(self):
        return f"Field({self.name}, {self.dtype}, {self.shape}, off={self.offset}, {self.nbytes}B)"


def layout(fields: List[Field]) -> int:
    """Assign row offsets: wide fields first on 16-byte boundaries, scalars packed behind."""
    order = sorted(range(len(fields)), key=lambda i: (fields[i].nbytes < 16, i))
    off = 0
    for i in order:
        f = fields[i]
        align = 16 if f.nbytes >= 16 else max(1, t.empty((), dtype=f.dtype).element_size())
        off = (off + align - 1) // align * align
        f.offset = off
        off += f.nbytes
    return max(16, (off + 15) // 16 * 16)


class DeviceRing:
    def __init__(self, capacity: int, device: t.device):
        self.max_size = int(capacity)
        self.device = device
        self.fields: "OrderedDict[str, Field]" = OrderedDict()
        self.pitch = 0
        self.data: Optional[t.Tensor] = None
        self.count = 0   # len(storage)
        self.index = 0   # ring cursor (storage.py:80)

    def __len__(self):
        return self.count

    def clear(self):
        self.count = 0   # storage.clear() empties data; the cursor restarts with the next append

    # ------------------------------------------------------------------ schema
    def set_schema(self, fields: List[Field]):
        if self.data is not None:
            raise RuntimeError("ring schema already fixed")
        if len(fields) > _lib.MB_MAX_FIELDS:
            raise ValueError(f"at most {_lib.MB_MAX_FIELDS} tensor fields per transition are supported")
        self.pitch = layout(fields)
        self.fields = OrderedDict((f.name, f) for f in fields)
        self.data = t.zeros((self.max_size, self.pitch), dtype=t.uint8, device=self.device)

    def _field_array(self, names, dense: Dict[str, t.Tensor]):
        arr = (MbField * len(names))()
        for i, n in enumerate(names):
            f = self.fields[n]
            d = dense[n]
            assert d.is_cuda and d.is_contiguous() and d.dtype == f.dtype, (n, d.dtype, f.dtype)
            arr[i] = MbField(f.offset, f.nbytes, d.data_ptr(), f.nbytes)
        return arr

    # ------------------------------------------------------------------ data movement
    def write(self, first_slot: int, n: int, dense: Dict[str, t.Tensor], stream=None):
        names = [k for k in self.fields if k in dense]
        arr = self._field_array(names, dense)
        _lib.check(lib.mb_rows_scatter(_lib.ptr(self.data), self.pitch, self.max_size, first_slot, n,
                                       arr, len(names), _lib.stream_ptr(stream)))

    def gather(self, index: t.Tensor, names=None, stream=None, out: Dict[str, t.Tensor] = None) -> Dict[str, t.Tensor]:
        """``out``: optional pre-allocated dense tensors to fill (static batch buffers)."""
        names = list(self.fields) if names is None else [n for n in names if n in self.fields]
        n = int(index.shape[0])
        if out is None:
            out = {k: t.empty((n,) + self.fields[k].shape, dtype=self.fields[k].dtype, device=self.device)
                   for k in names}
        if n and names:
            arr = self._field_array(names, out)
            _lib.check(lib.mb_rows_gather(_lib.ptr(self.data), self.pitch, self.max_size, _lib.ptr(index), n,
                                          arr, len(names), _lib.stream_ptr(stream)))
        return out

    def unpack(self, rows: t.Tensor, names=None, stream=None, out: Dict[str, t.Tensor] = None) -> Dict[str, t.Tensor]:
        """Split ``n`` packed rows (u8[n, >= pitch], e.g. wire rows received from a peer shard, which carry
        metadata behind the payload) into dense fields.  ``out``: optional pre-allocated dense tensors."""
        names = list(self.fields) if names is None else [n for n in names if n in self.fields]
        n = int(rows.shape[0])
        assert rows.is_contiguous() and rows.shape[1] >= self.pitch
        if out is None:
            out = {k: t.empty((n,) + self.fields[k].shape, dtype=self.fields[k].dtype, device=rows.device) for k in names}
        if n and names:
            arr = self._field_array(names, out)
            _lib.check(lib.mb_rows_gather(_lib.ptr(rows), int(rows.shape[1]), n, None, n, arr, len(names),
                                          _lib.stream_ptr(stream)))
        return out

    def gather_into_wire(self, index: t.Tensor, wire: t.Tensor, stream=None):
        """wire[i, :pitch] = ring[index[i]] for a wide (pitch + metadata) contiguous wire buffer."""
        n = int(index.shape[0])
        if n:
            arr = (MbField * 1)(MbField(0, self.pitch, wire.data_ptr(), int(wire.shape[1])))
            _lib.check(lib.mb_rows_gather(_lib.ptr(self.data), self.pitch, self.max_size, _lib.ptr(index), n, arr, 1,
                                          _lib.stream_ptr(stream)))

    def gather_packed(self, index: t.Tensor, out: Optional[t.Tensor] = None, stream=None) -> t.Tensor:
        n = int(index.shape[0])
        if out is None:
            out = t.empty((n, self.pitch), dtype=t.uint8, device=self.device)
        if n:
            _lib.check(lib.mb_rows_gather_packed(_lib.ptr(self.data), self.pitch, self.max_size, _lib.ptr(index),
                                                 n, _lib.ptr(out), _lib.stream_ptr(stream)))
        return out
