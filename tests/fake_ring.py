"""TEST DOUBLE -- a host-memory stand-in for machin_b200.frame.buffers.DeviceRing, so that the PYTHON logic of the
buffers (schema building, sample methods, attribute selection, window reshaping, custom attributes) can run in the
CPU suite.  Same interface and the same packed-row layout (machin_b200.frame.buffers.device_ring.layout); the row
movement that the CUDA kernels do is plain tensor indexing here.  It is not part of the library and nothing under
machin_b200/ imports it: product buffers refuse to be constructed without a CUDA device."""
from collections import OrderedDict

import torch as t

from machin_b200.frame.buffers import EpisodeBook
from machin_b200.frame.buffers.device_ring import layout


class FakeRing:
    def __init__(self, capacity):
        self.max_size, self.device = int(capacity), t.device("cpu")
        self.fields, self.pitch, self.data, self.count, self.index = OrderedDict(), 0, None, 0, 0

    def __len__(self):
        return self.count

    def clear(self):
        self.count = 0

    def set_schema(self, fields):
        self.pitch = layout(fields)
        self.fields = OrderedDict((f.name, f) for f in fields)
        self.data = t.zeros((self.max_size, self.pitch), dtype=t.uint8)

    def _bytes(self, f, dense, n):
        return dense.contiguous().view(t.uint8).reshape(n, f.nbytes)

    def write(self, first_slot, n, dense, stream=None):
        slots = (first_slot + t.arange(n)) % self.max_size
        for name, f in self.fields.items():
            if name in dense:
                assert dense[name].dtype == f.dtype
                self.data[slots, f.offset:f.offset + f.nbytes] = self._bytes(f, dense[name], n)

    def _split(self, rows, names, out):
        n = rows.shape[0]
        names = list(self.fields) if names is None else [k for k in names if k in self.fields]
        res = {} if out is None else out
        for k in names:
            f = self.fields[k]
            val = rows[:, f.offset:f.offset + f.nbytes].contiguous().view(f.dtype).reshape((n,) + f.shape)
            if out is None:
                res[k] = val
            else:
                out[k].copy_(val)
        return res

    def gather(self, index, names=None, stream=None, out=None):
        return self._split(self.data[index.clamp(0, self.max_size - 1)], names, out)

    def unpack(self, rows, names=None, stream=None, out=None):
        return self._split(rows, names, out)

    def gather_packed(self, index, out=None, stream=None):
        return self.data[index.clamp(0, self.max_size - 1)].clone()


def make_cpu_buffer(cls, capacity, **attrs):
    """An instance of a product buffer class over a FakeRing, built without its CUDA-only constructor."""
    buf = cls.__new__(cls)
    buf.buffer_device = t.device("cpu")
    buf.storage = FakeRing(capacity)
    buf.book = EpisodeBook(capacity)
    buf._major_attr, buf._sub_attr, buf._custom_attr, buf._host_custom = [], [], [], {}
    for k, v in attrs.items():
        setattr(buf, k, v)
    return buf
