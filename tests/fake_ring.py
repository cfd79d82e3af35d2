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


class FakeTree:
    """TEST DOUBLE for machin_b200.frame.buffers.WeightTree over the oracle tree (host memory)."""

    def __init__(self, size):
        from oracle.sumtree import OracleTree
        self.ora = OracleTree(size)
        self.size = size

    @property
    def weights(self):
        return self.ora.weights

    @property
    def max_leaf(self):
        return self.ora.max_leaf

    def get_leaf_max(self):
        return self.ora.get_leaf_max()

    def get_weight_sum(self):
        return self.ora.get_weight_sum()

    def update_leaf_batch(self, weights, indexes, valid=None):
        assert valid is None
        self.ora.update_leaf_batch(weights, indexes)

    def clear(self):
        self.__init__(self.size)


def make_cpu_prioritized(cls, capacity, epsilon=1e-2, alpha=0.6, beta=0.4, beta_increment_per_sampling=1e-3, **attrs):
    """A product prioritized-buffer class over FakeRing + FakeTree; the draw (the CUDA kernel of the product) is the
    oracle's restatement of prioritized_buffer.py:397-428, everything around it is the product's own Python."""
    import numpy as np
    from oracle import per as oper

    class Cpu(cls):
        def _draw(self, batch_size, all_weight_sum=None, uniform=None, use_device_params=False, out=None):
            idx, isw = oper.sample_index_and_weight(self.wt_tree.ora, batch_size, len(self.storage), self.curr_beta,
                                                    all_weight_sum)
            self.curr_beta = np.min([1.0, self.curr_beta + self.beta_increment_per_sampling])
            return t.from_numpy(np.asarray(idx, dtype=np.int64)), t.from_numpy(np.asarray(isw))

    buf = make_cpu_buffer(Cpu, capacity, epsilon=epsilon, alpha=alpha, beta=beta, curr_beta=beta,
                          beta_increment_per_sampling=beta_increment_per_sampling, wt_tree=FakeTree(capacity), **attrs)
    return buf
