"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's sequence ("RNN") buffers on the CPU.

Follows machin/frame/buffers/rnn_buffers.py: window selection :78-167, window reshape :169-182,
prioritized store :343-380 and draw :382-413, over the ring / episode bookkeeping of buffer.py:72-109 and
storage.py:97-114.  Pinned against tests/golden/ref_rnn.npz (produced by the unmodified reference through
oracle/make_golden.py).  Never imported by machin_b200/.
"""
import random

import numpy as np

from .per import normalize_priority, sample_index_and_weight
from .sumtree import OracleTree


class OracleRing:
    """Slots of a ring of ``capacity`` + which episode owns which slots (whole-episode eviction)."""

    def __init__(self, capacity):
        self.capacity, self.count, self.cursor, self.episodes = capacity, 0, 0, 0
        self.handles = {}                        # episode -> list of slots, insertion ordered
        self.owner = {}                          # slot -> episode
        self.payload = [None] * capacity

    def store(self, items):
        n = len(items)
        start = self.count if self.count < self.capacity else self.cursor        # storage.py:97-114
        slots = [(start + i) % self.capacity for i in range(n)]
        self.count = min(self.capacity, self.count + n)
        self.cursor = (slots[-1] + 1) % self.capacity
        for s in slots:                                                           # buffer.py:93-105
            old = self.owner.get(s)
            if old is not None and old in self.handles:
                for o in self.handles.pop(old):
                    self.owner.pop(o, None)
        ep = self.episodes
        self.episodes += 1
        self.handles[ep] = slots
        for s, it in zip(slots, items):
            self.owner[s] = ep
            self.payload[s] = it
        return slots


def windows_random_unique(ring, length, batch_size):
    valid = [ep for ep, h in ring.handles.items() if len(h) >= length]            # :88-92
    k = min(len(valid), batch_size)
    chosen = random.sample(valid, k=k)                                            # :94
    starts = [random.randint(0, len(ring.handles[ep]) - length) for ep in chosen]  # :95-100
    return k, [s for ep, p in zip(chosen, starts) for s in ring.handles[ep][p:p + length]]


def windows_random(ring, length, batch_size):
    valid = [ep for ep, h in ring.handles.items() if len(h) >= length]
    k = min(len(valid), batch_size)
    chosen = random.choices(valid, k=k)                                           # :126
    starts = [random.randint(0, len(ring.handles[ep]) - length) for ep in chosen]
    return k, [s for ep, p in zip(chosen, starts) for s in ring.handles[ep][p:p + length]]


def windows_all(ring, length):
    count, out = 0, []
    for ep, h in ring.handles.items():                                            # :151-165
        if len(h) < length:
            continue
        for p in range(len(h) - length + 1):
            out += h[p:p + length]
            count += 1
    return count, out


class OracleRNNPrioritized:
    def __init__(self, capacity, length, epsilon=1e-2, alpha=0.6, beta=0.4, beta_inc=1e-3):
        self.ring, self.tree, self.length = OracleRing(capacity), OracleTree(capacity), length
        self.eps, self.alpha, self.beta, self.beta_inc = epsilon, alpha, beta, beta_inc

    def norm(self, p):
        return normalize_priority(p, self.eps, self.alpha)

    def store(self, items, priorities=None):
        slots, n, L = self.ring.store(items), len(items), self.length
        if priorities is None:                                                    # :362-368
            p = self.norm(self.tree.get_leaf_max())
            leaf = [p if i + L <= n else 0 for i in range(n)]
        elif n < L:                                                               # :370-371
            leaf = np.zeros(n)
        else:                                                                     # :373-374
            leaf = self.norm(np.asarray(priorities))
            leaf[n - L + 1:] = 0
        self.tree.update_leaf_batch(leaf, slots)

    def sample(self, batch_size, uniform=None):
        index, isw = sample_index_and_weight(self.tree, batch_size, self.ring.count, self.beta, uniform=uniform)
        self.beta = np.min([1.0, self.beta + self.beta_inc])
        seq = [self.ring.payload[i] for idx in index for i in range(idx, idx + self.length)]   # :391-395
        return index, isw, seq

    def update_priority(self, priorities, index):
        self.tree.update_leaf_batch(self.norm(priorities), index)
