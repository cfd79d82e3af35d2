"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference sum tree.

Follows machin/frame/buffers/prioritized_buffer.py:8-231 (class WeightTree).  Same
flattened layout (leaves first, root last, leaf level padded to a power of two), same
float64 storage, same per-level operation order, so contents are bit-identical.
"""
import numpy as np


class OracleTree:
    def __init__(self, size: int):
        # prioritized_buffer.py:45-52 -- geometry
        self.size = int(size)
        self.max_leaf = 0
        self.depth = int(np.ceil(np.log2(self.size))) + 1
        self.sizes = 2 ** np.arange(self.depth - 1, -1, -1, dtype=np.int64)
        self.offsets = np.concatenate(([0], np.cumsum(self.sizes))).astype(np.int64)
        self.weights = np.zeros(int(self.offsets[-1]), dtype=np.float64)

    # prioritized_buffer.py:54-94
    def get_weight_sum(self):
        return self.weights[-1]

    def get_leaf_max(self):
        return self.max_leaf

    def get_leaf_all_weights(self):
        return self.weights[: self.size]

    def get_leaf_weight(self, index):
        index = np.asarray(index).reshape(-1)
        if np.any(index >= self.size) or np.any(index < 0):
            raise ValueError("Index has elements out of boundary!")
        picked = self.weights[index]
        return float(picked[0]) if index.shape[0] == 1 else picked

    def find_leaf_index(self, weight):
        # prioritized_buffer.py:106-125 -- level-by-level descent, strict '>'
        w = np.array(weight, dtype=np.float64).reshape(-1)
        idx = np.zeros(w.shape[0], dtype=np.int64)
        for level in range(self.depth - 2, -1, -1):
            left = self.weights[self.offsets[level] + 2 * idx]
            go_right = w > left
            idx = 2 * idx + go_right
            w = w - left * go_right
        idx = np.clip(idx, 0, self.size - 1)
        return int(idx[0]) if idx.shape[0] == 1 else idx

    def update_leaf(self, weight, index):
        # prioritized_buffer.py:127-150 -- single path-to-root, value + sibling.
        # (the reference's bound check admits index == size, :135)
        if not 0 <= index <= self.size:
            raise ValueError("Index has elements out of boundary!")
        self.max_leaf = max(weight, self.max_leaf)
        self.weights[index] = weight
        node = int(index)
        for level in range(1, self.depth):
            node //= 2
            lo = self.offsets[level - 1] + 2 * node
            self.weights[self.offsets[level] + node] = self.weights[lo] + self.weights[lo + 1]

    def update_leaf_batch(self, weights, indexes):
        # prioritized_buffer.py:152-186
        if len(weights) != len(indexes):
            raise ValueError("Dimension of weights and indexes doesn't match!")
        if len(weights) == 0:
            return
        weights = np.array(weights)
        indexes = np.array(indexes)
        if np.any(indexes >= self.size) or np.any(indexes < 0):
            raise ValueError("Index has elements out of boundary!")
        self.max_leaf = max(np.max(weights), self.max_leaf)
        self.weights[indexes] = weights  # duplicates: last one wins (NumPy semantics)
        dirty = indexes
        for level in range(1, self.depth):
            dirty = np.unique(dirty // 2)
            lo = self.offsets[level - 1] + 2 * dirty
            self.weights[self.offsets[level] + dirty] = self.weights[lo] + self.weights[lo + 1]

    def update_all_leaves(self, weights):
        # prioritized_buffer.py:188-199,216-231
        if len(weights) != self.size:
            raise ValueError("Weights size must match tree size!")
        self.weights[: self.size] = np.array(weights)
        self.max_leaf = np.max(self.weights[: self.size])
        for level in range(self.depth - 1):
            lo, n = self.offsets[level], self.sizes[level]
            self.weights[lo + n : lo + n + n // 2] = self.weights[lo : lo + n].reshape(-1, 2).sum(axis=1)
