"""TEST INFRASTRUCTURE ONLY -- restatement of the reference PER draw / priority math.

Follows machin/frame/buffers/prioritized_buffer.py:397-434 (sample_index_and_weight,
_normalize_priority) and prioritized_buffer_d.py:204-303 (shard apportioning).
"""
import numpy as np


def normalize_priority(priority, epsilon=1e-2, alpha=0.6):
    # prioritized_buffer.py:430-434 -- dtype follows the input (float32 from torch)
    return (np.abs(priority) + epsilon) ** alpha


def stratified_queries(uniform, weight_sum, batch_size):
    # prioritized_buffer.py:409-416 -- three separately rounded float64 ops, then clip
    seg = weight_sum / batch_size
    q = uniform * seg
    q += np.arange(batch_size, dtype=np.float64) * seg
    return np.clip(q, 0, max(weight_sum - 1e-6, 0))


def sample_index_and_weight(tree, batch_size, n_stored, beta, all_weight_sum=None,
                            uniform=None):
    """Returns (index i64[B], is_weight f64[B]).  ``uniform`` defaults to the global
    NumPy MT19937 stream exactly like the reference (prioritized_buffer.py:411)."""
    if uniform is None:
        uniform = np.random.uniform(size=batch_size)
    q = stratified_queries(uniform, tree.get_weight_sum(), batch_size)
    index = tree.find_leaf_index(q)
    priority = tree.get_leaf_weight(index)
    total = all_weight_sum or tree.get_weight_sum()  # :421 -- 0.0 falls back to local
    prob = priority / total
    isw = np.power(n_stored * prob, -beta)
    isw = isw / np.max(isw)
    return np.asarray(index).reshape(-1), np.asarray(isw).reshape(-1)


def shard_sample_sizes(shard_sums, batch_size):
    # prioritized_buffer_d.py:229-234 -- (+1e-6, ceil) apportioning
    total = sum(shard_sums) + 1e-6
    return [int(s) for s in np.ceil(np.array(shard_sums) * batch_size / total)], total
