"""GPU: sequence ("RNN") buffers over the device ring against fixtures of the unmodified reference
(tests/golden/ref_rnn.npz; rnn_buffers.py:78-182, :343-413).  SURVEY §8(f) rank 4."""
import os
import random

import numpy as np
import pytest
import torch as t

from conftest import GOLDEN

pytestmark = [pytest.mark.gpu]
G = np.load(os.path.join(GOLDEN, "ref_rnn.npz"))


def _episode(n, state_dim, rng, counter):
    """Same generator, same draw order as oracle/make_golden.py:_episode / _indexed_episode."""
    ep = []
    for i in range(n):
        ep.append(dict(
            state={"state": t.from_numpy(rng.randn(1, state_dim).astype(np.float32))},
            action={"action": t.tensor([[int(rng.randint(0, 6))]], dtype=t.long)},
            next_state={"state": t.from_numpy(rng.randn(1, state_dim).astype(np.float32))},
            reward=float(np.float32(rng.randn())),
            terminal=bool(rng.rand() < 0.1),
            index=counter + i,
        ))
    return ep


def _trim(a):
    return a[a >= 0]


@pytest.mark.parametrize("tag", ["a", "b"])
def test_rnn_buffer_windows_match_reference(tag):
    from machin_b200.frame.buffers import RNNBuffer
    cap, L, lens = int(G[f"u{tag}_cap"]), int(G[f"u{tag}_L"]), G[f"u{tag}_lens"]
    buf = RNNBuffer(sample_length=L, buffer_size=cap)
    rng, counter = np.random.RandomState(11), 0
    kw = dict(sample_attrs=["index"], additional_concat_custom_attrs=["index"])
    for k, n in enumerate(lens):
        buf.store_episode(_episode(int(n), 2, rng, counter))
        counter += int(n)
        for seed, method, key in ((5, "random_unique", "unique"), (6, "random", "random")):
            want = _trim(G[f"u{tag}_{key}"][k])
            random.seed(seed)
            bs, res = buf.sample_batch(4, sample_method=method, **kw)
            assert bs == len(want) // L
            if bs:
                assert res[0].shape == (bs, L, 1)
                assert np.array_equal(res[0].cpu().numpy().reshape(-1), want)
            else:
                assert res is None
        bs, res = buf.sample_batch(0, sample_method="all", **kw)
        assert bs == G[f"u{tag}_all_count"][k]
        if bs:
            assert np.array_equal(res[0].cpu().numpy().reshape(-1), G[f"u{tag}_all{k}"])
    random.seed(9)
    bs, (state, reward, idx) = buf.sample_batch(3, sample_method="random_unique", sample_attrs=["state", "reward", "index"],
                                                additional_concat_custom_attrs=["index"])
    assert np.array_equal(state["state"].cpu().numpy(), G[f"u{tag}_batch_state"])
    assert np.array_equal(reward.cpu().numpy(), G[f"u{tag}_batch_reward"])
    assert np.array_equal(idx.cpu().numpy(), G[f"u{tag}_batch_index"])
    # not concatenated: lists of windows (:173-177)
    random.seed(9)
    bs, (idx_list,) = buf.sample_batch(3, concatenate=False, sample_method="random_unique", sample_attrs=["index"])
    assert [list(w) for w in idx_list] == G[f"u{tag}_batch_index"].reshape(bs, L).tolist()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_rnn_prioritized_buffer_matches_reference(tag):
    from machin_b200.frame.buffers import RNNPrioritizedBuffer
    cap, L, lens = int(G[f"p{tag}_cap"]), int(G[f"p{tag}_L"]), G[f"p{tag}_lens"]
    buf = RNNPrioritizedBuffer(sample_length=L, buffer_size=cap)
    pri, states, o = G[f"p{tag}_pri"], G[f"p{tag}_states"], 0
    kw = dict(sample_attrs=["state", "reward", "index"], additional_concat_custom_attrs=["index"])
    for k, n in enumerate(lens):
        n = int(n)
        ep = [dict(state={"state": t.from_numpy(states[o + i:o + i + 1].copy())}, action={"action": t.tensor([[1]])},
                   next_state={"state": t.zeros(1, 3)}, reward=0.5, terminal=False, index=o + i) for i in range(n)]
        p = pri[o:o + n].copy()
        buf.store_episode(ep, priorities=None if p[0] < 0 else p)
        o += n
        assert np.array_equal(buf.wt_tree.weights, G[f"p{tag}_trees"][k]), (tag, k)
        if k == 3:
            np.random.seed(21)
            n_s, _, index, isw = buf.sample_batch(6, sample_attrs=["index"], additional_concat_custom_attrs=["index"])
            assert np.array_equal(index, G[f"p{tag}_mid_index"])
            buf.update_priority(G[f"p{tag}_mid_newp"], index)
            assert np.array_equal(buf.wt_tree.weights, G[f"p{tag}_mid_tree"])
    np.random.seed(7)
    for step in range(3):
        q = f"p{tag}_s{step}_"
        assert buf.curr_beta == float(G[q + "beta"])
        n_s, (state, reward, seq), index, isw = buf.sample_batch(len(G[q + "uniform"]), **kw)
        assert np.array_equal(index, G[q + "index"])                       # bit-exact draw
        np.testing.assert_allclose(isw, G[q + "isw"], rtol=1e-12)
        assert np.array_equal(state["state"].cpu().numpy(), G[q + "state"])
        assert reward.shape == (n_s, L, 1)
        assert np.array_equal(seq.cpu().numpy(), G[q + "seq"])
        buf.update_priority(G[q + "newp"], index)
        assert np.array_equal(buf.wt_tree.weights, G[q + "tree_after"])
    buf.wt_tree.check()


def test_rnn_prioritized_device_path_and_ring_wrap():
    """Tensor-native store + device draw; a window that starts near the end of a wrapped episode continues at
    slot 0 (the reference would raise IndexError at rnn_buffers.py:393)."""
    from machin_b200.frame.buffers import RNNPrioritizedBuffer
    cap, L = 16, 4
    buf = RNNPrioritizedBuffer(sample_length=L, buffer_size=cap)

    def store(n, first):
        v = t.arange(first, first + n, dtype=t.float32).view(n, 1)
        buf.store_episode_tensors({"state": {"s": v}, "action": {"a": t.zeros(n, 1, dtype=t.int64)}, "next_state": {"s": v + 1}},
                                  {"reward": v.clone(), "terminal": t.zeros(n, 1, dtype=t.bool)}, priorities=np.ones(n))
    store(12, 0)          # slots 0..11
    store(8, 100)         # slots 12..15, 0..3 (wraps)
    leaves = buf.wt_tree.get_leaf_all_weights()
    assert np.count_nonzero(leaves[12:16]) == 4 and leaves[0] > 0 and np.all(leaves[1:4] == 0)
    n, fields, index, isw = buf.sample_device(64)
    s = fields["state/s"].cpu().numpy().reshape(64, L)
    idx = index.cpu().numpy()
    assert fields["state/s"].shape == (64, L, 1)
    for i, w in zip(idx, s):
        assert np.all(np.diff(w) == 1), (i, w)            # consecutive steps of ONE episode
    assert np.any(idx >= 13)                              # some windows crossed the end of the ring
