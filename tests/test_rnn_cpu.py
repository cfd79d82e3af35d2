"""CPU: sequence ("RNN") buffers -- the oracle restatement (oracle/rnn.py) and the product's host logic
(machin_b200/frame/buffers/rnn_buffers.py) against tests/golden/ref_rnn.npz, which the unmodified reference
produced (oracle/make_golden.py gen_rnn; rnn_buffers.py:78-182, :343-413)."""
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import rnn as orn

G = np.load(os.path.join(GOLDEN, "ref_rnn.npz"))


def _pad(seq, width):
    return np.array(list(seq) + [-1] * (width - len(seq)), dtype=np.int64)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_window_selection_matches_reference(tag):
    cap, L, lens = int(G[f"u{tag}_cap"]), int(G[f"u{tag}_L"]), G[f"u{tag}_lens"]
    ring, counter = orn.OracleRing(cap), 0
    for k, n in enumerate(lens):
        ring.store(list(range(counter, counter + int(n))))
        counter += int(n)
        random.seed(5)
        bs, slots = orn.windows_random_unique(ring, L, 4)
        assert np.array_equal(_pad([ring.payload[s] for s in slots], 4 * L), G[f"u{tag}_unique"][k])
        random.seed(6)
        bs, slots = orn.windows_random(ring, L, 4)
        assert np.array_equal(_pad([ring.payload[s] for s in slots], 4 * L), G[f"u{tag}_random"][k])
        bs, slots = orn.windows_all(ring, L)
        assert bs == G[f"u{tag}_all_count"][k]
        assert np.array_equal([ring.payload[s] for s in slots], G[f"u{tag}_all{k}"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_host_window_selection_matches_reference(tag):
    """The product's own selection functions over its EpisodeBook (no device involved)."""
    from machin_b200.frame.buffers import EpisodeBook
    from machin_b200.frame.buffers import rnn_buffers as rb
    cap, L, lens = int(G[f"u{tag}_cap"]), int(G[f"u{tag}_L"]), G[f"u{tag}_lens"]
    book, counter = EpisodeBook(cap), 0
    label = np.full(cap, -1, dtype=np.int64)          # slot -> global step counter (the "index" attribute)
    for k, n in enumerate(lens):
        _, positions = book.assign(int(n))
        label[positions] = np.arange(counter, counter + int(n))
        book.record(positions)
        counter += int(n)
        random.seed(5)
        bs, slots = rb.rnn_sample_random_unique(book.handles, L, 4)
        assert len(slots) == bs * L
        assert np.array_equal(_pad(label[slots], 4 * L), G[f"u{tag}_unique"][k])
        random.seed(6)
        bs, slots = rb.rnn_sample_random(book.handles, L, 4)
        assert np.array_equal(_pad(label[slots], 4 * L), G[f"u{tag}_random"][k])
        bs, slots = rb.rnn_sample_all(book.handles, L)
        assert bs == G[f"u{tag}_all_count"][k]
        assert np.array_equal(label[slots] if slots else [], G[f"u{tag}_all{k}"])


def test_window_reshape_matches_reference():
    import torch as t
    from machin_b200.frame.buffers import rnn_buffers as rb
    ref = G["ua_batch_state"]                          # [B, L, 2]
    flat = t.from_numpy(ref.reshape(-1, ref.shape[-1]).copy())
    assert np.array_equal(rb._chunk(flat, ref.shape[1]).numpy(), ref)
    assert G["ua_batch_reward"].shape == (ref.shape[0], ref.shape[1], 1)
    assert rb._chunk(list(range(6)), 3) == [[0, 1, 2], [3, 4, 5]]


def _replay_prioritized(tag, store, sample, update, tree_weights, skip_draw):
    """Drive one implementation through the fixture's schedule and compare after every step."""
    cap, L, lens = int(G[f"p{tag}_cap"]), int(G[f"p{tag}_L"]), G[f"p{tag}_lens"]
    pri, o, counter = G[f"p{tag}_pri"], 0, 0
    for k, n in enumerate(lens):
        n = int(n)
        p = pri[o:o + n].copy()
        store(list(range(counter, counter + n)), None if p[0] < 0 else p)
        o, counter = o + n, counter + n
        assert np.array_equal(tree_weights(), G[f"p{tag}_trees"][k]), (tag, k)
        if k == 3:
            skip_draw()                                   # the fixture drew here (result given): beta advanced once
            update(G[f"p{tag}_mid_newp"], G[f"p{tag}_mid_index"])
            assert np.array_equal(tree_weights(), G[f"p{tag}_mid_tree"])
    for step in range(3):
        index, isw, seq = sample(G[f"p{tag}_s{step}_uniform"])
        assert np.array_equal(index, G[f"p{tag}_s{step}_index"])
        np.testing.assert_allclose(isw, G[f"p{tag}_s{step}_isw"], rtol=1e-12)
        assert np.array_equal(np.asarray(seq).reshape(len(index), L), G[f"p{tag}_s{step}_seq"].reshape(len(index), L))
        update(G[f"p{tag}_s{step}_newp"], index)
        assert np.array_equal(tree_weights(), G[f"p{tag}_s{step}_tree_after"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_prioritized_sequences_match_reference(tag):
    cap, L = int(G[f"p{tag}_cap"]), int(G[f"p{tag}_L"])
    buf = orn.OracleRNNPrioritized(cap, L)

    def skip_draw():
        buf.beta = np.min([1.0, buf.beta + buf.beta_inc])
    _replay_prioritized(tag, buf.store, lambda u: buf.sample(len(u), uniform=u), buf.update_priority,
                        lambda: buf.tree.weights, skip_draw)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_host_episode_priorities_match_reference(tag):
    """rnn_episode_priorities (product host logic) fed to the oracle tree reproduces the reference's leaves."""
    from machin_b200.frame.buffers import rnn_buffers as rb
    from oracle.per import normalize_priority
    cap, L = int(G[f"p{tag}_cap"]), int(G[f"p{tag}_L"])
    o = orn.OracleRNNPrioritized(cap, L)

    def store(items, p):
        slots = o.ring.store(items)
        default = normalize_priority(o.tree.get_leaf_max()) if p is None else None
        o.tree.update_leaf_batch(rb.rnn_episode_priorities(len(items), L, p, default, normalize_priority), slots)

    def skip_draw():
        o.beta = np.min([1.0, o.beta + o.beta_inc])
    _replay_prioritized(tag, store, lambda u: o.sample(len(u), uniform=u), o.update_priority, lambda: o.tree.weights,
                        skip_draw)


def test_vectorised_episode_sampling_matches_list_sampling():
    """IMPALABuffer.sample_method_episode builds the slot list of the drawn episodes with array arithmetic (an
    episode's slots are consecutive modulo the capacity, the live episodes are a contiguous range of numbers): same
    draws from Python's random stream and the same slots as `random.choices(list(keys))` + per-episode lists
    (impala.py:52-66), through wrap-around and evictions."""
    from machin_b200.frame.buffers import EpisodeBook
    rng = np.random.RandomState(0)
    book = EpisodeBook(5000)
    for step in range(1500):
        n = int(rng.randint(1, 40))
        _, pos = book.assign(n)
        book.record(pos)
        if step % 40:
            continue
        handles = book.handles
        random.seed(step)
        episodes = random.choices(list(handles.keys()), k=min(len(handles), 64))
        want = np.concatenate([np.asarray(handles[e], dtype=np.int64) for e in episodes])
        after = random.random()
        random.seed(step)
        live = book.live_episode_range()
        assert live is not None and live[1] - live[0] == len(handles)
        picks = random.choices(range(live[1] - live[0]), k=min(len(handles), 64))
        got, lens = book.episode_slots(live[0] + np.asarray(picks, dtype=np.int64))
        assert random.random() == after                            # the stream was consumed identically
        assert np.array_equal(got, want) and lens.tolist() == [len(handles[e]) for e in episodes]
    book.clear()
    assert book.live_episode_range() is None


def test_window_index_wraps_around_the_ring():
    """RNNPrioritizedBuffer._window_index: start slot -> sample_length consecutive slots modulo the capacity (an
    episode that wrapped around the end of the ring continues at slot 0)."""
    import types
    import torch as t
    from machin_b200.frame.buffers.rnn_buffers import RNNPrioritizedBuffer
    buf = RNNPrioritizedBuffer.__new__(RNNPrioritizedBuffer)       # no device needed for the index arithmetic
    buf.sample_length = 4
    buf.storage = types.SimpleNamespace(max_size=16)
    got = buf._window_index(t.tensor([0, 5, 13, 15]))
    assert got.tolist() == [0, 1, 2, 3, 5, 6, 7, 8, 13, 14, 15, 0, 15, 0, 1, 2]
