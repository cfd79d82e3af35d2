"""CPU: the cases of the reference's own Buffer test (test/frame/buffers/test_buffer.py:8-340) run against the PYTHON
logic of machin_b200's Buffer over a host-memory test double of the device ring (tests/fake_ring.py): store_episode
validation and messages, size / clear, and every combination of batch size x concatenate x device x sample method x
attribute selection of sample_batch with the reference's expectations on the form of the result.  The same cases
run against the real device ring in tests/test_gpu_zz_buffer_ref_cases.py.  Also: RNNBuffer (rnn_buffers.py:78-182)
end to end against the fixture of the unmodified reference."""
import os
import random

import numpy as np
import pytest
import torch as t

from conftest import GOLDEN
from fake_ring import make_cpu_buffer

BUFFER_SIZE, SAMPLE_BUFFER_SIZE = 5, 10
REQUIRED = ("state", "action", "next_state", "reward", "terminal")


def _buffer(cap):
    from machin_b200.frame.buffers import Buffer
    return make_cpu_buffer(Buffer, cap)


def _tr():
    return {"state": {"state_1": t.zeros([1, 2])}, "action": {"action_1": t.zeros([1, 3])},
            "next_state": {"next_state_1": t.zeros([1, 2])}, "reward": 1, "terminal": True, "some_custom_attr": None}


def test_store_episode_size_clear_cases():
    from machin_b200.frame.transition import Transition
    _buffer(BUFFER_SIZE).store_episode([_tr(), Transition(**_tr())], REQUIRED)
    with pytest.raises(ValueError, match="Episode must be non-empty"):
        _buffer(BUFFER_SIZE).store_episode([], REQUIRED)
    with pytest.raises(ValueError, match="Transition object must be a dict"):
        _buffer(BUFFER_SIZE).store_episode([[]], REQUIRED)
    with pytest.raises(ValueError, match="Transition object missing attributes"):
        _buffer(BUFFER_SIZE).store_episode([_tr()], REQUIRED + ("some_special_attr",))
    buffer = _buffer(BUFFER_SIZE)
    buffer.store_episode([_tr()])
    assert buffer.size() == 1
    buffer.clear()
    assert buffer.size() == 0
    with pytest.raises(ValueError, match="Failed to store episode"):          # longer than the ring (storage.py:102-104)
        _buffer(BUFFER_SIZE).store_episode([_tr() for _ in range(BUFFER_SIZE + 1)])


@pytest.fixture(scope="module")
def const_buffer():
    episode = [{"state": {"state_1": t.zeros([1, 2])}, "action": {"action_1": t.zeros([1, 3])},
                "next_state": {"next_state_1": t.zeros([1, 4])}, "reward": 10.0, "terminal": True, "data_index": i,
                "not_concatenable": (i, "some_str")} for i in range(SAMPLE_BUFFER_SIZE)]
    buffer = _buffer(SAMPLE_BUFFER_SIZE)
    buffer.store_episode(episode)
    return buffer


def tensor_equal(a, b):
    return a.shape == b.shape and bool(t.all(a == b))


ALL = ["state", "action", "next_state", "reward", "terminal"]
CASES = [
    (None, None, ALL + ["data_index", "not_concatenable"]),
    (ALL, [], ALL),
    (ALL + ["data_index"], ["data_index"], ALL + ["data_index"]),
    (ALL + ["*"], ["data_index"], ALL + ["*"]),
    (ALL + ["*"], ["not_concatenable"], ALL + ["*"]),
]


@pytest.mark.parametrize("batch_size", [0, SAMPLE_BUFFER_SIZE // 2, SAMPLE_BUFFER_SIZE, SAMPLE_BUFFER_SIZE * 2])
@pytest.mark.parametrize("concat", [True, False])
@pytest.mark.parametrize("sample_method", ["random", "random_unique", "all", "some_invalid_method"])
@pytest.mark.parametrize("sample_attrs,concat_attrs,should_be_attrs", CASES)
def test_sample_batch(const_buffer, batch_size, concat, sample_method, sample_attrs, concat_attrs, should_be_attrs):
    sample_not_empty = batch_size != 0 or sample_method == "all"
    will_concat_custom = isinstance(concat_attrs, list) and concat and sample_not_empty
    kw = dict(concatenate=concat, device="cpu", sample_method=sample_method, sample_attrs=sample_attrs,
              additional_concat_custom_attrs=concat_attrs)
    if sample_method == "some_invalid_method":
        with pytest.raises(RuntimeError, match="Cannot find specified sample method"):
            const_buffer.sample_batch(batch_size, **kw)
        return
    if will_concat_custom and "not_concatenable" in concat_attrs:
        with pytest.raises(ValueError, match="Batch not concatenable"):
            const_buffer.sample_batch(batch_size, **kw)
        return
    bsize, b = const_buffer.sample_batch(batch_size, **kw)
    if bsize == 0:
        assert b is None
        return
    assert bsize == (SAMPLE_BUFFER_SIZE if sample_method == "all" else min(batch_size, SAMPLE_BUFFER_SIZE))
    assert len(b) == len(should_be_attrs)
    widths = {"state": ("state_1", 2), "action": ("action_1", 3), "next_state": ("next_state_1", 4)}
    for data, attr in zip(b, should_be_attrs):
        if attr in widths:
            key, width = widths[attr]
            if concat:
                assert tensor_equal(data[key], t.zeros([bsize, width]))
            else:
                assert isinstance(data[key], list) and len(data[key]) == bsize
                assert tensor_equal(data[key][0], t.zeros([1, width]))
        elif attr == "reward":
            if concat:
                assert tensor_equal(data, t.full([bsize, 1], 10.0))
            else:
                assert isinstance(data, list) and len(data) == bsize and data[0] == 10
        elif attr == "terminal":
            if concat:
                assert tensor_equal(data, t.full([bsize, 1], True, dtype=t.bool))
            else:
                assert isinstance(data, list) and len(data) == bsize and data[0] is True
        elif attr == "data_index":
            if will_concat_custom:
                assert list(data.shape) == [bsize, 1]
            else:
                assert isinstance(data, list) and len(data) == bsize and isinstance(data[0], int)
        elif attr == "not_concatenable":
            assert isinstance(data, list) and len(data) == bsize and isinstance(data[0], tuple)
        elif attr == "*":
            assert isinstance(data, dict) and set(data) == {"data_index", "not_concatenable"} - set(should_be_attrs)


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


@pytest.mark.parametrize("tag", ["a", "b"])
def test_rnn_buffer_end_to_end_vs_reference_fixture(tag):
    from machin_b200.frame.buffers import RNNBuffer
    G = np.load(os.path.join(GOLDEN, "ref_rnn.npz"))
    cap, L, lens = int(G[f"u{tag}_cap"]), int(G[f"u{tag}_L"]), G[f"u{tag}_lens"]
    buf = make_cpu_buffer(RNNBuffer, cap, sample_length=L, sample_dimension=1)
    rng, counter = np.random.RandomState(11), 0
    kw = dict(sample_attrs=["index"], additional_concat_custom_attrs=["index"])
    for k, n in enumerate(lens):
        buf.store_episode(_episode(int(n), 2, rng, counter))
        counter += int(n)
        for seed, method, key in ((5, "random_unique", "unique"), (6, "random", "random")):
            want = G[f"u{tag}_{key}"][k]
            want = want[want >= 0]
            random.seed(seed)
            bs, res = buf.sample_batch(4, sample_method=method, **kw)
            assert bs == len(want) // L
            if bs:
                assert res[0].shape == (bs, L, 1) and np.array_equal(res[0].numpy().reshape(-1), want)
            else:
                assert res is None
        bs, res = buf.sample_batch(0, sample_method="all", **kw)
        assert bs == G[f"u{tag}_all_count"][k]
        if bs:
            assert np.array_equal(res[0].numpy().reshape(-1), G[f"u{tag}_all{k}"])
    random.seed(9)
    bs, (state, reward, idx) = buf.sample_batch(3, sample_method="random_unique", sample_attrs=["state", "reward", "index"],
                                                additional_concat_custom_attrs=["index"])
    assert np.array_equal(state["state"].numpy(), G[f"u{tag}_batch_state"])
    assert np.array_equal(reward.numpy(), G[f"u{tag}_batch_reward"])
    assert np.array_equal(idx.numpy(), G[f"u{tag}_batch_index"])
    random.seed(9)
    bs, (idx_list,) = buf.sample_batch(3, concatenate=False, sample_method="random_unique", sample_attrs=["index"])
    assert [list(w) for w in idx_list] == G[f"u{tag}_batch_index"].reshape(bs, L).tolist()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_rnn_prioritized_buffer_python_logic_vs_reference_fixture(tag):
    """RNNPrioritizedBuffer's own Python (store priorities with the unstartable tail zeroed, default priority and its
    dtype, window gathering and reshaping, beta) over test doubles of ring, tree and draw -- against the fixture of
    the unmodified reference (rnn_buffers.py:343-413).  The same schedule runs on the device in
    tests/test_gpu_zz_rnn.py."""
    from fake_ring import make_cpu_prioritized
    from machin_b200.frame.buffers import RNNPrioritizedBuffer
    G = np.load(os.path.join(GOLDEN, "ref_rnn.npz"))
    cap, L, lens = int(G[f"p{tag}_cap"]), int(G[f"p{tag}_L"]), G[f"p{tag}_lens"]
    buf = make_cpu_prioritized(RNNPrioritizedBuffer, cap, sample_length=L, sample_dimension=1)
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
        assert np.array_equal(index, G[q + "index"])
        np.testing.assert_allclose(isw, G[q + "isw"], rtol=1e-12)
        assert np.array_equal(state["state"].numpy(), G[q + "state"])
        assert reward.shape == (n_s, L, 1)
        assert np.array_equal(seq.numpy(), G[q + "seq"])
        buf.update_priority(G[q + "newp"], index)
        assert np.array_equal(buf.wt_tree.weights, G[q + "tree_after"])
