"""GPU: the cases of the reference's own Buffer test (test/frame/buffers/test_buffer.py:8-340) run against the
device-resident Buffer: store_episode validation and messages, size / clear, and every combination of
batch size x concatenate x device x sample method x attribute selection of sample_batch, with the reference's
expectations on the FORM of the result (dict of tensors per major attribute, [B, 1] tensors or Python lists for
scalar attributes, lists for custom attributes, "*" collecting the remaining custom attributes)."""
import pytest
import torch as t



pytestmark = [pytest.mark.gpu]

BUFFER_SIZE, SAMPLE_BUFFER_SIZE = 5, 10
REQUIRED = ("state", "action", "next_state", "reward", "terminal")


def _tr():
    return {"state": {"state_1": t.zeros([1, 2])}, "action": {"action_1": t.zeros([1, 3])},
            "next_state": {"next_state_1": t.zeros([1, 2])}, "reward": 1, "terminal": True, "some_custom_attr": None}


def test_store_episode_cases():
    from machin_b200.frame.buffers import Buffer
    from machin_b200.frame.transition import Transition
    Buffer(BUFFER_SIZE).store_episode([_tr(), Transition(**_tr())], REQUIRED)
    with pytest.raises(ValueError, match="Episode must be non-empty"):
        Buffer(BUFFER_SIZE).store_episode([], REQUIRED)
    with pytest.raises(ValueError, match="Transition object must be a dict"):
        Buffer(BUFFER_SIZE).store_episode([[]], REQUIRED)
    with pytest.raises(ValueError, match="Transition object missing attributes"):
        Buffer(BUFFER_SIZE).store_episode([_tr()], REQUIRED + ("some_special_attr",))


def test_size_and_clear():
    from machin_b200.frame.buffers import Buffer
    buffer = Buffer(BUFFER_SIZE)
    buffer.store_episode([_tr()])
    assert buffer.size() == 1
    buffer.clear()
    assert buffer.size() == 0


@pytest.fixture(scope="module")
def const_buffer():
    from machin_b200.frame.buffers import Buffer
    episode = [{"state": {"state_1": t.zeros([1, 2])}, "action": {"action_1": t.zeros([1, 3])},
                "next_state": {"next_state_1": t.zeros([1, 4])}, "reward": 10.0, "terminal": True, "data_index": i,
                "not_concatenable": (i, "some_str")} for i in range(SAMPLE_BUFFER_SIZE)]
    buffer = Buffer(buffer_size=SAMPLE_BUFFER_SIZE, buffer_device="cuda:0")
    buffer.store_episode(episode)
    return buffer


def tensor_equal(a, b):
    return a.shape == b.shape and bool(t.all(a == b.to(a.device)))


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
@pytest.mark.parametrize("dev", [None, "cpu"])
@pytest.mark.parametrize("sample_method", ["random", "random_unique", "all", "some_invalid_method"])
@pytest.mark.parametrize("sample_attrs,concat_attrs,should_be_attrs", CASES)
def test_sample_batch(const_buffer, batch_size, concat, dev, sample_method, sample_attrs, concat_attrs, should_be_attrs):
    sample_not_empty = batch_size != 0 or sample_method == "all"
    will_concat_custom = isinstance(concat_attrs, list) and concat and sample_not_empty
    kw = dict(concatenate=concat, device=dev, sample_method=sample_method, sample_attrs=sample_attrs,
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
