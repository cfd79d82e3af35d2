"""CPU: host-side logic of the path -- Transition validation (mirrors the cases of the reference's
test/frame/test_transition.py), row layout properties, shard apportioning (prioritized_buffer_d.py:229-234),
safe_call argument routing (algorithms/utils.py:52-161)."""
import pytest
import torch as t
from hypothesis import given, settings
from hypothesis import strategies as st

from machin_b200.frame.buffers.device_ring import Field, layout
from machin_b200.frame.transition import Transition
from machin_b200.parallel.exchange import shard_sample_sizes
from oracle import per as oper


def _tr(**kw):
    base = dict(state={"s": t.zeros(1, 3)}, action={"a": t.zeros(1, 2)}, next_state={"s": t.zeros(1, 3)},
                reward=1.0, terminal=False)
    base.update(kw)
    return base


def test_transition_surface():
    tr = Transition(**_tr(extra=5, another="x"))
    assert tr.major_attr == ["state", "action", "next_state"] and tr.sub_attr == ["reward", "terminal"]
    assert tr.custom_attr == ["another", "extra"] and len(tr) == 7
    assert tr.has_keys(["state", "extra"]) and not tr.has_keys(["nope"])
    assert tr["reward"] == 1.0 and tr.extra == 5
    assert [k for k, _ in tr.items()] == tr.keys()
    tr["reward"] = 2.0
    assert tr["reward"] == 2.0
    with pytest.raises(RuntimeError, match="dynamically set"):
        tr["new_key"] = 1
    assert tr.to("cpu") is tr
    grad = t.zeros(1, 3, requires_grad=True) * 2
    assert not Transition(**_tr(state={"s": grad}))["state"]["s"].requires_grad     # detached, like :150-169
    assert Transition(**_tr(reward=t.ones(1, 1), terminal=t.zeros(1, 1, dtype=t.bool)))["reward"].shape == (1, 1)


@pytest.mark.parametrize("bad,match", [
    (dict(state={"s": t.zeros(2, 3)}, action={"a": t.zeros(2, 2)}, next_state={"s": t.zeros(2, 3)}), "must be 1|scalar, but batch size"),
    (dict(state={"s": t.zeros(1, 3)}, action={"a": t.zeros(2, 2)}), "invalid batch size"),
    (dict(state={"s": t.tensor(3.0)}), "invalid tensor"),          # 0-dim tensor (a bare float has no .detach in the reference either)
    (dict(reward=[1.0]), "requires scalar or tensor"),
    (dict(reward=t.ones(2, 1)), "invalid batch size"),
])
def test_transition_validation(bad, match):
    with pytest.raises(ValueError, match=match):
        Transition(**_tr(**bad))


def test_transition_terminal_must_be_bool():
    with pytest.raises(AssertionError):
        Transition(**_tr(terminal=1))


@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(st.sampled_from([t.float32, t.float64, t.int64, t.bool, t.uint8, t.float16]),
                          st.integers(1, 400)), min_size=1, max_size=12))
def test_row_layout_properties(spec):
    fields = [Field(f"f{i}", "a", None, dt, (n,), "sub") for i, (dt, n) in enumerate(spec)]
    pitch = layout(fields)
    assert pitch % 16 == 0 and pitch >= 16
    spans = sorted((f.offset, f.offset + f.nbytes) for f in fields)
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= pitch
    for f in fields:
        esz = t.empty((), dtype=f.dtype).element_size()
        assert f.offset % (16 if f.nbytes >= 16 else esz) == 0
    assert pitch - sum(f.nbytes for f in fields) < 16 * (len(fields) + 1)       # padding stays small


@settings(max_examples=300, deadline=None)
@given(st.lists(st.floats(0, 1e9, allow_nan=False), min_size=1, max_size=16), st.integers(1, 4096))
def test_shard_apportioning_matches_oracle_and_covers_batch(sums, batch):
    ssize, total = shard_sample_sizes(sums, batch)
    assert (ssize, total) == tuple(oper.shard_sample_sizes(sums, batch))
    assert total == sum(sums) + 1e-6
    assert all(0 <= s <= batch + 1 for s in ssize)
    if sum(sums) > 1e-3:
        assert batch <= sum(ssize) <= batch + len(sums)        # ceil: at most one extra row per shard
    assert all(s == 0 for s, w in zip(ssize, sums) if w == 0)


def test_safe_call_routes_named_arguments():
    from machin_b200.frame.algorithms.utils import safe_call

    class Critic(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = t.nn.Linear(5, 1)

        def forward(self, state, action, scale=1.0):
            return self.fc(t.cat([state, action], dim=1)) * scale

    net = Critic()
    out = safe_call(net, {"state": t.zeros(4, 3), "ignored": 1}, {"action": t.zeros(4, 2)})
    assert isinstance(out, tuple) and out[0].shape == (4, 1)
    with pytest.raises(RuntimeError, match="missing required arguments"):
        safe_call(net, {"state": t.zeros(4, 3)})


def test_framework_base_registry_save_and_load(tmp_path):
    """TorchFramework surface (machin/frame/algorithms/base.py:11-184): model registry, replaceable backward function,
    save as <name>_<version>.pt, load a given version or the newest one, network_map renames."""
    from machin_b200.frame.algorithms.base import TorchFramework

    class Algo(TorchFramework):
        _is_top = ["net", "net_target"]
        _is_restorable = ["net_target"]

        def __init__(self):
            self.net, self.net_target = t.nn.Linear(3, 2), t.nn.Linear(3, 2)
            super().__init__()

    a = Algo()
    assert Algo.get_top_model_names() == ["net", "net_target"] and Algo.get_restorable_model_names() == ["net_target"]
    assert a.top_models == [a.net, a.net_target] and a.restorable_models == [a.net_target]
    assert not Algo.is_distributed() and a.backward_function is t.autograd.backward
    with pytest.raises(NotImplementedError):
        a.optimizers
    with pytest.raises(AssertionError, match="Backward function must be callable."):
        a.set_backward_function(3)
    marker = lambda *args, **kw: None
    a.set_backward_function(marker)
    assert a.backward_function is marker
    a.save(str(tmp_path), version=2)
    first = a.net_target.weight.detach().clone()
    with t.no_grad():
        a.net_target.weight.add_(1.0)
    a.save(str(tmp_path), version=10)
    a.save(str(tmp_path), network_map={"net_target": "renamed"}, version=1)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["net_target_10.pt", "net_target_2.pt", "renamed_1.pt"]
    b = Algo()
    b.load(str(tmp_path), version=2)
    assert t.equal(b.net_target.weight, first)
    b.load(str(tmp_path))                                        # newest: 10, not the lexicographically larger "2"
    assert t.equal(b.net_target.weight, first + 1.0)
    with pytest.raises(RuntimeError, match="Cannot find a valid version"):   # the reference's message (utils/prepare.py:109)
        b.load(str(tmp_path), network_map={"net_target": "missing"})


def test_load_picks_newest_version_common_to_all_models(tmp_path):
    """utils/prepare.py:52-117: after a partial save (actor_5 present, critic_5 missing) the newest COMMON version is
    loaded -- never a mixed pair -- and a missing directory / no common version raise RuntimeError."""
    import torch as t
    from machin_b200.frame.algorithms.base import TorchFramework

    class Two(TorchFramework):
        _is_top = ["actor", "critic"]
        _is_restorable = ["actor", "critic"]

        def __init__(self):
            self.actor, self.critic = t.nn.Linear(1, 1, bias=False), t.nn.Linear(1, 1, bias=False)
            super().__init__()

    def fill(algo, a, c):
        with t.no_grad():
            algo.actor.weight.fill_(a)
            algo.critic.weight.fill_(c)

    algo = Two()
    fill(algo, 3.0, 30.0)
    algo.save(str(tmp_path), version=3)
    fill(algo, 5.0, 50.0)
    algo.save(str(tmp_path), version=5)
    (tmp_path / "critic_5.pt").unlink()                      # partial save
    fill(algo, 0.0, 0.0)
    algo.load(str(tmp_path))                                 # newest common = 3
    assert algo.actor.weight.item() == 3.0 and algo.critic.weight.item() == 30.0
    fill(algo, 0.0, 0.0)
    algo.load(str(tmp_path), version=5)                      # not available for every model: falls back to the common one
    assert algo.actor.weight.item() == 3.0 and algo.critic.weight.item() == 30.0
    (tmp_path / "critic_3.pt").unlink()
    with pytest.raises(RuntimeError, match="valid version"):
        algo.load(str(tmp_path))
    with pytest.raises(RuntimeError, match="doesn't exist"):
        algo.load(str(tmp_path / "nope"))


def test_step_options_and_capture_stream(monkeypatch):
    """MB_STEP_OFF switches the stream-level options of the captured update off by name; capture_graph hands the warm-up
    stream to torch.cuda.graph unless `capture_stream` is off."""
    import torch as t
    from machin_b200.frame.algorithms import utils
    seen = []
    monkeypatch.setattr(t.cuda, "graph", lambda graph, **kw: seen.append((graph, kw)) or "ctx")
    monkeypatch.delenv("MB_STEP_OFF", raising=False)
    assert all(utils.step_option(n) for n in ("capture_stream", "last_grad_main", "third_branch"))
    assert utils.capture_graph("g", "warmup-stream", capture_error_mode="thread_local") == "ctx"
    assert seen[-1] == ("g", {"stream": "warmup-stream", "capture_error_mode": "thread_local"})
    monkeypatch.setenv("MB_STEP_OFF", "capture_stream,third_branch")
    assert not utils.step_option("capture_stream") and not utils.step_option("third_branch") and utils.step_option("last_grad_main")
    utils.capture_graph("g", "warmup-stream")
    assert seen[-1] == ("g", {})
