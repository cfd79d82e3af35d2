"""CPU: the drop-in boundary as data.  tests/golden/ref_surface.json lists, for every reference class this
repository mirrors, its public methods / properties and their parameter names (written by
oracle/make_golden.py::gen_surface from the unmodified reference).  Every one of them must exist here, callables must
accept the same parameter names -- apart from the differences listed below, each one a design decision."""
import inspect
import json
import os

import pytest
import torch as t

from conftest import GOLDEN

# parameters that do not carry over, and why
DIFFERENT = {
    # batches are assembled from ring slots on the device, not from a list of Transition objects
    ("*Buffer", "post_process_batch"): {"batch"},
    # a custom storage object cannot back a device ring (the reference's PrioritizedBuffer forbids it as well)
    ("DistributedBuffer", "__init__"): {"storage"},
    # servers sit on a ShardGroup (NCCL broadcast / all-reduce + version), not on an ordered RPC server
    ("PushPullModelServer", "__init__"): {"o_server"},
    ("PushPullGradServer", "__init__"): {"model_name", "secondary_reducers", "o_server"},
}


def _mine():
    import machin_b200.frame.algorithms as A
    import machin_b200.frame.buffers as B
    from machin_b200.parallel.server import PushPullGradServer, PushPullModelServer
    out = {n: getattr(A, n) for n in ("DQN", "DQNPer", "DDPGPer", "DQNApex", "DDPGApex", "IMPALA", "A2C", "A3C", "RAINBOW")}
    out.update({n: getattr(B, n) for n in ("WeightTree", "Buffer", "PrioritizedBuffer", "DistributedBuffer",
                                           "DistributedPrioritizedBuffer", "RNNBuffer", "RNNPrioritizedBuffer")})
    out.update(PushPullModelServer=PushPullModelServer, PushPullGradServer=PushPullGradServer)
    return out


REF = json.load(open(os.path.join(GOLDEN, "ref_surface.json")))


@pytest.mark.parametrize("name", sorted(REF))
def test_public_surface_matches_reference(name):
    cls = _mine()[name]
    problems = []
    for member, params in REF[name].items():
        if not hasattr(cls, member):
            problems.append(f"missing {member}")
            continue
        if params == "property":
            if not isinstance(inspect.getattr_static(cls, member), property):
                problems.append(f"{member} is not a property")
            continue
        if not isinstance(params, list):
            continue
        sig = inspect.signature(getattr(cls, member))
        if any(p.kind == p.VAR_KEYWORD for p in sig.parameters.values()):
            continue                                    # accepts any keyword, like the reference's own **__ catch-alls
        allowed = DIFFERENT.get((name, member), set()) | (DIFFERENT.get(("*Buffer", member), set()) if "Buffer" in name else set())
        lacking = [p for p in params if p not in sig.parameters and p not in ("_", "__", "args", "kwargs") and p not in allowed]
        if lacking:
            problems.append(f"{member} lacks parameters {lacking}")
    assert not problems, problems


def test_hooks_behave_like_the_reference():
    from machin_b200.frame.algorithms import DDPGPer, DQNPer
    reward, nxt = t.tensor([[1.0], [2.0]]), t.tensor([[10.0], [20.0]])
    terminal = t.tensor([[False], [True]])
    for cls in (DQNPer, DDPGPer):
        assert t.equal(cls.reward_function(reward, 0.5, nxt, terminal, None), t.tensor([[6.0], [2.0]]))

    class Probs(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = t.nn.Parameter(t.zeros(1))

        def forward(self, state):
            return t.softmax(state, dim=1) + 0 * self.p

    ddpg = DDPGPer.__new__(DDPGPer)                     # acting only touches the actors
    ddpg.actor = ddpg.actor_target = Probs()
    t.manual_seed(0)
    action, probs = ddpg.act_discrete_with_noise({"state": t.tensor([[0.0, 50.0, 0.0]])})
    assert action.shape == (1, 1) and probs.shape == (1, 3) and abs(float(probs.detach().max()) - 0.95) < 1e-4
    with pytest.raises(ValueError, match="not a probability tensor"):
        class Raw(Probs):
            def forward(self, state):
                return state + 0 * self.p
        ddpg.actor = Raw()
        ddpg.act_discrete_with_noise({"state": t.tensor([[0.3, 0.3, 0.3]])})
