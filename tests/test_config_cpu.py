"""CPU: generate_config / init_from_config of the mirrored algorithms (SURVEY 8b lists them in the boundary).
Default tables against what the unmodified reference produces (tests/golden/ref_configs.json from
oracle/make_golden.py::gen_configs), the name resolvers, and construction from a config over host-memory doubles
(product buffers need a device; tests/test_gpu_zz_config.py builds the real thing)."""
import json
import os

import numpy as np
import pytest
import torch as t

from conftest import GOLDEN

NAMES = ["DQN", "DQNPer", "DDPGPer", "DQNApex", "DDPGApex", "IMPALA", "A2C", "A3C", "RAINBOW"]


def _plain(x):
    """tuples -> lists (JSON has no tuples), recursively."""
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    return x


@pytest.mark.parametrize("name", NAMES)
def test_generate_config_matches_reference(name):
    import machin_b200.frame.algorithms as algos
    ref = json.load(open(os.path.join(GOLDEN, "ref_configs.json")))[name]
    cls = getattr(algos, name)
    assert _plain(cls.generate_config({})) == ref["empty"]
    given = {"frame_config": {"batch_size": 7, "my_own_key": 1}, "other": 2}
    got = cls.generate_config(given)
    assert _plain(got) == ref["prefilled"]                       # the defaults win over what was there (dqn.py:530-535)
    assert given == {"frame_config": {"batch_size": 7, "my_own_key": 1}, "other": 2}          # the input is copied
    assert got["frame_config"]["gradient_max"] == np.inf
    got["frame_config"]["model_kwargs"][0]["x"] = 1             # and the tables are not shared between calls
    assert cls.generate_config({})["frame_config"]["model_kwargs"][0] == {}


class GlobalNet(t.nn.Module):
    def __init__(self, width=3):
        super().__init__()
        self.fc = t.nn.Linear(width, 2)

    def forward(self, state):
        return self.fc(state)


def my_loss(a, b):
    return (a - b).abs()


def test_resolvers():
    from machin_b200.frame.algorithms.config import resolve
    assert resolve(GlobalNet, "model") is GlobalNet
    assert resolve("GlobalNet", "model") is GlobalNet                       # a global of a frame of the call stack
    assert resolve("torch.nn.Linear", "model") is t.nn.Linear               # dotted import path
    assert resolve("Adam", "optimizer") is t.optim.Adam and resolve(t.optim.SGD, "optimizer") is t.optim.SGD
    assert resolve("machin_b200.model.FusedAdam", "optimizer").__name__ == "FusedAdam"
    assert resolve("StepLR", "lr_scheduler") is t.optim.lr_scheduler.StepLR
    assert resolve("MSELoss", "criterion") is t.nn.MSELoss and resolve(my_loss, "criterion") is my_loss
    assert resolve("my_loss", "criterion") is my_loss
    for spec, kind in (("NoSuchNet", "model"), (3, "model"), (GlobalNet, "optimizer"), ("nosuch", "lr_scheduler"), (None, "criterion")):
        with pytest.raises(ValueError, match=f"Invalid {kind}"):
            resolve(spec, kind)


class QNet(t.nn.Module):
    def __init__(self, state_dim, action_num):
        super().__init__()
        self.fc = t.nn.Linear(state_dim, action_num)

    def forward(self, state):
        return self.fc(state)


class Actor(t.nn.Module):
    def __init__(self, state_dim, action_dim):
        super().__init__()
        self.fc = t.nn.Linear(state_dim, action_dim)

    def forward(self, state, action=None):
        return t.tanh(self.fc(state))


class Critic(t.nn.Module):
    def __init__(self, state_dim, action_dim=0):
        super().__init__()
        self.fc = t.nn.Linear(state_dim + action_dim, 1)

    def forward(self, state, action=None):
        return self.fc(state if action is None else t.cat([state, action], dim=1))


@pytest.fixture
def doubles(monkeypatch):
    """Product buffers over host-memory doubles, the library as the signature-checking recorder."""
    from threading import RLock
    import machin_b200.frame.algorithms.apex as apex_mod
    import machin_b200.frame.algorithms.impala as impala_mod
    import machin_b200.frame.buffers as buffers_mod
    import machin_b200.frame.algorithms.a2c as a2c_mod
    import machin_b200.frame.algorithms.dqn_per as dqn_mod
    import machin_b200.frame.algorithms.ddpg_per as ddpg_mod
    from fake_ring import make_cpu_buffer
    from test_learner_dry_run_cpu import _Ring, _Tree, _dry_run_patches
    rec = _dry_run_patches(monkeypatch)
    real_per, real_buf = buffers_mod.PrioritizedBuffer, buffers_mod.Buffer
    real_dper, real_imp = buffers_mod.DistributedPrioritizedBuffer, buffers_mod.IMPALABuffer

    def per(buffer_size=1000, buffer_device=None, **kw):
        buf = make_cpu_buffer(real_per, buffer_size, epsilon=1e-2, alpha=0.6, beta=0.4, curr_beta=0.4,
                              beta_increment_per_sampling=1e-3, wt_tree=_Tree(buffer_size))
        buf.storage, buf.buffer_device = _Ring(buffer_size), t.device("cpu", 0)
        return buf

    def dper(buffer_name, group, buffer_size, requesters=None, **kw):
        buf = make_cpu_buffer(real_dper, buffer_size, epsilon=1e-2, alpha=0.6, beta=0.4, curr_beta=0.4,
                              beta_increment_per_sampling=1e-3, wt_tree=_Tree(buffer_size), buffer_name=buffer_name, group=group,
                              wr_lock=RLock(), _requesters=requesters, _sampler=None,
                              buffer_version_table=t.zeros(buffer_size, dtype=t.int64))
        buf.storage, buf.buffer_device = _Ring(buffer_size), t.device("cpu", 0)
        buf.weight_sum_tensor = lambda: t.tensor([max(buf.wt_tree.get_weight_sum(), 1.0)], dtype=t.float64)
        return buf

    for mod in (dqn_mod, ddpg_mod):
        monkeypatch.setattr(mod, "PrioritizedBuffer", per)
    import machin_b200.frame.algorithms.dqn as uniform_mod
    for mod in (a2c_mod, uniform_mod):
        monkeypatch.setattr(mod, "Buffer", lambda buffer_size=1000, buffer_device=None, **kw: make_cpu_buffer(real_buf, buffer_size))
    monkeypatch.setattr(apex_mod, "DistributedPrioritizedBuffer", dper)
    monkeypatch.setattr(impala_mod, "IMPALABuffer", lambda buffer_name, group, buffer_size, requesters=None: make_cpu_buffer(
        real_imp, buffer_size, buffer_name=buffer_name, group=group, wr_lock=RLock(), requesters=requesters or [0],
        _schema_agreed=False))
    return rec


def _fill(config, **model_kwargs):
    n = len(config["frame_config"]["models"])
    config["frame_config"]["model_kwargs"] = [dict(model_kwargs) for _ in range(n)]
    return config


def test_init_from_config_local_frameworks(doubles):
    from machin_b200.frame.algorithms import A2C, DDPGPer, DQNPer, RAINBOW
    cfg = _fill(DQNPer.generate_config({}), state_dim=4, action_num=3)
    cfg["frame_config"].update(optimizer="SGD", learning_rate=0.05, batch_size=9, lr_scheduler="StepLR",
                               lr_scheduler_args=((5,),), lr_scheduler_kwargs=({"gamma": 0.5},))
    dqn = DQNPer.init_from_config(cfg, model_device="cpu")
    assert isinstance(dqn.qnet, QNet) and isinstance(dqn.qnet_optim, t.optim.SGD) and dqn.batch_size == 9
    assert dqn.qnet_optim.param_groups[0]["lr"] == 0.05 and dqn.criterion.reduction == "none"
    assert isinstance(dqn.lr_schedulers[0], t.optim.lr_scheduler.StepLR) and dqn.lr_schedulers[0].gamma == 0.5
    cfg = _fill(RAINBOW.generate_config({}), state_dim=4, action_num=3)
    rb = RAINBOW.init_from_config(cfg, model_device="cpu")
    assert (rb.v_min, rb.v_max, rb.reward_future_steps) == (-1.0, 1.0, 3) and isinstance(rb.qnet_optim, t.optim.Adam)
    cfg = DDPGPer.generate_config({})
    cfg["frame_config"]["model_kwargs"] = [{"state_dim": 4, "action_dim": 2}] * 4
    ddpg = DDPGPer.init_from_config(cfg, model_device="cpu")
    assert isinstance(ddpg.actor, Actor) and isinstance(ddpg.critic_target, Critic)
    assert ddpg.actor_optim.param_groups[0]["lr"] == 0.0005 and ddpg.critic_optim.param_groups[0]["lr"] == 0.001
    cfg = A2C.generate_config({})
    cfg["frame_config"]["model_kwargs"] = [{"state_dim": 4, "action_dim": 2}, {"state_dim": 4}]
    cfg["frame_config"]["criterion"] = "SmoothL1Loss"            # (whatever the name resolves to is CALLED to build it)
    a2c = A2C.init_from_config(cfg, model_device="cpu")
    assert isinstance(a2c.criterion, t.nn.SmoothL1Loss) and a2c.actor_update_times == 5 and a2c.gae_lambda == 1.0


def test_init_from_config_distributed_frameworks_on_a_group_of_one(doubles):
    """Outside torch.distributed every group is a group of one and the only member is a learner."""
    from machin_b200.frame.algorithms import A3C, DDPGApex, DQNApex, IMPALA
    from machin_b200.parallel.server import PushPullGradServer, PushPullModelServer
    cfg = _fill(DQNApex.generate_config({}), state_dim=4, action_num=3)
    cfg["frame_config"]["criterion_kwargs"] = {"reduction": "sum"}
    apex = DQNApex.init_from_config(cfg, model_device="cpu")
    assert apex.role == "learner" and apex.criterion.reduction == "none" and apex.apex_group.size() == 1
    assert isinstance(apex.qnet_model_server, PushPullModelServer) and apex.learner_group is None
    assert apex.apex_group.get_group_name() == "dqn_apex" and apex.replay_buffer.buffer_name == "buffer"
    cfg = DDPGApex.generate_config({})
    cfg["frame_config"]["model_kwargs"] = [{"state_dim": 4, "action_dim": 2}] * 4
    dd = DDPGApex.init_from_config(cfg, model_device="cpu")
    assert dd.role == "learner" and isinstance(dd.actor_model_server, PushPullModelServer)
    cfg = IMPALA.generate_config({})
    cfg["frame_config"]["model_kwargs"] = [{"state_dim": 4, "action_dim": 2}, {"state_dim": 4}]
    imp = IMPALA.init_from_config(cfg, model_device="cpu")
    assert imp.role == "learner" and imp.batch_size == 5 and imp.impala_group.get_group_name() == "impala"
    cfg = A3C.generate_config({})
    cfg["frame_config"]["model_kwargs"] = [{"state_dim": 4, "action_dim": 2}, {"state_dim": 4}]
    cfg["frame_config"].update(actor_learning_rate=0.01, critic_learning_rate=0.02)
    a3c = A3C.init_from_config(cfg, model_device="cpu")
    servers = (a3c.actor_grad_server, a3c.critic_grad_server)
    assert all(isinstance(s, PushPullGradServer) for s in servers)
    assert isinstance(servers[0].model, Actor) and isinstance(servers[1].model, Critic)       # each manages its own kind
    assert servers[0].optimizer.param_groups[0]["lr"] == 0.01 and servers[1].optimizer.param_groups[0]["lr"] == 0.02
    assert servers[0].model is not a3c.actor


# ------------------------------------------------------------------ SPMD: roles and groups over gloo, world 3
class _Recorded:
    """Stands in for the framework class: records what init_sharded passes to the constructor."""

    def __init__(self, *args, **kwargs):
        self.args, self.kwargs, self.served = args, kwargs, 0

    def serve_update(self):
        self.served += 1

    def update(self):
        return "learned"


def _spmd_roles(rank, world):
    from machin_b200.frame.algorithms import DQNApex
    from machin_b200.frame.algorithms import config as builders
    cfg = DQNApex.generate_config({})
    cfg["frame_config"].update(models=[t.nn.Identity, t.nn.Identity], learner_process_number=2)
    frame = builders.init_sharded(_Recorded, cfg, "cpu", "apex")
    qnet, qtgt, optimizer, criterion, group, servers = frame.args
    learner_group = frame.kwargs["learner_group"]
    out = {"role": frame.role, "update": frame.update(), "served": frame.served, "group": (group.size(), group.index),
           "learners": (learner_group.size(), learner_group.is_member()), "requesters": frame.kwargs.get("requesters"),
           "server_group": servers[0].group.size(), "reduction": criterion.reduction,
           "optimizer": optimizer.__name__, "lr": frame.kwargs["learning_rate"]}
    # the learner group is a working process group of its members
    if learner_group.is_member():
        x = t.ones(1, dtype=t.float64) * (rank + 1)
        learner_group.all_reduce_sum(x)
        out["sum"] = float(x)
    group.barrier()
    return out


def test_init_sharded_roles_over_gloo():
    from test_distributed_cpu import run_multi
    res = run_multi(_spmd_roles, 3)
    assert [r["role"] for r in res] == ["learner", "learner", "sampler"]
    assert [r["update"] for r in res] == ["learned", "learned", None] and [r["served"] for r in res] == [0, 0, 1]
    assert [r["group"] for r in res] == [(3, 0), (3, 1), (3, 2)] and all(r["server_group"] == 3 for r in res)
    assert [r["learners"] for r in res] == [(2, True), (2, True), (2, False)]
    assert all(r["requesters"] == [0, 1] and r["reduction"] == "none" and r["optimizer"] == "Adam" for r in res)
    assert res[0]["sum"] == res[1]["sum"] == 3.0 and "sum" not in res[2]


def test_server_helpers_on_a_group_of_one(doubles):
    """machin.frame.helpers.servers surface: names, one server per model, the primary manages model + optimizer +
    scheduler, per-model learning rates and optimizer keyword arguments."""
    from machin_b200.frame.helpers import grad_server_helper, model_server_helper
    dev = t.device("cpu", 0)
    servers = model_server_helper(3, group_name="weights", device=dev)
    assert [s.model_name for s in servers] == ["model_server_0", "model_server_1", "model_server_2"]
    assert len({id(s.group) for s in servers}) == 1 and servers[0].group.get_group_name() == "weights"
    made = []
    creators = [lambda: made.append("a") or GlobalNet(3), lambda: made.append("b") or GlobalNet(5)]
    grads = grad_server_helper(creators, group_name="grads", optimizer=t.optim.SGD, learning_rate=[0.1, 0.2],
                               optimizer_kwargs=[{"momentum": 0.5}, {}], lr_scheduler=t.optim.lr_scheduler.StepLR,
                               lr_scheduler_args=[(3,), (4,)], device=dev)
    assert made == ["a", "b"] and [g.server_name for g in grads] == ["grad_server_0", "grad_server_1"]
    assert grads[0].model.fc.in_features == 3 and grads[1].model.fc.in_features == 5
    assert grads[0].optimizer.param_groups[0]["lr"] == 0.1 and grads[0].optimizer.param_groups[0]["momentum"] == 0.5
    assert grads[1].optimizer.param_groups[0]["lr"] == 0.2 and grads[1].lr_scheduler.step_size == 4
    same = grad_server_helper([lambda: GlobalNet()], learning_rate=0.01, device=dev)
    assert isinstance(same[0].optimizer, t.optim.Adam) and same[0].optimizer.param_groups[0]["lr"] == 0.01
