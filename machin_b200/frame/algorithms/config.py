"""``generate_config`` / ``init_from_config`` for the mirrored algorithms.

Reference surface: the classmethods of dqn.py:503-563, dqn_per.py:163-196, ddpg.py / ddpg_per.py, a2c.py, a3c.py:176-248,
impala.py:408-509, apex.py:152-254 and :432-532, rainbow.py:332-339, and the name resolvers of
algorithms/utils.py:171-312.  A config is a mapping with ``frame`` (class name), ``frame_config`` (constructor
arguments with models / optimizer / criterion / lr_scheduler given as classes or names) and, for the distributed
frameworks, ``batch_num``.  The default tables below reproduce the reference's values (pinned by
tests/golden/ref_configs.json, which the unmodified reference wrote).

Kept as in the reference: ``generate_config`` lets the DEFAULTS win over values already present in ``frame_config``
(``{**config["frame_config"], **default_values}``), and a name is looked up in the globals of every frame of the call
stack, outermost frame winning.

Different by design: processes are SPMD ranks of one torch.distributed job, not named RPC processes -- group
members are ranks (or "all"), learners average gradients with one all-reduce instead of a DistributedDataParallel
wrapper, a non-learner member's ``update`` is its collective ``serve_update`` (the reference disables it because its
shards are served by RPC threads), and replay memory lives in HBM: ``replay_device="cpu"`` means "next to the model".
"""
import inspect
from copy import deepcopy
from importlib import import_module
from typing import Any, Dict

import numpy as np
import torch as t
import torch.nn as nn

# ------------------------------------------------------------------------------------------------ default tables
_COMMON = {"optimizer": "Adam", "criterion": "MSELoss", "criterion_args": (), "criterion_kwargs": {},
           "lr_scheduler": None, "lr_scheduler_args": None, "lr_scheduler_kwargs": None}
_LOCAL_REPLAY = {"replay_size": 500000, "replay_device": "cpu", "replay_buffer": None}
_VISUAL = {"visualize": False, "visualize_dir": ""}
_BATCH_NUM = {"sampler": 10, "learner": 1}


def _models(*names):
    return {"models": list(names), "model_args": tuple(() for _ in names), "model_kwargs": tuple({} for _ in names)}


def _sharded(prefix, group_key):
    return {"learner_process_number": 1, "model_server_group_name": prefix + "_model_server",
            "model_server_members": "all", group_key + "_group_name": prefix, group_key + "_members": "all"}


_DQN = {**_models("QNet", "QNet"), **_COMMON, "batch_size": 100, "epsilon_decay": 0.9999, "update_rate": 0.005,
        "update_steps": None, "learning_rate": 0.001, "discount": 0.99, "gradient_max": np.inf}
_DDPG = {**_models("Actor", "Actor", "Critic", "Critic"), **_COMMON, "batch_size": 100, "update_steps": None,
         "gradient_max": np.inf, "discount": 0.99}
_A2C = {**_models("Actor", "Critic"), **_COMMON, "batch_size": 100, "actor_update_times": 5, "critic_update_times": 10,
        "actor_learning_rate": 0.001, "critic_learning_rate": 0.001, "entropy_weight": None, "value_weight": 0.5,
        "gradient_max": np.inf, "gae_lambda": 1.0, "discount": 0.99, "normalize_advantage": True,
        **_LOCAL_REPLAY, **_VISUAL}

DEFAULTS: Dict[str, Dict[str, Any]] = {
    "DQN": {**_DQN, **_LOCAL_REPLAY, "mode": "double", **_VISUAL},
    "DQNPer": {**_DQN, **_LOCAL_REPLAY, "mode": "double", **_VISUAL},
    "RAINBOW": {**_DQN, **_LOCAL_REPLAY, "mode": "double", **_VISUAL, "value_min": -1.0, "value_max": 1.0,
                "reward_future_steps": 3},
    "DDPGPer": {**_DDPG, "update_rate": 0.001, "actor_learning_rate": 0.0005, "critic_learning_rate": 0.001,
                **_LOCAL_REPLAY, **_VISUAL},
    "A2C": _A2C,
    "A3C": {"grad_server_group_name": "a3c_grad_server", "grad_server_members": "all", **_A2C},
    "DQNApex": {**_sharded("dqn_apex", "apex"), **_DQN, "replay_size": 500000},
    "DDPGApex": {**_sharded("ddpg_apex", "apex"), **_DDPG, "update_rate": 0.005, "learning_rate": 0.001,
                 "replay_size": 500000},
    "IMPALA": {**_sharded("impala", "impala"), **_models("Actor", "Critic"), **_COMMON, "batch_size": 5,
               "learning_rate": 0.001, "isw_clip_c": 1.0, "isw_clip_rho": 1.0, "entropy_weight": None,
               "value_weight": 0.5, "gradient_max": np.inf, "discount": 0.99, "replay_size": 500},
}
_DISTRIBUTED = ("DQNApex", "DDPGApex", "IMPALA")


def generate_config(frame: str, config) -> Any:
    """A copy of ``config`` with ``frame`` set and ``frame_config`` filled from the defaults of ``frame``."""
    config = deepcopy(config)
    config["frame"] = frame
    if frame in _DISTRIBUTED:
        config["batch_num"] = dict(_BATCH_NUM)
    defaults = deepcopy(DEFAULTS[frame])
    config["frame_config"] = {**config["frame_config"], **defaults} if "frame_config" in config else defaults
    return config


# ------------------------------------------------------------------------------------------------ name resolution
def _stack_globals():
    found = {}
    frame = inspect.currentframe()
    while frame is not None:                        # innermost first, so the outermost definition of a name wins
        found.update((k, v) for k, v in frame.f_globals.items() if not k.startswith("__"))
        frame = frame.f_back
    return found


def _import_dotted(path: str):
    parts = path.split(".")
    for cut in range(len(parts) - 1, 0, -1):        # longest importable module prefix, attributes after it
        try:
            obj = import_module(".".join(parts[:cut]))
        except ImportError:
            continue
        for name in parts[cut:]:
            obj = getattr(obj, name)
        return obj
    raise ImportError(f"cannot import {path}")


_KINDS = {
    "model": (None, lambda c: inspect.isclass(c) and issubclass(c, nn.Module)),
    "optimizer": (t.optim, lambda c: inspect.isclass(c) and issubclass(c, t.optim.Optimizer)),
    "lr_scheduler": (t.optim.lr_scheduler, lambda c: inspect.isclass(c) and issubclass(c, t.optim.lr_scheduler.LRScheduler)),
    "criterion": (nn.modules.loss, callable),
}


def resolve(spec, kind: str):
    """``spec``: an object of the right kind (a class; for a criterion any callable), the name of one in torch's
    namespace for that kind, the name of a global of any frame of the call stack, or a dotted import path."""
    namespace, accepts = _KINDS[kind]
    if not isinstance(spec, str):
        if accepts(spec):
            return spec
    else:
        if namespace is not None and hasattr(namespace, spec):
            return getattr(namespace, spec)
        candidate = _stack_globals().get(spec)
        if candidate is not None and accepts(candidate):
            return candidate
        if "." in spec:
            return _import_dotted(spec)
    where = {"model": "", "optimizer": "torch.optim, ", "lr_scheduler": "torch.optim.lr_scheduler, ",
             "criterion": "torch.nn.modules.loss, "}[kind]
    raise ValueError(f"Invalid {kind}: {spec}, it needs to be a suitable class"
                     f"{' or callable' if kind == 'criterion' else ''}, or the name of one in {where}"
                     "the globals of a frame of the call stack, or an importable dotted path")


# ------------------------------------------------------------------------------------------------ construction
def _build_models(f, device):
    classes = [resolve(m, "model") for m in f["models"]]
    return classes, [c(*a, **k).to(device) for c, a, k in zip(classes, f["model_args"], f["model_kwargs"])]


def _resolve_training_objects(f):
    optimizer = resolve(f["optimizer"], "optimizer")
    criterion = resolve(f["criterion"], "criterion")(*f["criterion_args"], **f["criterion_kwargs"])
    lr_scheduler = f["lr_scheduler"] and resolve(f["lr_scheduler"], "lr_scheduler")
    return optimizer, criterion, lr_scheduler


def init_local(cls, config, model_device):
    """DQNPer / DDPGPer / A2C / RAINBOW: ``cls(*models, **frame_config)`` with the names resolved."""
    f = deepcopy(config["frame_config"])
    _, models = _build_models(f, model_device)
    f["optimizer"], f["criterion"], f["lr_scheduler"] = _resolve_training_objects(f)
    if f.get("replay_device") in (None, "cpu"):
        f["replay_device"] = None                   # HBM of the model's device (there is no host replay memory)
    return cls(*models, **f)


def _device(spec):
    """torch.device of ``spec``; a CUDA device without an index becomes the current one (groups and servers allocate
    on it and the library wants an ordinal)."""
    dev = t.device(spec)
    return t.device("cuda", t.cuda.current_device()) if dev.type == "cuda" and dev.index is None else dev


def _members_group(name, members, device):
    """ShardGroup over the given ranks ("all": every rank), or a group of one outside torch.distributed."""
    from ..helpers.servers import _group
    return _group(name, members, device)


def init_sharded(cls, config, model_device, group_key):
    """DQNApex / DDPGApex / IMPALA (apex.py:188-254, impala.py:446-509).  Collective: every rank calls it."""
    import torch.distributed as dist
    from ...parallel import ShardGroup
    from ..helpers.servers import model_server_helper
    f = deepcopy(config["frame_config"])
    device = _device(model_device)
    group = _members_group(f[group_key + "_group_name"], f[group_key + "_members"], device)
    _, models = _build_models(f, device)
    n_learners = int(f["learner_process_number"])
    spmd = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if spmd else 0
    learner_group = ShardGroup(f[group_key + "_group_name"] + "_learners", ranks=list(range(n_learners)),
                               device=device) if spmd else None
    optimizer, criterion, lr_scheduler = _resolve_training_objects(f)
    if group_key == "apex":
        criterion.reduction = "none"                # apex.py:225
    servers = model_server_helper(1, f["model_server_group_name"], f["model_server_members"], device)
    for key in ("optimizer", "criterion", "lr_scheduler"):
        del f[key]
    if spmd and n_learners < group.size():
        f.setdefault("requesters", list(range(n_learners)))
    extra = {"learner_group": learner_group} if group_key == "apex" else {}
    frame = cls(*models, optimizer, criterion, group, servers, lr_scheduler=lr_scheduler, **extra, **f)
    frame.role = "learner" if rank < n_learners else "sampler"
    if frame.role == "sampler" and hasattr(frame, "serve_update"):
        frame.update = lambda *_, **__: (frame.serve_update(), None)[1]
    return frame


def init_a3c(cls, config, model_device):
    """a3c.py:207-248.  Collective.  The member with index 0 of the gradient-server group owns the managed copies of
    actor and critic and their optimizers.  (The reference builds its managed copies from late-binding lambdas that all
    create the LAST model class, a3c.py:221-224; here each server manages a copy of its own model.)"""
    from ..helpers.servers import grad_server_helper
    f = deepcopy(config["frame_config"])
    device = _device(model_device)
    classes, models = _build_models(f, device)
    optimizer, criterion, lr_scheduler = _resolve_training_objects(f)
    creators = [(lambda c=c, a=a, k=k: c(*a, **k)) for c, a, k in zip(classes, f["model_args"], f["model_kwargs"])]
    servers = grad_server_helper(creators, group_name=f["grad_server_group_name"], members=f["grad_server_members"],
                                 optimizer=optimizer, learning_rate=[f["actor_learning_rate"], f["critic_learning_rate"]],
                                 lr_scheduler=lr_scheduler, lr_scheduler_args=f["lr_scheduler_args"] or ((), ()),
                                 lr_scheduler_kwargs=f["lr_scheduler_kwargs"] or ({}, {}), device=device)
    del f["criterion"]
    if f.get("replay_device") in (None, "cpu"):
        f["replay_device"] = None                   # HBM of the model's device (there is no host replay memory)
    return cls(*models, criterion, tuple(servers), **f)
