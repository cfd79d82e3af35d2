"""Episode datasets over gym-style environments and the launch recipe (surface of machin/auto/envs/openai_gym.py:32-343).

The actor side: one environment per dataset, one ``act`` per step with the observation cast to the models' precision
and moved to their device, transitions collected as the dictionaries ``store_episode`` takes.  ``gym`` itself is optional:
without it the names in ``classic_control.REGISTRY`` resolve to the built-in tasks, anything else raises ImportError.

Kept from the reference: the environment must have a simple (array) observation and a finite horizon; Q networks and
actors take ``state`` (and ``action``) by those names; the DDPG family stores the probability row, not the sampled index,
as the action of a discrete environment (openai_gym.py:121-122); distributed frameworks pull the newest weights once
per episode and not per step (``manual_sync`` / ``set_sync``, :92-94, :138).
Pinned: ``tests/golden/ref_auto_dataset.npz`` -- two cart-pole episodes collected by the unmodified discrete dataset with
the reference's DQN acting; the dataset here collects the same transitions bit for bit (tests/test_auto_cpu.py).
Different: the reference's continuous dataset indexes ``[0]`` into the tensor ``act_with_noise`` returns for a plain actor
(openai_gym.py:215-217) and stores an action WITHOUT its batch dimension; here the stored action is ``[1, action_dim]``
like every other attribute.
Added: an IMPALA transition carries ``action_log_prob`` (its ``store_episode`` requires it, impala.py:224-237; the
reference's dataset omits it and cannot feed IMPALA).  Rendering to video (:97-98, :134-135) belongs to the media
utilities and is not done; ``render_every_episode`` is accepted and ignored."""
import os
from copy import deepcopy
from typing import Any, Dict

import numpy as np
import torch as t

from ...frame.algorithms import A2C, DDPGPer, DQNPer, IMPALA
from ..config import fill_default, is_algorithm_distributed
from ..dataset import DatasetResult, RLDataset, determine_precision
from ..launcher import DistributedLauncher, EarlyStopping, JsonLinesLogger, Launcher, ModelCheckpoint, Trainer
from . import classic_control

_DISCRETE, _CONTINUOUS = ("Discrete", "MultiDiscrete"), ("Box", "MultiBinary")


def _kind(space) -> str:
    return type(space).__name__


def _reset(env):
    out = env.reset()
    return out[0] if isinstance(out, tuple) and len(out) == 2 and isinstance(out[1], dict) else out      # (obs, info) API


def _step(env, action):
    out = env.step(action)
    if len(out) == 5:                                  # obs, reward, terminated, truncated, info
        return out[0], out[1], bool(out[2] or out[3]), out[4]
    return out[0], out[1], bool(out[2]), out[3]


class _GymEpisodes(RLDataset):
    early_stopping_monitor = "total_reward"
    action_kinds = ()

    def __init__(self, frame, env, render_every_episode: int = 100, act_kwargs: Dict[str, Any] = None):
        super().__init__()
        self.frame, self.env, self.render_every_episode = frame, env, render_every_episode
        self.act_kwargs = act_kwargs or {}
        models = [getattr(frame, name) for name in frame.get_top_model_names()]
        self._precision = determine_precision(models)
        self._device = next(models[0].parameters()).device
        self.counter = 0
        assert _kind(env.action_space) in self.action_kinds
        assert _kind(env.observation_space) in _DISCRETE + _CONTINUOUS

    def _observe(self, raw):
        return t.as_tensor(np.asarray(raw), dtype=self._precision).flatten().unsqueeze(0)

    def _act(self, state):
        """-> (what the environment gets, the action to store, extra transition attributes)"""
        raise NotImplementedError

    def __next__(self) -> DatasetResult:
        result, total_reward, terminal = DatasetResult(), 0.0, False
        state = self._observe(_reset(self.env))
        getattr(self.frame, "manual_sync", lambda: None)()
        getattr(self.frame, "set_sync", lambda _: None)(False)
        while not terminal:
            with t.no_grad():
                env_action, stored, extra = self._act({"state": state.to(self._device)})
            raw, reward, terminal, info = _step(self.env, env_action)
            next_state = self._observe(raw)
            total_reward += float(reward)
            result.add_observation({"state": {"state": state}, "action": {"action": stored},
                                    "next_state": {"state": next_state}, "reward": float(reward), "terminal": terminal,
                                    **extra})
            if info:
                result.add_log(info)
            state = next_state
        result.add_log({"total_reward": total_reward})
        getattr(self.frame, "set_sync", lambda _: None)(True)
        self.counter += 1
        return result

    def _policy_gradient_act(self, state):
        action, log_prob, *_ = self.frame.act(state, **self.act_kwargs)
        extra = {"action_log_prob": float(log_prob.reshape(-1)[0])} if isinstance(self.frame, IMPALA) else {}
        return action, extra


class RLGymDiscActDataset(_GymEpisodes):
    """Episodes of an environment that takes one python int per step (openai_gym.py:32-140)."""
    action_kinds = ("Discrete",)

    def _act(self, state):
        if isinstance(self.frame, (A2C, IMPALA)):
            action, extra = self._policy_gradient_act(state)
            return int(action.reshape(-1)[0]), action.detach().cpu().view(1, -1), extra
        if isinstance(self.frame, DQNPer):
            action = self.frame.act_discrete_with_noise(state, **self.act_kwargs)
            action = action[0] if isinstance(action, tuple) else action
            return int(action.reshape(-1)[0]), action.detach().cpu().view(1, 1), {}
        if isinstance(self.frame, DDPGPer):
            action, probs, *_ = self.frame.act_discrete_with_noise(state, **self.act_kwargs)
            return int(action.reshape(-1)[0]), probs.detach().cpu(), {}
        raise RuntimeError(f"Unsupported framework: {type(self.frame)}")


class RLGymContActDataset(_GymEpisodes):
    """Episodes of an environment that takes a float array per step (openai_gym.py:143-248)."""
    action_kinds = ("Box",)

    def _act(self, state):
        if isinstance(self.frame, (A2C, IMPALA)):
            action, extra = self._policy_gradient_act(state)
        elif isinstance(self.frame, DDPGPer):
            action = self.frame.act_with_noise(state, **self.act_kwargs)
            action, extra = (action[0] if isinstance(action, tuple) else action), {}
        else:
            raise RuntimeError(f"Unsupported framework: {type(self.frame)}")
        action = action.detach().cpu().view(1, -1)
        return action.numpy().reshape(self.env.action_space.shape), action, extra


def dataset_for(frame, env, env_config):
    kind = _kind(env.action_space)
    options = dict(render_every_episode=env_config.get("render_every_episode", 100), act_kwargs=env_config.get("act_kwargs"))
    if kind in _DISCRETE:
        return RLGymDiscActDataset(frame, env, **options)
    if kind in _CONTINUOUS:
        return RLGymContActDataset(frame, env, **options)
    raise ValueError(f"Gym environment {env_config.get('env_name')} has action space of type {kind}, which is not supported.")


def _make_env(name: str):
    try:
        import gym
    except ImportError:
        if name in classic_control.REGISTRY:
            return classic_control.make(name)
        raise ImportError(f"gym is not installed and {name} is not one of the built-in tasks "
                          f"{sorted(classic_control.REGISTRY)}") from None
    return gym.make(name)


def gym_env_dataset_creator(frame, env_config):
    return dataset_for(frame, _make_env(env_config["env_name"]), env_config)


def generate_env_config(config: Dict[str, Any] = None):
    """The example config: CartPole-v1 for training and testing (openai_gym.py:272-293)."""
    config = deepcopy(config) if config else {}
    part = {"env_name": "CartPole-v1", "render_every_episode": 100, "act_kwargs": {}}
    return fill_default({"env": "openai_gym", "train_env_config": dict(part), "test_env_config": dict(part)}, config)


def launch_with(config, callbacks, dataset_creator):
    """The recipe of openai_gym.py:295-343: keep the best models by ``total_reward``, stop after
    ``early_stopping_patience`` epochs without improvement, ``episode_per_epoch`` batches per epoch, ``max_episodes``
    batches in all; scalars go to ``<root_dir>/log/train_log/metrics.jsonl``."""
    root = config.get("root_dir") or "./trial"
    trainer = Trainer(
        max_steps=config["max_episodes"], limit_train_batches=config["episode_per_epoch"],
        callbacks=[ModelCheckpoint(os.path.join(root, "model"), monitor="total_reward", mode="max"),
                   EarlyStopping(monitor="total_reward", mode="max", patience=config["early_stopping_patience"]),
                   JsonLinesLogger(os.path.join(root, "log", "train_log", "metrics.jsonl"))] + list(callbacks or []))
    device = config.get("model_device", "cuda")
    module = (DistributedLauncher if is_algorithm_distributed(config) else Launcher)(config, dataset_creator, model_device=device)
    return trainer.fit(module)


def launch(config: Dict[str, Any], pl_callbacks=None):
    return launch_with(config, pl_callbacks, gym_env_dataset_creator)
