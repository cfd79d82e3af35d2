"""A2C / A3C with returns + GAE computed on the device: reference ``A2C`` / ``A3C`` surface.

Mirror of machin/frame/algorithms/a2c.py:23-497 and a3c.py:12-248.  ``store_episode`` replaces the
reference's per-transition Python loop -- one critic forward and one ``.item()`` per step
(a2c.py:275-313) -- by one batched critic forward over the episode and one reverse-scan kernel
(csrc/scan.cu, float64 like the reference's Python floats).  A3C pushes gradients through the NCCL
gradient server (parallel/server.py) instead of the RPC reduce tree.
"""
from typing import Any, Callable, Dict, List, Tuple, Union

import numpy as np
import torch as t
import torch.nn as nn

from ... import ops
from ..buffers import Buffer
from ..transition import Transition
from .base import TorchFramework
from .utils import model_device, safe_call


class A2C(TorchFramework):
    _config_frame, _config_kind = "A2C", "local"
    _is_top = ["actor", "critic"]
    _is_restorable = ["actor", "critic"]
    on_policy = True      # update() consumes the buffer (a2c.py:419): machin_b200.auto runs one update per stored episode

    def __init__(self, actor: nn.Module, critic: nn.Module, optimizer: Callable, criterion: Callable, *_,
                 lr_scheduler: Callable = None, lr_scheduler_args=(), lr_scheduler_kwargs=(), batch_size: int = 100,
                 actor_update_times: int = 5, critic_update_times: int = 10, actor_learning_rate: float = 0.001,
                 critic_learning_rate: float = 0.001, entropy_weight: float = None, value_weight: float = 0.5,
                 gradient_max: float = np.inf, gae_lambda: float = 1.0, discount: float = 0.99,
                 normalize_advantage: bool = True, replay_size: int = 500000, replay_device=None,
                 replay_buffer: Buffer = None, visualize: bool = False, visualize_dir: str = "", **__):
        self.batch_size = batch_size
        self.actor_update_times, self.critic_update_times = actor_update_times, critic_update_times
        self.discount, self.value_weight, self.entropy_weight = discount, value_weight, entropy_weight
        self.gradient_max, self.gae_lambda, self.normalize_advantage = gradient_max, gae_lambda, normalize_advantage
        self.visualize, self.visualize_dir = visualize, visualize_dir
        self.actor, self.critic = actor, critic
        self.actor_optim = optimizer(self.actor.parameters(), lr=actor_learning_rate)
        self.critic_optim = optimizer(self.critic.parameters(), lr=critic_learning_rate)
        device = model_device(critic)
        self.replay_buffer = (Buffer(replay_size, replay_device if replay_device is not None else device)
                              if replay_buffer is None else replay_buffer)
        if lr_scheduler is not None:
            self.actor_lr_sch = lr_scheduler(self.actor_optim, *lr_scheduler_args[0], **lr_scheduler_kwargs[0])
            self.critic_lr_sch = lr_scheduler(self.critic_optim, *lr_scheduler_args[1], **lr_scheduler_kwargs[1])
        self.criterion = criterion
        super().__init__()

    @property
    def optimizers(self):
        return [self.actor_optim, self.critic_optim]

    @property
    def lr_schedulers(self):
        return [self.actor_lr_sch, self.critic_lr_sch] if hasattr(self, "actor_lr_sch") else []

    def act(self, state: Dict[str, Any], *_, **__):
        return safe_call(self.actor, state)

    def _eval_act(self, state: Dict[str, Any], action: Dict[str, Any], *_, **__):
        return safe_call(self.actor, state, action)

    def _criticize(self, state: Dict[str, Any], *_, **__):
        return safe_call(self.critic, state)[0]

    # ------------------------------------------------------------------ returns + GAE (a2c.py:269-326)
    def compute_value_and_gae(self, episode: List[Union[Transition, Dict]]):
        """-> (returns f64[L], gae f64[L]) CUDA tensors for one episode."""
        dev = model_device(self.critic)
        cat = lambda attr: {k: t.cat([tr[attr][k] for tr in episode], dim=0).to(dev) for k in episode[0][attr]}
        with t.no_grad():
            cv = self._criticize(cat("state")).reshape(-1).float()
            ncv = self._criticize(cat("next_state")).reshape(-1).float() if self.gae_lambda == 0.0 else cv
        reward = t.tensor([float(tr["reward"]) for tr in episode], dtype=t.float64, device=dev)
        terminal = t.tensor([bool(tr["terminal"]) for tr in episode], dtype=t.bool, device=dev)
        out = ops.gae(reward, terminal, cv, ncv, self.discount, self.gae_lambda, n_episodes=1)
        return out["returns"], out["gae"]

    def store_episode(self, episode: List[Union[Transition, Dict]]):
        ret, adv = self.compute_value_and_gae(episode)
        ret, adv = ret.tolist(), adv.tolist()
        for tr, v, a in zip(episode, ret, adv):
            tr["value"] = v
            tr["gae"] = a
        self.replay_buffer.store_episode(
            episode, required_attrs=("state", "action", "next_state", "reward", "value", "gae", "terminal"))

    # ------------------------------------------------------------------ update (a2c.py:328-422)
    def _draw(self, attrs, concatenate):
        """One mini-batch of unique stored steps with the given attributes; the last one is the custom column."""
        return self.replay_buffer.sample_batch(self.batch_size, sample_method="random_unique", concatenate=concatenate,
                                               sample_attrs=attrs, additional_concat_custom_attrs=attrs[-1:])

    def _descend(self, net, optim, loss):
        net.zero_grad()
        self._backward(loss)
        nn.utils.clip_grad_norm_(net.parameters(), self.gradient_max)
        optim.step()

    def _policy_round(self, concatenate, apply):
        n, (state, action, adv) = self._draw(["state", "action", "gae"], concatenate)
        if self.normalize_advantage:
            adv = (adv - adv.mean()) / (adv.std() + 1e-6)
        evaluated = self._eval_act(state, action)
        log_prob = evaluated[1].view(n, 1)
        loss = -(log_prob * adv.type_as(log_prob))
        if self.entropy_weight is not None and evaluated[2] is not None:
            loss += self.entropy_weight * evaluated[2].mean()
        loss = loss.mean()
        if apply:
            self._descend(self.actor, self.actor_optim, loss)
        return loss.item()

    def _value_round(self, concatenate, apply):
        _, (state, target) = self._draw(["state", "value"], concatenate)
        predicted = self._criticize(state)
        loss = self.criterion(target.type_as(predicted), predicted) * self.value_weight
        if apply:
            self._descend(self.critic, self.critic_optim, loss)
        return loss.item()

    def update(self, update_value=True, update_policy=True, concatenate_samples=True, **__):
        """``actor_update_times`` policy rounds, then ``critic_update_times`` value rounds, each on its own mini-batch;
        the buffer is cleared afterwards (on-policy).  Returns (mean policy objective, mean value loss)."""
        for net in (self.actor, self.critic):
            net.train()
        policy = sum(self._policy_round(concatenate_samples, update_policy) for _ in range(self.actor_update_times))
        value = sum(self._value_round(concatenate_samples, update_value) for _ in range(self.critic_update_times))
        self.replay_buffer.clear()
        for net in (self.actor, self.critic):
            net.eval()
        return -policy / self.actor_update_times, value / self.critic_update_times

    def update_lr_scheduler(self):
        if hasattr(self, "actor_lr_sch"):
            self.actor_lr_sch.step()
            self.critic_lr_sch.step()


class _NoStepOptimizer:
    """a3c.py:89-111: the local optimizers never step; the gradient server applies the update."""

    def __init__(self, *_, **__):
        pass

    def zero_grad(self):
        pass

    def step(self):
        pass


class A3C(A2C):
    _config_frame, _config_kind = "A3C", "a3c"

    def __init__(self, actor: nn.Module, critic: nn.Module, criterion: Callable, grad_server: Tuple, *_,
                 batch_size: int = 100, actor_update_times: int = 5, critic_update_times: int = 10,
                 entropy_weight: float = None, value_weight: float = 0.5, gradient_max: float = np.inf,
                 gae_lambda: float = 1.0, discount: float = 0.99, normalize_advantage: bool = True,
                 replay_size: int = 500000, replay_device=None, replay_buffer: Buffer = None, **__):
        super().__init__(actor, critic, _NoStepOptimizer, criterion, batch_size=batch_size,
                         actor_update_times=actor_update_times, critic_update_times=critic_update_times,
                         entropy_weight=entropy_weight, value_weight=value_weight, gradient_max=gradient_max,
                         gae_lambda=gae_lambda, discount=discount, normalize_advantage=normalize_advantage,
                         replay_size=replay_size, replay_device=replay_device, replay_buffer=replay_buffer)
        self.actor_grad_server, self.critic_grad_server = grad_server[0], grad_server[1]
        self.is_syncing = True

    @property
    def optimizers(self):
        return []

    @property
    def lr_schedulers(self):
        return []

    @classmethod
    def is_distributed(cls):
        return True

    def set_sync(self, is_syncing):
        self.is_syncing = is_syncing

    def manual_sync(self):
        self.actor_grad_server.pull(self.actor)
        self.critic_grad_server.pull(self.critic)

    def update(self, update_value=True, update_policy=True, concatenate_samples=True, **__):
        org_sync = self.is_syncing
        self.is_syncing = False
        result = super().update(update_value, update_policy, concatenate_samples)
        self.is_syncing = org_sync
        self.actor_grad_server.push(self.actor)      # a3c.py:163-165: collective over the server's group
        self.critic_grad_server.push(self.critic)
        return result
