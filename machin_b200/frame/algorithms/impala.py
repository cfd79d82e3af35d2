"""IMPALA learner with the V-trace correction on the device: reference ``IMPALA`` surface.

Mirror of machin/frame/algorithms/impala.py:69-509.  ``update()`` samples whole episodes from the
(sharded) trajectory buffer, re-evaluates pi(a|s) and V with the user's networks, and hands the flattened
episode-major arrays to ONE fused reverse-scan kernel (csrc/scan.cu) instead of the reference's per-step
Python loop of scalar CPU tensors (:340-362, ~38 us per step).  Reference behaviour kept: vs = 0 on the
last step of every episode, ``next_value[terminal] = 0``, rho/c clipping, sum-reduced policy loss,
``criterion(value, vs)``, actor pushed after every update.
"""
from typing import Any, Callable, Dict, List, Tuple, Union

import numpy as np
import torch as t
import torch.nn as nn

from ... import ops
from ..buffers import IMPALABuffer
from ..transition import Transition
from .base import TorchFramework
from .utils import safe_call


class IMPALA(TorchFramework):
    _config_frame, _config_kind = "IMPALA", "impala"
    _is_top = ["actor", "critic"]
    _is_restorable = ["actor", "critic"]

    def __init__(self, actor: nn.Module, critic: nn.Module, optimizer: Callable, criterion: Callable, impala_group,
                 model_server: Tuple, *_, lr_scheduler: Callable = None, lr_scheduler_args=(), lr_scheduler_kwargs=(),
                 batch_size: int = 5, learning_rate: float = 0.001, isw_clip_c: float = 1.0, isw_clip_rho: float = 1.0,
                 entropy_weight: float = None, value_weight: float = 0.5, gradient_max: float = np.inf,
                 discount: float = 0.99, replay_size: int = 500, requesters: List[int] = None, **__):
        self.batch_size = batch_size
        self.discount = discount
        self.value_weight = value_weight
        self.entropy_weight = entropy_weight
        self.grad_max = gradient_max
        self.isw_clip_c = isw_clip_c
        self.isw_clip_rho = isw_clip_rho
        self.impala_group = impala_group
        self.actor, self.critic = actor, critic
        self.actor_optim = optimizer(self.actor.parameters(), lr=learning_rate)
        self.critic_optim = optimizer(self.critic.parameters(), lr=learning_rate)
        self.replay_buffer = IMPALABuffer(buffer_name="buffer", group=impala_group, buffer_size=replay_size,
                                          requesters=requesters)
        self.is_syncing = True
        self.actor_model_server = model_server[0] if model_server else None
        if lr_scheduler is not None:
            self.actor_lr_sch = lr_scheduler(self.actor_optim, *lr_scheduler_args[0], **lr_scheduler_kwargs[0])
            self.critic_lr_sch = lr_scheduler(self.critic_optim, *lr_scheduler_args[1], **lr_scheduler_kwargs[1])
        self.criterion = criterion
        self._is_using_DP_or_DDP = isinstance(self.actor, (nn.parallel.DataParallel, nn.parallel.DistributedDataParallel))
        super().__init__()

    @property
    def optimizers(self):
        return [self.actor_optim, self.critic_optim]

    @property
    def lr_schedulers(self):
        return [self.actor_lr_sch, self.critic_lr_sch] if hasattr(self, "actor_lr_sch") else []

    @classmethod
    def is_distributed(cls):
        return True

    def set_sync(self, is_syncing):
        self.is_syncing = is_syncing

    def manual_sync(self):
        if not self._is_using_DP_or_DDP and self.actor_model_server is not None:
            self.actor_model_server.pull(self.actor)

    def act(self, state: Dict[str, Any], *_, **__):
        if self.is_syncing and not self._is_using_DP_or_DDP and self.actor_model_server is not None:
            self.actor_model_server.pull(self.actor)
        return safe_call(self.actor, state)

    def _eval_act(self, state: Dict[str, Any], action: Dict[str, Any], *_, **__):
        return safe_call(self.actor, state, action)

    def _criticize(self, state: Dict[str, Any], *_, **__):
        return safe_call(self.critic, state)[0]

    def store_episode(self, episode: List[Union[Transition, Dict]]):
        if len(episode) == 0:
            raise ValueError("Episode must be non-empty.")
        episode[0]["episode_length"] = len(episode)      # impala.py:224-227
        for transition in episode[1:]:
            transition["episode_length"] = 0
        self.replay_buffer.store_episode(
            episode, required_attrs=("state", "action", "next_state", "reward", "action_log_prob", "terminal",
                                     "episode_length"))

    # ------------------------------------------------------------------ learner step (impala.py:242-397)
    def update_device(self, update_value=True, update_policy=True):
        """Collective over ``impala_group``; returns (policy value, value loss) as CUDA scalars on the
        learner(s) and ``None`` on members that only served trajectories."""
        self.actor.train()
        self.critic.train()
        buf = self.replay_buffer
        n_ep, fields, all_length = buf.sample_episodes_device(self.batch_size)
        if not n_ep:
            self._push_actor()
            return None
        group = lambda attr: {f.key: fields[f.name] for f in buf.storage.fields.values()
                              if f.attr == attr and f.kind == "major"}
        state, action, next_state = group("state"), group("action"), group("next_state")
        reward, terminal, action_log_prob = fields["reward"], fields["terminal"], fields["action_log_prob"]
        sum_length = int(sum(all_length))
        if sum_length != terminal.shape[0]:
            raise RuntimeError("Sum length is unequal to tensor total length, an unknown error has occurred.")
        dev = reward.device
        offsets = t.as_tensor(np.concatenate(([0], np.cumsum(all_length))), dtype=t.int64).to(dev)

        _, cur_action_log_prob, entropy, *__ = self._eval_act(state, action)
        cur_action_log_prob = cur_action_log_prob.view(sum_length, 1)
        entropy = entropy.view(sum_length, 1) if entropy.numel() == sum_length else entropy
        value = self._criticize(state).view(sum_length, 1)
        next_value = self._criticize(next_state).view(sum_length, 1)
        vt = ops.vtrace(value.detach(), next_value.detach(), reward.view(sum_length, 1), terminal.view(sum_length, 1),
                        cur_action_log_prob.detach(), action_log_prob.view(sum_length, 1), self.discount,
                        self.isw_clip_c, self.isw_clip_rho, offsets=offsets, max_len=int(max(all_length)))
        act_policy_loss = -(vt["rho"] * cur_action_log_prob * vt["adv"])
        if self.entropy_weight is not None:
            act_policy_loss = act_policy_loss + self.entropy_weight * entropy
        act_policy_loss = act_policy_loss.sum()
        value_loss = self.criterion(value, vt["vs"].type_as(value))
        if update_policy:
            self.actor.zero_grad()
            self._backward(act_policy_loss)
            nn.utils.clip_grad_norm_(self.actor.parameters(), self.grad_max)
            self.actor_optim.step()
        if update_value:
            self.critic.zero_grad()
            self._backward(value_loss)
            nn.utils.clip_grad_norm_(self.critic.parameters(), self.grad_max)
            self.critic_optim.step()
        self._push_actor()
        self.actor.eval()
        self.critic.eval()
        self.last_vtrace = vt
        return -act_policy_loss, value_loss

    def _push_actor(self):
        srv = self.actor_model_server
        if srv is None:
            return
        net = self.actor.module if self._is_using_DP_or_DDP else self.actor
        if self.impala_group.index == srv.src_index:
            srv.push(net, pull_on_fail=False)
        else:
            srv.pull(net)

    def update(self, update_value=True, update_policy=True, **__):
        out = self.update_device(update_value, update_policy)
        if out is None:
            return None
        return out[0].item(), out[1].item()

    def update_lr_scheduler(self):
        if hasattr(self, "actor_lr_sch"):
            self.actor_lr_sch.step()
            self.critic_lr_sch.step()
