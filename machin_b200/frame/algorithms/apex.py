"""APEX learners over sharded device replay: reference ``DQNApex`` / ``DDPGApex`` surface.

Mirror of machin/frame/algorithms/apex.py:18-532.  The reference wires DQNPer/DDPGPer to a
``DistributedPrioritizedBuffer`` reached over RPC and to a ``PushPullModelServer`` (push after every
update, pull before acting).  Here the buffer shards live in the GPUs' HBM and exchange rows over NCCL,
weights travel in one broadcast, and -- when several members are learners -- gradients are averaged
with one all-reduce of a flat buffer (the reference wraps the nets in DistributedDataParallel over gloo,
apex.py:212-221; a DDP-wrapped net is accepted as well).

Execution is SPMD: every member of ``apex_group`` calls ``update()`` (learners) or
``serve_update()`` (members that only hold a shard) in the same order.
"""
from typing import Any, Callable, Dict, Tuple, Union

import numpy as np
import torch as t
import torch.nn as nn

from ..buffers import DistributedPrioritizedBuffer
from .dqn_per import DQNPer


def _flat_grad_all_reduce(params, group):
    """Average gradients over the learner group: one NCCL all-reduce of one flat buffer."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or group is None or group.size() == 1:
        return
    flat = t.cat([g.reshape(-1) for g in grads])
    group.all_reduce_sum(flat)
    flat /= group.size()
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


class DQNApex(DQNPer):
    def __init__(self, qnet: nn.Module, qnet_target: nn.Module, optimizer: Callable, criterion: Callable,
                 apex_group, model_server: Tuple, *_, lr_scheduler: Callable = None, lr_scheduler_args=(),
                 lr_scheduler_kwargs=(), batch_size: int = 100, epsilon_decay: float = 0.9999,
                 update_rate: float = 0.005, update_steps: Union[int, None] = None, learning_rate: float = 0.001,
                 discount: float = 0.99, gradient_max: float = np.inf, replay_size: int = 500000,
                 learner_group=None, requesters=None, **__):
        buffer = DistributedPrioritizedBuffer(buffer_name="buffer", group=apex_group, buffer_size=replay_size,
                                              requesters=requesters)
        super().__init__(qnet, qnet_target, optimizer, criterion, lr_scheduler=lr_scheduler,
                         lr_scheduler_args=lr_scheduler_args or None, lr_scheduler_kwargs=lr_scheduler_kwargs or None,
                         batch_size=batch_size, epsilon_decay=epsilon_decay, update_rate=update_rate,
                         update_steps=update_steps, learning_rate=learning_rate, discount=discount,
                         gradient_max=gradient_max, replay_buffer=buffer)
        self._is_using_DP_or_DDP = isinstance(self.qnet, (nn.parallel.DataParallel, nn.parallel.DistributedDataParallel))
        self.apex_group = apex_group
        self.learner_group = learner_group
        self.qnet_model_server = model_server[0] if model_server else None
        self.is_syncing = True

    @classmethod
    def is_distributed(cls):
        return True

    def set_sync(self, is_syncing):
        self.is_syncing = is_syncing

    def manual_sync(self):
        if not self._is_using_DP_or_DDP and self.qnet_model_server is not None:
            self.qnet_model_server.pull(self.qnet)

    def act_discrete(self, state: Dict[str, Any], use_target: bool = False, **__):
        if self.is_syncing and not use_target and not self._is_using_DP_or_DDP and self.qnet_model_server is not None:
            self.qnet_model_server.pull(self.qnet)
        return super().act_discrete(state, use_target)

    def act_discrete_with_noise(self, state: Dict[str, Any], use_target: bool = False, decay_epsilon: bool = True, **__):
        if self.is_syncing and not use_target and not self._is_using_DP_or_DDP and self.qnet_model_server is not None:
            self.qnet_model_server.pull(self.qnet)
        return super().act_discrete_with_noise(state, use_target, decay_epsilon)

    # ------------------------------------------------------------------ learner step
    def _sample(self, uniform=None, use_device_params=False):
        n, fields, index, d_isw = self.replay_buffer.sample_device(self.batch_size)
        group = lambda attr: {f.key: fields[f.name] for f in self.replay_buffer.storage.fields.values()
                              if f.attr == attr and f.kind == "major"}
        return n, group("state"), group("action"), fields["reward"], group("next_state"), fields["terminal"], index, d_isw

    def _sync_grads(self):
        _flat_grad_all_reduce(list(self.qnet.parameters()), self.learner_group)

    def serve_update(self):
        """For members of ``apex_group`` that hold a replay shard but do not learn (the samplers of
        apex.py's role split): take part in the collective sample / priority write-back of one update."""
        self.replay_buffer.sample_device(self.batch_size)
        self.replay_buffer.update_priority(None, None)
        if self.qnet_model_server is not None:
            self.qnet_model_server.pull(self.qnet)

    def update(self, update_value=True, update_target=True, concatenate_samples=True, **__):
        result = super().update(update_value, update_target, concatenate_samples)
        self.push_weights()
        return result

    def push_weights(self):
        """apex.py:141-150: publish the online net after every update (one broadcast; the lowest learner
        pushes, every other member pulls)."""
        srv = self.qnet_model_server
        if srv is None:
            return
        net = self.qnet.module if self._is_using_DP_or_DDP else self.qnet
        if self.apex_group.index == srv.src_index:
            srv.push(net, pull_on_fail=False)
        else:
            srv.pull(net)


from .ddpg_per import DDPGPer  # noqa: E402


class DDPGApex(DDPGPer):
    """apex.py:256-532: DDPGPer over the sharded buffer; pushes the ACTOR after every update."""

    def __init__(self, actor: nn.Module, actor_target: nn.Module, critic: nn.Module, critic_target: nn.Module,
                 optimizer: Callable, criterion: Callable, apex_group, model_server: Tuple, *_,
                 lr_scheduler: Callable = None, lr_scheduler_args=(), lr_scheduler_kwargs=(), batch_size: int = 100,
                 update_rate: float = 0.005, update_steps: Union[int, None] = None,
                 actor_learning_rate: float = 0.0005, critic_learning_rate: float = 0.001, discount: float = 0.99,
                 gradient_max: float = np.inf, replay_size: int = 500000, learner_group=None, requesters=None, **__):
        buffer = DistributedPrioritizedBuffer(buffer_name="buffer", group=apex_group, buffer_size=replay_size,
                                              requesters=requesters)
        super().__init__(actor, actor_target, critic, critic_target, optimizer, criterion, lr_scheduler=lr_scheduler,
                         lr_scheduler_args=lr_scheduler_args or None, lr_scheduler_kwargs=lr_scheduler_kwargs or None,
                         batch_size=batch_size, update_rate=update_rate, update_steps=update_steps,
                         actor_learning_rate=actor_learning_rate, critic_learning_rate=critic_learning_rate,
                         discount=discount, gradient_max=gradient_max, replay_buffer=buffer)
        self._is_using_DP_or_DDP = isinstance(self.actor, (nn.parallel.DataParallel, nn.parallel.DistributedDataParallel))
        self.apex_group, self.learner_group = apex_group, learner_group
        self.actor_model_server = model_server[0] if model_server else None
        self.is_syncing = True

    @classmethod
    def is_distributed(cls):
        return True

    def set_sync(self, is_syncing):
        self.is_syncing = is_syncing

    def manual_sync(self):
        if not self._is_using_DP_or_DDP and self.actor_model_server is not None:
            self.actor_model_server.pull(self.actor)

    def act(self, state: Dict[str, Any], use_target: bool = False, **kw):
        if self.is_syncing and not use_target and not self._is_using_DP_or_DDP and self.actor_model_server is not None:
            self.actor_model_server.pull(self.actor)
        return super().act(state, use_target, **kw)

    def _sync_grads(self, params):
        _flat_grad_all_reduce(params, self.learner_group)

    def serve_update(self):
        self.replay_buffer.sample_device(self.batch_size)
        self.replay_buffer.update_priority(None, None)
        if self.actor_model_server is not None:
            self.actor_model_server.pull(self.actor)

    def update(self, update_value=True, update_policy=True, update_target=True, concatenate_samples=True, **__):
        result = super().update(update_value, update_policy, update_target, concatenate_samples)
        srv = self.actor_model_server
        if srv is not None:
            net = self.actor.module if self._is_using_DP_or_DDP else self.actor
            if self.apex_group.index == srv.src_index:
                srv.push(net, pull_on_fail=False)
            else:
                srv.pull(net)
        return result
