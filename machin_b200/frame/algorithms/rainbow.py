"""RAINBOW (distributional double DQN, n-step values, prioritized replay) on the device: reference ``RAINBOW`` surface.

Mirror of machin/frame/algorithms/rainbow.py:8-339 -- SURVEY 8f rank 4, a caller of the same tree / draw / gather /
priority write-back kernels as DQNPer.  The reference moves every network output to the CPU and builds the projected
target distribution from ~30 eager tensor ops (:212-306); here the value block is ONE kernel (csrc/dist_device.cuh,
``ops.td_rainbow``) that also yields the new priorities and d loss / d q_dist, so the learner step keeps every tensor in
HBM like DQNPer's:

    PER draw -> row gather -> user Q-net forward x3 (torch) -> categorical target, row loss, priorities (td.cu) ->
    priority write-back -> backward, clip, optimizer step (torch) -> target update

Reference behaviour kept: the greedy next action comes from the ONLINE net (:229-231), a Bellman target exactly on an
atom loses its mass (:257-262), and the "IS-weighted" loss is the mean of the [B, B] product of :303.  ``qnet`` must
return ``[batch, action_num, atom_num]`` with the last dimension soft-maxed (:44-48).
"""
from typing import Any, Callable, Dict, List, Tuple, Union

import numpy as np
import torch as t

from ... import ops
from ..buffers import PrioritizedBuffer
from ..transition import Transition
from .dqn_per import DQNPer
from .utils import safe_call


class _NoCriterion:
    """RAINBOW has no criterion (the reference passes ``lambda: None``, :89); DQNPer only looks at ``reduction``."""
    reduction = "none"


class RAINBOW(DQNPer):
    _config_frame, _config_kind = "RAINBOW", "local"

    def __init__(self, qnet, qnet_target, optimizer: Callable, value_min, value_max, *_,
                 lr_scheduler: Callable = None, lr_scheduler_args: Tuple[Tuple] = None,
                 lr_scheduler_kwargs: Tuple[Dict] = None, batch_size: int = 100, epsilon_decay: float = 0.9999,
                 update_rate: Union[float, None] = 0.001, update_steps: Union[int, None] = None,
                 learning_rate: float = 0.001, discount: float = 0.99, gradient_max: float = np.inf,
                 reward_future_steps: int = 3, replay_size: int = 500000, replay_device: Union[str, t.device] = None,
                 replay_buffer: PrioritizedBuffer = None, visualize: bool = False, visualize_dir: str = "", **__):
        super().__init__(qnet, qnet_target, optimizer, _NoCriterion(), lr_scheduler=lr_scheduler,
                         lr_scheduler_args=lr_scheduler_args, lr_scheduler_kwargs=lr_scheduler_kwargs,
                         batch_size=batch_size, epsilon_decay=epsilon_decay, update_rate=update_rate,
                         update_steps=update_steps, learning_rate=learning_rate, discount=discount,
                         gradient_max=gradient_max, replay_size=replay_size, replay_device=replay_device,
                         replay_buffer=replay_buffer, visualize=visualize, visualize_dir=visualize_dir)
        self.v_min = value_min
        self.v_max = value_max
        self.reward_future_steps = reward_future_steps

    # ------------------------------------------------------------------ acting (:110-172)
    def _expected_q(self, state, use_target):
        q_dist, *others = safe_call(self.qnet_target if use_target else self.qnet, state)
        support = ops.value_support(self.v_min, self.v_max, q_dist.shape[-1], q_dist.device).view(1, -1)
        return t.sum(support.type_as(q_dist) * q_dist, dim=-1), others

    def act_discrete(self, state: Dict[str, Any], use_target: bool = False, **__):
        q_value, others = self._expected_q(state, use_target)
        result = t.argmax(q_value, dim=1).view(-1, 1)
        return result if len(others) == 0 else (result, *others)

    def act_discrete_with_noise(self, state: Dict[str, Any], use_target: bool = False, decay_epsilon: bool = True, **__):
        q_value, others = self._expected_q(state, use_target)
        action_dim = q_value.shape[1]
        result = t.argmax(q_value, dim=1).view(-1, 1)
        if t.rand([1]).item() < self.epsilon:
            result = t.randint(0, action_dim, [result.shape[0], 1])
        if decay_epsilon:
            self.epsilon *= self.epsilon_decay
        return result if len(others) == 0 else (result, *others)

    # ------------------------------------------------------------------ store (:174-208)
    def store_episode(self, episode: List[Union[Transition, Dict]]):
        """Adds the (truncated) n-step value of every transition as the custom attribute "value", then stores."""
        episode[-1]["value"] = episode[-1]["reward"]
        length = len(episode)
        for i in reversed(range(length)):
            value_sum = 0
            for j in reversed(range(min(self.reward_future_steps, length - i))):
                value_sum = value_sum * self.discount + episode[i + j]["reward"]
            # (tensor rewards give a one-element tensor: its value survives float() exactly, and a plain number is what
            #  the device ring stores as a scalar column -- batches come out as f32[B, 1] either way)
            episode[i]["value"] = float(value_sum) if t.is_tensor(value_sum) else value_sum
        self.replay_buffer.store_episode(
            episode, required_attrs=("state", "action", "next_state", "reward", "value", "terminal"))

    # ------------------------------------------------------------------ learner step (:210-330)
    def _step(self, update_value, update_target, uniform, use_device_params) -> t.Tensor:
        self.qnet.train()
        buf = self.replay_buffer
        _, fields, d_index, d_isw = buf.sample_device(self.batch_size, uniform=uniform, use_device_params=use_device_params)
        group = lambda attr: {f.key: fields[f.name] for f in buf.storage.fields.values()
                              if f.attr == attr and f.kind == "major"}
        state, action, next_state = group("state"), group("action"), group("next_state")
        with t.no_grad():
            next_q_dist = self._criticize(next_state)            # act_discrete(next_state): the online net (:229-231)
            target_next_q_dist = self._criticize(next_state, True)
        q_dist = self._criticize(state)
        td = ops.td_rainbow(q_dist.detach(), next_q_dist, target_next_q_dist, self.action_get_function(action),
                            fields["value"], fields["terminal"], self.v_min, self.v_max,
                            self.discount ** self.reward_future_steps, is_weight=d_isw, want_grad=True)
        value_loss = td["loss_elem"].mean()
        buf.update_priority(td["abs_error"], d_index)
        if update_value:
            self.qnet_optim.zero_grad(set_to_none=True)
            q_dist.backward(td["grad_q"])
            self._sync_grads()
        self._optimizer_and_target(update_value, update_target)
        self.qnet.eval()
        self.last_target_dist = td["target_dist"]
        return value_loss
