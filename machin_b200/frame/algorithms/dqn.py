"""DQN with uniform replay on the device: reference ``DQN`` surface (BASELINE config C1).

Mirror of machin/frame/algorithms/dqn.py:17-563.  The three modes differ in where the bootstrap value comes from
(:356-458): "vanilla" -- max_a Q(s') of the online net, no target net at all (``qnet_target`` IS ``qnet``);
"fixed_target" -- max_a Q_target(s'), soft update only; "double" -- Q_target(s')[argmax_a Q(s')], soft or periodic
hard update.  The replay buffer is the uniform ``Buffer`` (slots drawn with Python's ``random.sample`` like
buffer.py:234, rows gathered on the device); the target comes from the same kernel as DQNPer's (csrc/td.cu), and the
user's criterion -- with whatever reduction it has -- is applied to ``(Q(s)[a], y)`` by torch.
"""
from typing import Callable, Dict, Tuple, Union

import numpy as np
import torch as t
import torch.nn as nn

from ... import ops
from ..buffers import Buffer
from .dqn_per import DQNPer
from .utils import hard_update, model_device, soft_update


class _Reducible:
    reduction = "none"


class DQN(DQNPer):
    _config_frame, _config_kind = "DQN", "local"

    def __init__(self, qnet: nn.Module, qnet_target: nn.Module, optimizer: Callable, criterion: Callable, *_,
                 lr_scheduler: Callable = None, lr_scheduler_args: Tuple[Tuple] = None,
                 lr_scheduler_kwargs: Tuple[Dict] = None, batch_size: int = 100, epsilon_decay: float = 0.9999,
                 update_rate: Union[float, None] = 0.005, update_steps: Union[int, None] = None,
                 learning_rate: float = 0.001, discount: float = 0.99, gradient_max: float = np.inf,
                 replay_size: int = 500000, replay_device: Union[str, t.device] = None, replay_buffer: Buffer = None,
                 mode: str = "double", visualize: bool = False, visualize_dir: str = "", **__):
        if mode not in {"vanilla", "fixed_target", "double"}:
            raise ValueError(f"Unknown DQN mode: {mode}")
        if mode == "vanilla":
            qnet_target = qnet                          # dqn.py:174-175
        if replay_buffer is None:
            replay_buffer = Buffer(replay_size, replay_device if replay_device is not None else model_device(qnet))
        # DQNPer's constructor switches its criterion to reduction="none" (the IS weights need element-wise losses);
        # DQN keeps the criterion exactly as the user built it
        super().__init__(qnet, qnet_target, optimizer, _Reducible(), lr_scheduler=lr_scheduler,
                         lr_scheduler_args=lr_scheduler_args, lr_scheduler_kwargs=lr_scheduler_kwargs, batch_size=batch_size,
                         epsilon_decay=epsilon_decay, update_rate=update_rate, update_steps=update_steps,
                         learning_rate=learning_rate, discount=discount, gradient_max=gradient_max,
                         replay_buffer=replay_buffer, visualize=visualize, visualize_dir=visualize_dir)
        self.criterion, self.mode = criterion, mode

    def enable_cuda_graph(self, warmup: int = 3):
        raise RuntimeError("DQN samples slots with Python's random on the host (buffer.py:234): nothing to capture; "
                           "DQNPer's update is the capturable one")

    # ------------------------------------------------------------------ learner step (dqn.py:333-463)
    def _step(self, update_value, update_target, uniform=None, use_device_params=False) -> t.Tensor:
        batch_size, (state, action, reward, next_state, terminal, _others) = self.replay_buffer.sample_batch(
            self.batch_size, True, sample_method="random_unique",
            sample_attrs=["state", "action", "reward", "next_state", "terminal", "*"])
        self.qnet.train()
        q_value = self._criticize(state)
        action_value = q_value.gather(dim=1, index=self.action_get_function(action).to(device=q_value.device, dtype=t.long))
        with t.no_grad():
            if self.mode == "double":
                online_next, target_next, which = self._criticize(next_state), self._criticize(next_state, True), "double"
            else:                                       # the maximum of ONE net: the online one ("vanilla") or the target
                online_next, which = None, "fixed_target"
                target_next = self._criticize(next_state, self.mode == "fixed_target")
            y_i = ops.td_dqn(None, online_next, target_next, None, reward, terminal, self.discount, mode=which)["y"]
        value_loss = self.criterion(action_value, y_i.type_as(action_value))
        if update_value:
            self.qnet.zero_grad()
            self._backward(value_loss)
            nn.utils.clip_grad_norm_(self.qnet.parameters(), self.grad_max)
            self.qnet_optim.step()
        if update_target and self.mode == "fixed_target":
            soft_update(self.qnet_target, self.qnet, self.update_rate)          # :419-420: soft update only
        elif update_target and self.mode == "double":
            if self.update_rate is not None:
                soft_update(self.qnet_target, self.qnet, self.update_rate)
            else:
                self._update_counter += 1
                if self._update_counter % self.update_steps == 0:
                    hard_update(self.qnet_target, self.qnet)
        self.qnet.eval()
        return value_loss
