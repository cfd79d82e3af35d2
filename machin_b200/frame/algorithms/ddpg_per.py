"""DDPG with prioritized replay on the device: reference ``DDPGPer`` surface.

Mirror of machin/frame/algorithms/ddpg_per.py:15-219 (+ ddpg.py for act*/store_episode).  Learner step:

    PER draw + row gather (sumtree.cu / rows.cu) -> target actor/critic forward (torch) ->
    y = r + gamma * ~terminal * Q'(s', mu'(s'))  (td.cu) -> IS-weighted critic loss + critic step ->
    act_value = Q(s, mu(s)) with the UPDATED critic -> actor step ->
    priority = |act_value - y| (td.cu) -> write-back (sumtree.cu) -> soft updates (td.cu)

Reference quirks kept: priorities use ``act_value`` (not the critic's value of the stored action),
IS weights are cast to the loss dtype here (ddpg_per.py:133-135) but kept float64 in DQNPer, and the
constructor copies the TARGET nets into the online nets (ddpg.py:160-162).
"""
from typing import Any, Callable, Dict, List, Tuple, Union

import numpy as np
import torch as t
import torch.nn as nn

from ... import ops
from ..buffers import PrioritizedBuffer
from ..transition import Transition
from .base import TorchFramework
from .utils import capture_graph, capture_safe, hard_update, model_device, safe_call, soft_update


class DDPGPer(TorchFramework):
    _config_frame, _config_kind = "DDPGPer", "local"
    _is_top = ["actor", "critic", "actor_target", "critic_target"]
    _is_restorable = ["actor_target", "critic_target"]

    def __init__(self, actor: nn.Module, actor_target: nn.Module, critic: nn.Module, critic_target: nn.Module,
                 optimizer: Callable, criterion: Callable, *_, lr_scheduler: Callable = None,
                 lr_scheduler_args: Tuple[Tuple, Tuple] = None, lr_scheduler_kwargs: Tuple[Dict, Dict] = None,
                 batch_size: int = 100, update_rate: float = 0.005, update_steps: Union[int, None] = None,
                 actor_learning_rate: float = 0.0005, critic_learning_rate: float = 0.001, discount: float = 0.99,
                 gradient_max: float = np.inf, replay_size: int = 500000, replay_device=None,
                 replay_buffer: PrioritizedBuffer = None, visualize: bool = False, visualize_dir: str = "", **__):
        self.batch_size = batch_size
        self.update_rate = update_rate
        self.update_steps = update_steps
        self.discount = discount
        self.gradient_max = gradient_max
        self.visualize, self.visualize_dir = visualize, visualize_dir
        self._update_counter = 0
        if update_rate is not None and update_steps is not None:
            raise ValueError("You can only specify one target network update scheme, either by update_rate or "
                             "update_steps, but not both.")
        self.actor, self.actor_target, self.critic, self.critic_target = actor, actor_target, critic, critic_target
        self.actor_optim = optimizer(self.actor.parameters(), lr=actor_learning_rate)
        self.critic_optim = optimizer(self.critic.parameters(), lr=critic_learning_rate)
        device = model_device(actor)
        self.replay_buffer = (PrioritizedBuffer(replay_size, replay_device if replay_device is not None else device)
                              if replay_buffer is None else replay_buffer)
        with t.no_grad():  # ddpg.py:160-162 -- online <- target
            hard_update(self.actor, self.actor_target)
            hard_update(self.critic, self.critic_target)
        if lr_scheduler is not None:
            args, kw = lr_scheduler_args or ((), ()), lr_scheduler_kwargs or ({}, {})
            self.actor_lr_sch = lr_scheduler(self.actor_optim, *args[0], **kw[0])
            self.critic_lr_sch = lr_scheduler(self.critic_optim, *args[1], **kw[1])
        self.criterion = criterion
        if not hasattr(self.criterion, "reduction"):
            raise RuntimeError("Criterion does not have the 'reduction' property, are you using a custom criterion?")
        if getattr(self.criterion, "reduction") != "none":
            self.criterion.reduction = "none"
        super().__init__()

    @property
    def optimizers(self):
        return [self.actor_optim, self.critic_optim]

    @property
    def lr_schedulers(self):
        return [self.actor_lr_sch, self.critic_lr_sch] if hasattr(self, "actor_lr_sch") else []

    # ------------------------------------------------------------------ acting (ddpg.py:179-330)
    def act(self, state: Dict[str, Any], use_target: bool = False, **__):
        return safe_call(self.actor_target if use_target else self.actor, state)[0]

    def act_with_noise(self, state: Dict[str, Any], noise_param: Any = (0.0, 1.0), ratio: float = 1.0,
                       mode: str = "uniform", use_target: bool = False, **__):
        action, *others = safe_call(self.actor_target if use_target else self.actor, state)
        if mode == "uniform":
            lo, hi = noise_param
            noise = t.rand_like(action) * (hi - lo) + lo
        elif mode == "normal":
            noise = t.randn_like(action) * noise_param[1] + noise_param[0]
        else:
            raise ValueError("Unknown noise type: " + str(mode))
        noisy = action + noise * ratio
        return noisy if not others else (noisy, *others)

    def act_discrete(self, state: Dict[str, Any], use_target: bool = False, **__):
        action, *others = safe_call(self.actor_target if use_target else self.actor, state)
        return (t.argmax(action, dim=1).view(action.shape[0], 1), action, *others)

    def act_discrete_with_noise(self, state: Dict[str, Any], use_target: bool = False, choose_max_prob: float = 0.95, **__):
        """ddpg.py:287-331: the actor outputs a probability row per state; near-one-hot rows are softened so that the
        largest component is chosen with ``choose_max_prob``, then one action is sampled per row."""
        probs, *others = safe_call(self.actor_target if use_target else self.actor, state)
        if not (probs.dim() == 2 and bool(t.all((probs.sum(dim=1) - 1.0).abs() < 1e-5)) and bool(t.all(probs >= 0))):
            raise ValueError("Input tensor is not a probability tensor, it must have 2 dimensions (0 and 1), a sum of 1.0 "
                             "for each row in dimension 1, and a positive value for each element.")
        batch_size, action_dim = probs.shape
        if action_dim > 1 and choose_max_prob < 1.0:
            probs = t.softmax(probs * np.log((action_dim - 1) / (1 - choose_max_prob) * choose_max_prob), dim=1)
        # (the reference draws a [batch, 1] block of samples PER ROW and reshapes it to [batch, 1], ddpg.py:327: that only
        #  fits for one state at a time -- which is how acting calls it; kept, so that the random stream is consumed alike)
        result = t.distributions.Categorical(probs).sample([batch_size, 1]).view(batch_size, 1)
        return (result, probs, *others)

    @staticmethod
    def reward_function(reward, discount, next_value, terminal, _):
        """ddpg.py:506-510, as torch ops; the learner step evaluates the same expression inside csrc/td.cu."""
        return reward + discount * ~terminal.to(reward.device) * next_value.to(reward.device)

    def _act(self, state: Dict[str, Any], use_target: bool = False, **__):
        return safe_call(self.actor_target if use_target else self.actor, state)[0]

    def _criticize(self, state: Dict[str, Any], action: Dict[str, Any], use_target: bool = False, **__):
        return safe_call(self.critic_target if use_target else self.critic, state, action)[0]

    def store_episode(self, episode: List[Union[Transition, Dict]]):
        self.replay_buffer.store_episode(
            episode, required_attrs=("state", "action", "reward", "next_state", "terminal"))

    @staticmethod
    def action_transform_function(raw_output_action: Any, *_):
        return {"action": raw_output_action}

    # ------------------------------------------------------------------ learner step (ddpg_per.py:92-185)
    def _sample(self, uniform=None, use_device_params=False):
        # (a sharded buffer's collective sample_device takes no uniforms: they are drawn per shard)
        extra = {"uniform": uniform, "use_device_params": use_device_params} if (uniform is not None or use_device_params) else {}
        n, fields, d_index, d_isw = self.replay_buffer.sample_device(self.batch_size, **extra)
        group = lambda attr: {f.key: fields[f.name] for f in self.replay_buffer.storage.fields.values()
                              if f.attr == attr and f.kind == "major"}
        return n, group("state"), group("action"), fields["reward"], group("next_state"), fields["terminal"], d_index, d_isw

    def _sync_grads(self, params):
        """Hook: gradient averaging across learner processes (see DDPGApex)."""

    def _clip(self, net):
        if np.isfinite(self.gradient_max):     # with max_norm = inf the reference's clip multiplies by exactly 1.0
            nn.utils.clip_grad_norm_(net.parameters(), self.gradient_max)

    def _step(self, update_value, update_policy, update_target, uniform=None, use_device_params=False):
        self.actor.train()
        self.critic.train()
        batch_size, state, action, reward, next_state, terminal, d_index, d_isw = self._sample(uniform, use_device_params)
        others = {}
        with t.no_grad():
            next_action = self.action_transform_function(self._act(next_state, True), next_state, others)
            next_value = self._criticize(next_state, next_action, True).view(batch_size, -1)
            y_i = ops.td_target(next_value, reward, terminal, self.discount)["y"]
        cur_value = self._criticize(state, action)
        value_loss = self.criterion(cur_value, y_i.type_as(cur_value))
        value_loss = (value_loss * d_isw.view(batch_size, 1).type_as(value_loss)).mean()
        if update_value:
            self.critic.zero_grad()
            self._backward(value_loss)
            self._sync_grads(list(self.critic.parameters()))
            self._clip(self.critic)
            self.critic_optim.step()
        cur_action = self.action_transform_function(self._act(state), state, others)
        act_value = self._criticize(state, cur_action)
        act_policy_loss = -act_value.mean()
        abs_error = ops.td_target(next_value, reward, terminal, self.discount, value=act_value.detach())["abs_error"]
        self.replay_buffer.update_priority(abs_error, d_index)
        if update_policy:
            self.actor.zero_grad()
            self._backward(act_policy_loss)
            self._sync_grads(list(self.actor.parameters()))
            self._clip(self.actor)
            self.actor_optim.step()
        if update_target:
            if self.update_rate is not None:
                soft_update(self.actor_target, self.actor, self.update_rate)
                soft_update(self.critic_target, self.critic, self.update_rate)
            else:
                self._update_counter += 1
                if self._update_counter % self.update_steps == 0:
                    hard_update(self.actor_target, self.actor)
                    hard_update(self.critic_target, self.critic)
        self.actor.eval()
        self.critic.eval()
        return -act_policy_loss, value_loss

    # ---- CUDA graph: the whole update (draw, gather, 6 forwards, both backward passes, both optimizers, priority
    # write-back, both soft updates) captured once and replayed; only the uniforms change.  Same protocol as DQNPer's.
    def enable_cuda_graph(self, warmup: int = 3):
        """Single-process prioritized buffer, soft target updates and capturable optimizers (SGD, or Adam / AdamW --
        ``capturable`` is switched on here).  torch's capture protocol needs ``warmup`` real updates first: networks,
        optimizer state, tree and beta are snapshotted before and restored in place after."""
        from ..buffers.prioritized_buffer import PrioritizedBuffer as single_process_per
        if type(self.replay_buffer) is not single_process_per or self.update_rate is None:
            raise RuntimeError("DDPGPer.enable_cuda_graph needs a single-process PrioritizedBuffer and soft target updates")
        buf = self.replay_buffer
        dev = buf.buffer_device
        nets = (self.actor, self.actor_target, self.critic, self.critic_target)
        optims = (self.actor_optim, self.critic_optim)
        tensors_of = lambda m: list(m.parameters()) + list(m.buffers())
        saved_nets = [[x.detach().clone() for x in tensors_of(m)] for m in nets]
        fresh = [len(o.state) == 0 for o in optims]
        saved_optim = [{id(p): {k: (v.clone() if t.is_tensor(v) else v) for k, v in st.items()} for p, st in o.state.items()}
                       for o in optims]
        saved_tree, saved_beta = buf.wt_tree.snapshot(), buf.curr_beta
        for o in optims:
            for group in o.param_groups:
                if "capturable" in group:
                    group["capturable"] = True
            for st in o.state.values():
                if "step" in st and t.is_tensor(st["step"]) and not st["step"].is_cuda:
                    st["step"] = st["step"].to(dev)
        buf.enable_graph_params()
        self._u_static = t.rand(self.batch_size, dtype=t.float64, device=dev)
        self._u_pinned = t.empty(self.batch_size, dtype=t.float64).pin_memory()
        self._u_pinned_np = self._u_pinned.numpy()
        self._h2d_done = t.cuda.Event()
        cur, side = t.cuda.current_stream(dev), t.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with t.cuda.stream(side):
            for _ in range(warmup):
                self._step(True, True, True, self._u_static, True)
                self._u_static.uniform_()
        cur.wait_stream(side)
        t.cuda.synchronize(dev)
        for o in optims:
            o.zero_grad(set_to_none=True)
        from ..._lib import lib
        graph = t.cuda.CUDAGraph()
        launched = lib.mb_launch_count()
        with capture_safe(), capture_graph(graph, side):
            self._static_losses = self._step(True, True, True, self._u_static, True)
        self._graph = graph
        self.graph_launches = lib.mb_launch_count() - launched   # libmachin_b200 kernels per replay
        with t.no_grad():
            for m, saved in zip(nets, saved_nets):
                for x, v in zip(tensors_of(m), saved):
                    x.copy_(v)
            for o, was_fresh, saved in zip(optims, fresh, saved_optim):
                for p, st in o.state.items():
                    for k, v in st.items():
                        if t.is_tensor(v):
                            if was_fresh:
                                v.zero_()
                            elif id(p) in saved:
                                v.copy_(saved[id(p)][k])
        buf.wt_tree.restore(saved_tree)
        buf.curr_beta = saved_beta
        buf.sync_graph_params()
        t.cuda.synchronize(dev)

    def release_graphs(self):
        graph, self._graph = getattr(self, "_graph", None), None
        if graph is not None:
            graph.reset()

    def update_device(self, update_value=True, update_policy=True, update_target=True, uniform: t.Tensor = None):
        """One learner update without host synchronisation -> (policy value, value loss) as CUDA scalars."""
        if getattr(self, "_graph", None) is not None and update_value and update_policy and update_target:
            buf = self.replay_buffer
            if uniform is None:
                self._h2d_done.synchronize()
                self._u_pinned_np[:] = np.random.uniform(size=self.batch_size)       # same stream as :411
                self._u_static.copy_(self._u_pinned, non_blocking=True)
                self._h2d_done.record()
            else:
                self._u_static.copy_(uniform)
            self._graph.replay()
            buf.curr_beta = np.float64(min(1.0, buf.curr_beta + buf.beta_increment_per_sampling))
            return self._static_losses
        return self._step(update_value, update_policy, update_target, uniform)

    def update(self, update_value=True, update_policy=True, update_target=True, concatenate_samples=True, **__):
        pol, val = self.update_device(update_value, update_policy, update_target)
        return pol.item(), val.item()

    def update_lr_scheduler(self):
        if hasattr(self, "actor_lr_sch"):
            self.actor_lr_sch.step()
            self.critic_lr_sch.step()

    def load(self, model_dir, network_map=None, version=-1):
        super().load(model_dir, network_map, version)
        with t.no_grad():
            hard_update(self.actor, self.actor_target)
            hard_update(self.critic, self.critic_target)
