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
from .utils import capture_graph, capture_safe, safe_call


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
        graph = getattr(self, "_graph", None)
        if graph is not None and update_value and update_policy:
            return self._replay_fixed_shape(graph)
        buf = self.replay_buffer
        n_ep, fields, all_length = buf.sample_episodes_device(self.batch_size)
        if not n_ep:
            self._push_actor()
            return None
        sum_length = int(sum(all_length))
        if sum_length != fields["terminal"].shape[0]:
            raise RuntimeError("Sum length is unequal to tensor total length, an unknown error has occurred.")
        offsets = t.as_tensor(np.concatenate(([0], np.cumsum(all_length))), dtype=t.int64).to(fields["reward"].device)
        out = self._learn(fields, sum_length, offsets, int(max(all_length)), update_value, update_policy)
        self._push_actor()
        return out

    def _clip(self, net):
        if np.isfinite(self.grad_max):     # with max_norm = inf the reference's clip multiplies by exactly 1.0
            nn.utils.clip_grad_norm_(net.parameters(), self.grad_max)

    def _learn(self, fields, sum_length, offsets, max_len, update_value=True, update_policy=True):
        """Everything between the sampled rows and the weight push: re-evaluate pi(a|s) and V, V-trace (one kernel),
        both losses, both optimizer steps."""
        self.actor.train()
        self.critic.train()
        buf = self.replay_buffer
        group = lambda attr: {f.key: fields[f.name] for f in buf.storage.fields.values()
                              if f.attr == attr and f.kind == "major"}
        state, action, next_state = group("state"), group("action"), group("next_state")
        reward, terminal, action_log_prob = fields["reward"], fields["terminal"], fields["action_log_prob"]
        _, cur_action_log_prob, entropy, *__ = self._eval_act(state, action)
        cur_action_log_prob = cur_action_log_prob.view(sum_length, 1)
        entropy = entropy.view(sum_length, 1) if entropy.numel() == sum_length else entropy
        value = self._criticize(state).view(sum_length, 1)
        next_value = self._criticize(next_state).view(sum_length, 1)
        vt = ops.vtrace(value.detach(), next_value.detach(), reward.view(sum_length, 1), terminal.view(sum_length, 1),
                        cur_action_log_prob.detach(), action_log_prob.view(sum_length, 1), self.discount,
                        self.isw_clip_c, self.isw_clip_rho, offsets=offsets, max_len=max_len)
        act_policy_loss = -(vt["rho"] * cur_action_log_prob * vt["adv"])
        if self.entropy_weight is not None:
            act_policy_loss = act_policy_loss + self.entropy_weight * entropy
        act_policy_loss = act_policy_loss.sum()
        value_loss = self.criterion(value, vt["vs"].type_as(value))
        if update_policy:
            self.actor.zero_grad()
            self._backward(act_policy_loss)
            self._clip(self.actor)
            self.actor_optim.step()
        if update_value:
            self.critic.zero_grad()
            self._backward(value_loss)
            self._clip(self.critic)
            self.critic_optim.step()
        self.actor.eval()
        self.critic.eval()
        self.last_vtrace = vt
        return -act_policy_loss, value_loss

    # ---- CUDA graph for trajectories of ONE length (BASELINE C3 / C5: T = 20) on a group of one: the batch has a fixed
    # shape then, so gather -> unpack -> 3 forwards -> V-trace -> both backward passes -> both optimizers is one graph
    # replay; the host still picks the episodes with Python's random like impala.py:58 and copies their slots in.
    def enable_cuda_graph(self, episode_length: int, warmup: int = 3):
        """Every stored episode must have ``episode_length`` steps (a batch with another length falls back to the eager
        update).  Needs a group of one, at least ``batch_size`` stored episodes and capturable optimizers; networks and
        optimizer state are snapshotted before torch's warm-up updates and restored in place after, and Python's random
        stream is put back, so enabling the graph does not change the trajectory."""
        import random
        from ..._lib import lib
        buf = self.replay_buffer
        if self.impala_group.size() != 1:
            raise RuntimeError("IMPALA.enable_cuda_graph: single-member groups only (the trajectory exchange has variable shapes)")
        T, B = int(episode_length), int(self.batch_size)
        dev, n = buf.buffer_device, B * T
        st = {"T": T, "n": n, "idx": t.zeros(n, dtype=t.int64, device=dev), "idx_pin": t.zeros(n, dtype=t.int64).pin_memory(),
              "offsets": t.arange(0, n + 1, T, dtype=t.int64, device=dev), "copied": t.cuda.Event()}
        st["idx_np"] = st["idx_pin"].numpy()
        nets, optims = (self.actor, self.critic), (self.actor_optim, self.critic_optim)
        tensors_of = lambda m: list(m.parameters()) + list(m.buffers())
        saved_nets = [[x.detach().clone() for x in tensors_of(m)] for m in nets]
        fresh = [len(o.state) == 0 for o in optims]
        saved_optim = [{id(p): {k: (v.clone() if t.is_tensor(v) else v) for k, v in s.items()} for p, s in o.state.items()}
                       for o in optims]
        for o in optims:
            for group in o.param_groups:
                if "capturable" in group:
                    group["capturable"] = True
            for s in o.state.values():
                if "step" in s and t.is_tensor(s["step"]) and not s["step"].is_cuda:
                    s["step"] = s["step"].to(dev)
        rng_state = random.getstate()
        if not self._stage_fixed_shape(st):
            random.setstate(rng_state)
            raise RuntimeError(f"IMPALA.enable_cuda_graph needs {B} sampled episodes of exactly {T} steps")
        random.setstate(rng_state)

        def body():
            fields = buf.storage.unpack(buf.storage.gather_packed(st["idx"]))
            return self._learn(fields, n, st["offsets"], T)

        cur, side = t.cuda.current_stream(dev), t.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with t.cuda.stream(side):
            for _ in range(warmup):
                body()
        cur.wait_stream(side)
        t.cuda.synchronize(dev)
        for o in optims:
            o.zero_grad(set_to_none=True)
        graph = t.cuda.CUDAGraph()
        launched = lib.mb_launch_count()
        with capture_safe(), capture_graph(graph, side):
            st["losses"] = body()
        self.graph_launches = lib.mb_launch_count() - launched
        with t.no_grad():
            for m, saved in zip(nets, saved_nets):
                for x, v in zip(tensors_of(m), saved):
                    x.copy_(v)
            for o, was_fresh, saved in zip(optims, fresh, saved_optim):
                for p, s in o.state.items():
                    for k, v in s.items():
                        if t.is_tensor(v):
                            if was_fresh:
                                v.zero_()
                            elif id(p) in saved:
                                v.copy_(saved[id(p)][k])
        t.cuda.synchronize(dev)
        self._graph, self._graph_state = graph, st

    def release_graphs(self):
        graph, self._graph = getattr(self, "_graph", None), None
        if graph is not None:
            graph.reset()

    def _stage_fixed_shape(self, st) -> bool:
        """Pick the episodes of one update on the host (Python's random, impala.py:58) and copy their slots into the
        static index buffer.  False if the batch does not have the captured shape."""
        buf = self.replay_buffer
        local_batch_size = int(np.ceil(self.batch_size / 1))                 # buffer_d.py:143-144 on a group of one
        _units, handles, sizes = buf._local_sample(local_batch_size, "episode")
        if len(sizes) != self.batch_size or len(handles) != st["n"] or any(int(x) != st["T"] for x in sizes):
            st["fallback"] = (handles, sizes)
            return False
        st["copied"].synchronize()                       # the copy that last read the pinned buffer has finished
        st["idx_np"][:] = np.asarray(handles, dtype=np.int64)
        st["idx"].copy_(st["idx_pin"], non_blocking=True)
        st["copied"].record()
        return True

    def _replay_fixed_shape(self, graph):
        st = self._graph_state
        if not self._stage_fixed_shape(st):
            # another shape (an episode of a different length, fewer stored episodes): this update runs eagerly on the
            # episodes that were just drawn, so Python's random stream is consumed exactly once
            handles, sizes = st.pop("fallback")
            if not len(handles):
                self._push_actor()
                return None
            buf = self.replay_buffer
            idx = t.as_tensor(np.asarray(handles, dtype=np.int64)).to(buf.buffer_device)
            fields = buf.storage.unpack(buf.storage.gather_packed(idx))
            offsets = t.as_tensor(np.concatenate(([0], np.cumsum(sizes))), dtype=t.int64).to(buf.buffer_device)
            out = self._learn(fields, int(sum(sizes)), offsets, int(max(sizes)))
            self._push_actor()
            return out
        graph.replay()
        self._push_actor()
        return st["losses"]

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
