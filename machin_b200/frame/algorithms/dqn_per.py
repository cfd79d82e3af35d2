"""Double DQN with prioritized replay on the device: reference ``DQNPer`` surface.

Mirror of machin/frame/algorithms/dqn_per.py:15-195 (+ dqn.py for act*/store_episode).  The
learner step keeps every tensor in HBM:

    PER draw (sumtree.cu) -> row gather (rows.cu) -> user Q-net forward x3 (torch) ->
    target / |delta| / IS-weighted loss (td.cu) -> priority write-back (sumtree.cu) ->
    backward, clip, optimizer step (torch) -> soft update (td.cu)

and only ``value_loss.item()`` -- which the reference also returns -- touches the host.
"""
from typing import Any, Callable, Dict, List, Tuple, Union

import numpy as np
import torch as t
import torch.nn as nn

from ... import ops
from ..buffers import PrioritizedBuffer
from ..transition import Transition
from .base import TorchFramework
from .utils import capture_graph, capture_safe, hard_update, model_device, safe_call, soft_update, step_option


class DQNPer(TorchFramework):
    _config_frame, _config_kind = "DQNPer", "local"
    _is_top = ["qnet", "qnet_target"]
    _is_restorable = ["qnet_target"]

    def __init__(self, qnet: nn.Module, qnet_target: nn.Module, optimizer: Callable, criterion: Callable, *_,
                 lr_scheduler: Callable = None, lr_scheduler_args: Tuple[Tuple] = None,
                 lr_scheduler_kwargs: Tuple[Dict] = None, batch_size: int = 100, epsilon_decay: float = 0.9999,
                 update_rate: Union[float, None] = 0.005, update_steps: Union[int, None] = None,
                 learning_rate: float = 0.001, discount: float = 0.99, gradient_max: float = np.inf,
                 replay_size: int = 500000, replay_device: Union[str, t.device] = None,
                 replay_buffer: PrioritizedBuffer = None, visualize: bool = False, visualize_dir: str = "", **__):
        self.batch_size = batch_size
        self.epsilon_decay = epsilon_decay
        self.update_rate = update_rate
        self.update_steps = update_steps
        self.discount = discount
        self.grad_max = gradient_max
        self.visualize = visualize
        self.visualize_dir = visualize_dir
        self.mode = "double"
        self.epsilon = 1
        self._update_counter = 0
        if update_rate is not None and update_steps is not None:
            raise ValueError("You can only specify one target network update scheme, either by update_rate or "
                             "update_steps, but not both.")
        self.qnet = qnet
        self.qnet_target = qnet_target
        self.qnet_optim = optimizer(self.qnet.parameters(), lr=learning_rate)
        device = model_device(qnet)
        self.replay_buffer = (PrioritizedBuffer(replay_size, replay_device if replay_device is not None else device)
                              if replay_buffer is None else replay_buffer)
        # "Make sure target and online networks have the same weight" -- the reference copies the
        # TARGET into the online net here (dqn.py:198-200, hard_update(self.qnet, self.qnet_target))
        with t.no_grad():
            hard_update(self.qnet, self.qnet_target)
        if lr_scheduler is not None:
            self.qnet_lr_sch = lr_scheduler(self.qnet_optim, *((lr_scheduler_args or ((),))[0]),
                                            **((lr_scheduler_kwargs or ({},))[0]))
        self.criterion = criterion
        if not hasattr(self.criterion, "reduction"):
            raise RuntimeError("Criterion does not have the 'reduction' property, are you using a custom criterion?")
        if getattr(self.criterion, "reduction") != "none":
            self.criterion.reduction = "none"   # dqn_per.py:76-87
        # MSELoss: loss, |delta| and dloss/dq all come out of one kernel (td.cu)
        self._fused_mse = isinstance(self.criterion, nn.MSELoss)
        self._graph = None
        self._program = None
        # opt-in: FusedMLP nets + MSE -> forward x3, TD and backward as ONE persistent cooperative kernel
        # (csrc/mlp.cu).  Correct (tests) but currently slower than cuBLAS-backed autograd, see DESIGN.md 8.
        self.use_fused_program = False
        # which fused executor: "program" (one cooperative launch, grid barriers between layers) or "rows" (row-tile
        # kernel without grid barriers + one dW launch; layer outputs <= 256 wide) -- csrc/mlp.cu
        self.fused_kind = "program"
        # opt-in: evaluate Q(s) and Q(s') of the ONLINE net in one forward over [s; s'] (they share weights;
        # only valid for row-wise networks, i.e. no BatchNorm/Dropout): 3 fewer GEMM + epilogue launches
        self.fuse_online_forwards = False
        # opt-in: the priority write-back (one small CTA, ~20 us) only feeds the NEXT draw, so it can run on a
        # second stream beside backward + optimizer + target update; in a captured update this becomes a
        # parallel branch of the graph.  Same kernels, same inputs -> same trajectory.  Single-process buffers only
        # (the sharded write-back is a collective and must stay ordered with the other NCCL calls).
        self.overlap_priority_update = False
        self._side_stream = None
        # opt-in: the three forwards of an update (online Q(s'), target Q(s'), online Q(s)) are independent and each
        # is a chain of sub-10-us kernels; run the two no-grad ones on side streams (parallel graph branches)
        self.parallel_forwards = False
        self._fwd_streams = None
        # opt-in: FusedAdam's step and the soft target update as ONE launch (one pass over the parameters)
        self.fuse_optimizer_and_target = False
        super().__init__()

    # ------------------------------------------------------------------ properties / acting (dqn.py:213-331)
    @property
    def optimizers(self):
        return [self.qnet_optim]

    @optimizers.setter
    def optimizers(self, optimizers):
        self.qnet_optim = optimizers[0]

    @property
    def lr_schedulers(self):
        return [self.qnet_lr_sch] if hasattr(self, "qnet_lr_sch") else []

    def act_discrete(self, state: Dict[str, Any], use_target: bool = False, **__):
        result, *others = safe_call(self.qnet_target if use_target else self.qnet, state)
        result = t.argmax(result, dim=1).view(-1, 1)
        return result if len(others) == 0 else (result, *others)

    def act_discrete_with_noise(self, state: Dict[str, Any], use_target: bool = False, decay_epsilon: bool = True, **__):
        result, *others = safe_call(self.qnet_target if use_target else self.qnet, state)
        action_dim = result.shape[1]
        result = t.argmax(result, dim=1).view(-1, 1)
        if t.rand([1]).item() < self.epsilon:
            result = t.randint(0, action_dim, [result.shape[0], 1])
        if decay_epsilon:
            self.epsilon *= self.epsilon_decay
        return result if len(others) == 0 else (result, *others)

    def _criticize(self, state: Dict[str, Any], use_target: bool = False, **__):
        return safe_call(self.qnet_target if use_target else self.qnet, state)[0]

    def store_episode(self, episode: List[Union[Transition, Dict]]):
        self.replay_buffer.store_episode(
            episode, required_attrs=("state", "action", "reward", "next_state", "terminal"))

    @staticmethod
    def action_get_function(sampled_actions):
        return sampled_actions["action"]

    @staticmethod
    def reward_function(reward, discount, next_value, terminal, _):
        """dqn.py:497-501, as torch ops.  The learner steps do not call it: csrc/td.cu evaluates exactly this expression
        (same operation order) inside the target kernel -- a subclass that changes it must override ``_step`` as well."""
        return reward + discount * ~terminal.to(reward.device) * next_value.to(reward.device)

    # ------------------------------------------------------------------ learner step (dqn_per.py:89-161)
    def _sample(self, uniform=None, use_device_params=False):
        n, fields, d_index, d_isw = self.replay_buffer.sample_device(self.batch_size, uniform=uniform,
                                                                     use_device_params=use_device_params)
        group = lambda attr: {f.key: fields[f.name] for f in self.replay_buffer.storage.fields.values()
                              if f.attr == attr and f.kind == "major"}
        return n, group("state"), group("action"), fields["reward"], group("next_state"), fields["terminal"], d_index, d_isw

    # ---- CUDA graph: the whole update is captured once and replayed; only the uniforms change
    def enable_cuda_graph(self, warmup: int = 3):
        """Capture :meth:`update_device` (draw -> gather -> forward -> TD -> priority write-back ->
        backward -> optimizer -> target update) into one CUDA graph.  Needs a capturable optimizer
        (SGD, or Adam/AdamW with ``capturable=True``, which is switched on here) and a Q-net without
        data-dependent Python control flow.  torch's capture protocol needs ``warmup`` real updates
        first; networks, optimizer state, tree and beta are snapshotted before and restored in place
        after, so enabling the graph does not change the training trajectory.

        Soft target updates only: a hard update every ``update_steps`` is a host-side decision that a replayed graph
        would freeze at whatever it was during capture.  The same holds for a float learning rate changed by an
        ``lr_scheduler``: use a tensor ``lr`` (updated in place) or keep such learners eager."""
        if self.update_rate is None:
            raise RuntimeError(f"{type(self).__name__}.enable_cuda_graph needs soft target updates (update_rate): the hard "
                               "update every update_steps is decided on the host and cannot be replayed")
        if hasattr(self, "qnet_lr_sch") and not all(t.is_tensor(g["lr"]) for g in self.qnet_optim.param_groups):
            raise RuntimeError(f"{type(self).__name__}.enable_cuda_graph: an lr_scheduler changes a float learning rate that "
                               "the captured optimizer step has baked in; give the optimizer a tensor lr or stay eager")
        buf = self.replay_buffer
        dev = buf.buffer_device
        fresh_optim = len(self.qnet_optim.state) == 0
        saved_nets = [[x.detach().clone() for x in list(m.parameters()) + list(m.buffers())]
                      for m in (self.qnet, self.qnet_target)]
        saved_optim = {id(p): {k: (v.clone() if t.is_tensor(v) else v) for k, v in st.items()}
                       for p, st in self.qnet_optim.state.items()}
        saved_tree, saved_beta, saved_counter = buf.wt_tree.snapshot(), buf.curr_beta, self._update_counter
        for group in self.qnet_optim.param_groups:
            if "capturable" in group:
                group["capturable"] = True
        for st in self.qnet_optim.state.values():
            if "step" in st and t.is_tensor(st["step"]) and not st["step"].is_cuda:
                st["step"] = st["step"].to(dev)
        buf.enable_graph_params()
        self._u_static = t.rand(self.batch_size, dtype=t.float64, device=dev)
        self._u_pinned = t.empty(self.batch_size, dtype=t.float64).pin_memory()
        self._u_pinned_np = self._u_pinned.numpy()       # (a view: taking it per update costs microseconds)
        self._h2d_done = t.cuda.Event()
        cur = t.cuda.current_stream(dev)
        side = t.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with t.cuda.stream(side):
            for _ in range(warmup):
                self._step(True, True, self._u_static, True)
                self._u_static.uniform_()
        cur.wait_stream(side)
        t.cuda.synchronize(dev)
        graph = t.cuda.CUDAGraph()
        if not self._program:
            self.qnet_optim.zero_grad(set_to_none=True)
        from ..._lib import lib
        launched = lib.mb_launch_count()
        with capture_safe(), capture_graph(graph, side):
            self._static_loss = self._step(True, True, self._u_static, True)
        self._graph = graph
        self.graph_launches = lib.mb_launch_count() - launched   # libmachin_b200 kernels per replay
        # restore everything the warm-up touched, in place (the graph holds these addresses)
        with t.no_grad():
            for m, saved in zip((self.qnet, self.qnet_target), saved_nets):
                for x, v in zip(list(m.parameters()) + list(m.buffers()), saved):
                    x.copy_(v)
            for p, st in self.qnet_optim.state.items():
                for k, v in st.items():
                    if t.is_tensor(v):
                        if fresh_optim:
                            v.zero_()
                        elif id(p) in saved_optim:
                            v.copy_(saved_optim[id(p)][k])
        buf.wt_tree.restore(saved_tree)
        buf.curr_beta, self._update_counter = saved_beta, saved_counter
        buf.sync_graph_params()
        t.cuda.synchronize(dev)

    def release_graphs(self):
        """Drop the captured update; later updates launch eagerly again.  The device must be idle (synchronise first)."""
        graph, self._graph = self._graph, None
        if graph is not None:
            graph.reset()

    def update_device(self, update_value=True, update_target=True, uniform: t.Tensor = None) -> t.Tensor:
        """One learner update without any host synchronisation; returns the loss as a CUDA scalar."""
        if self._graph is not None and update_value and update_target:
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
            return self._static_loss
        return self._step(update_value, update_target, uniform, False)

    # ---- fused value math: FusedMLP Q-nets + MSE -> one persistent kernel for forward x3, TD, backward
    def _fused_program(self):
        """Build (once) the mb_dqn_mlp_program for this learner, or return None if the nets / buffer do not
        qualify (then the generic autograd path runs)."""
        if not self.use_fused_program:
            return None
        if self._program is not None:
            return self._program or None
        from ...model import FusedMLP
        from ..buffers import PrioritizedBuffer
        buf = self.replay_buffer
        self._program = False
        fields = buf.storage.fields if buf.storage.data is not None else {}
        sk = [f for f in fields.values() if f.attr == "state" and f.kind == "major"]
        nk = [f for f in fields.values() if f.attr == "next_state" and f.kind == "major"]
        ak = [f for f in fields.values() if f.attr == "action" and f.kind == "major"]
        ok = (type(buf) is PrioritizedBuffer and isinstance(self.qnet, FusedMLP) and isinstance(self.qnet_target, FusedMLP)
              and self._fused_mse and self.qnet.sizes == self.qnet_target.sizes and len(sk) == len(nk) == len(ak) == 1
              and sk[0].dtype == t.float32 and sk[0].shape == (self.qnet.sizes[0],) and ak[0].dtype == t.int64
              and fields["reward"].dtype == t.float32 and fields["terminal"].dtype == t.bool
              and type(self)._criticize is DQNPer._criticize)
        if not ok:
            return None
        import ctypes as C
        from ..._lib import lib, ptr, check
        dev, B, sizes = buf.buffer_device, self.batch_size, self.qnet.sizes
        L = len(sizes) - 1
        names = [sk[0].name, ak[0].name, "reward", nk[0].name, "terminal"]
        st = {"names": names, "index": t.empty(B, dtype=t.int64, device=dev), "isw": t.empty(B, dtype=t.float64, device=dev),
              "fields": {n: t.empty((B,) + fields[n].shape, dtype=fields[n].dtype, device=dev) for n in names},
              "y": t.empty((B, 1), dtype=t.float32, device=dev), "av": t.empty((B, 1), dtype=t.float32, device=dev),
              "abs": t.empty(B, dtype=t.float32, device=dev), "loss": t.empty((B, 1), dtype=t.float64, device=dev)}
        plist = lambda net: [p for layer in net.layers for p in (layer.weight, layer.bias)]
        for p in plist(self.qnet):
            p.grad = t.zeros_like(p)       # static gradient buffers, fully overwritten by every run
        arr = lambda xs: (C.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
        handle = C.c_void_p()
        with t.cuda.device(dev):
            create = lib.mb_dqn_mlp_rows_create if self.fused_kind == "rows" else lib.mb_dqn_mlp_program_create
            check(create(
                dev.index, L, (C.c_int * (L + 1))(*sizes), B, arr(plist(self.qnet)), arr(plist(self.qnet_target)),
                arr([p.grad for p in plist(self.qnet)]), ptr(st["fields"][names[0]]), ptr(st["fields"][names[3]]),
                ptr(st["fields"][names[1]]), ptr(st["fields"]["reward"]), ptr(st["fields"]["terminal"]), ptr(st["isw"]),
                float(self.discount), 0 if self.mode == "double" else 1, ptr(st["y"]), ptr(st["av"]), ptr(st["abs"]),
                ptr(st["loss"]), C.byref(handle)))
        st["handle"] = handle
        import weakref
        weakref.finalize(self, lib.mb_mlp_program_destroy, handle)     # device buffers of the program go with the learner
        self._program = st
        return st

    def _step_fused(self, st, update_value, update_target, uniform, use_device_params) -> t.Tensor:
        from ..._lib import lib, check, stream_ptr
        buf = self.replay_buffer
        dev = buf.buffer_device
        side = None
        if self.overlap_priority_update:
            if self._side_stream is None:
                self._side_stream = t.cuda.Stream(device=dev)
            side, cur = self._side_stream, t.cuda.current_stream(dev)
            side.wait_stream(cur)
            with t.cuda.stream(side), t.cuda.device(dev):                 # weight copies: beside the draw and the gather
                check(lib.mb_mlp_program_run_part(st["handle"], 0, stream_ptr()))
        else:
            with t.cuda.device(dev):
                check(lib.mb_mlp_program_run_part(st["handle"], 0, stream_ptr()))
        buf._draw(self.batch_size, uniform=uniform, use_device_params=use_device_params, out=(st["index"], st["isw"]))
        buf.storage.gather(st["index"], names=st["names"], out=st["fields"])
        forked = None
        with t.cuda.device(dev):
            if side is not None:
                cur.wait_stream(side)
            check(lib.mb_mlp_program_run_part(st["handle"], 1, stream_ptr()))        # forward x3, TD, backward-input
            if side is not None:
                # |delta| is ready: the write-back runs beside the weight gradients, the optimizer and the target update
                forked = (cur, side)
                side.wait_stream(cur)
                with t.cuda.stream(side):
                    buf.update_priority(st["abs"], st["index"])
            check(lib.mb_mlp_program_run_part(st["handle"], 2, stream_ptr()))        # dW, db
        if forked is None:
            buf.update_priority(st["abs"], st["index"])
        if update_value:
            self._sync_grads()
        self._optimizer_and_target(update_value, update_target)
        if forked is not None:
            forked[0].wait_stream(forked[1])     # join before the next draw reads the tree
        return st["loss"].mean()

    def _step(self, update_value, update_target, uniform, use_device_params) -> t.Tensor:
        st = self._fused_program()
        if st is not None:
            return self._step_fused(st, update_value, update_target, uniform, use_device_params)
        self.qnet.train()
        fork = self.overlap_priority_update and type(self.replay_buffer) is PrioritizedBuffer
        # with a side branch, everything that only feeds the NEXT update (beta advance, priority write-back) or the host
        # (the scalar loss) runs there; the main stream carries draw -> gather -> forward -> TD -> backward -> optimizer
        self.replay_buffer.defer_beta_advance = bool(fork and use_device_params)
        try:
            batch_size, state, action, reward, next_state, terminal, d_index, d_isw = self._sample(uniform, use_device_params)
        finally:
            self.replay_buffer.defer_beta_advance = False
        if self.fuse_online_forwards and state.keys() == next_state.keys():
            both = self._criticize({k: t.cat([state[k], next_state[k]], dim=0) for k in state})
            q_value, next_q_value = both[:batch_size], both[batch_size:].detach()
            with t.no_grad():
                target_next_q_value = self._criticize(next_state, True)
        elif self.parallel_forwards:
            dev = reward.device
            if self._fwd_streams is None:
                self._fwd_streams = (t.cuda.Stream(device=dev), t.cuda.Stream(device=dev))
            cur = t.cuda.current_stream(dev)
            outs = []
            for side, use_target in zip(self._fwd_streams, (False, True)):
                side.wait_stream(cur)
                with t.cuda.stream(side), t.no_grad():
                    outs.append(self._criticize(next_state, use_target))
            q_value = self._criticize(state)
            for side in self._fwd_streams:
                cur.wait_stream(side)
            next_q_value, target_next_q_value = outs    # (kept alive until after the join; freed on `cur`'s side)
        else:
            with t.no_grad():
                next_q_value = self._criticize(next_state)
                target_next_q_value = self._criticize(next_state, True)
            q_value = self._criticize(state)
        act = self.action_get_function(action)
        td = ops.td_dqn(q_value.detach().contiguous(), next_q_value.contiguous(), target_next_q_value, act, reward, terminal, self.discount,
                        is_weight=d_isw, mode=self.mode, want_grad=self._fused_mse)
        batch_size = int(reward.shape[0])  # a sharded buffer may return more than batch_size rows
        value_loss = None
        if not self._fused_mse:
            action_value = q_value.gather(dim=1, index=act.to(dtype=t.long))
            value_loss = (self.criterion(action_value, td["y"]) * d_isw.view(batch_size, 1)).mean()
        forked = None
        if fork:
            dev = self.replay_buffer.buffer_device
            if self._side_stream is None:
                self._side_stream = t.cuda.Stream(device=dev)
            forked = (t.cuda.current_stream(dev), self._side_stream)
            # the beta advance and the scalar loss (three small torch kernels) do not depend on the tree update: behind it
            # they make the side branch as long as backward + optimizer on the main stream, so they get a branch of their
            # own -- the first forward stream, idle since its join in front of the TD kernel
            third = self._fwd_streams[0] if (step_option("third_branch") and self.parallel_forwards
                                             and self._fwd_streams is not None) else None
            forked[1].wait_stream(forked[0])
            with t.cuda.stream(forked[1]):
                self.replay_buffer.update_priority(td["abs_error"], d_index)
            if third is not None:
                third.wait_stream(forked[0])
            with t.cuda.stream(third if third is not None else forked[1]):
                if use_device_params:
                    self.replay_buffer.advance_beta_device()
                if value_loss is None:
                    value_loss = td["loss_elem"].mean()
            if third is not None:
                forked = (*forked, third)
        else:
            self.replay_buffer.update_priority(td["abs_error"], d_index)
            if value_loss is None:
                value_loss = td["loss_elem"].mean()
        if update_value:
            self.qnet_optim.zero_grad(set_to_none=True)
            if self._fused_mse:
                q_value.backward(td["grad_q"])
            else:
                self._backward(value_loss)
            self._sync_grads()
        self._optimizer_and_target(update_value, update_target)   # (FusedAdam + soft update as one launch when opted in)
        if forked is not None:
            for branch in forked[1:]:
                forked[0].wait_stream(branch)    # join before the next draw reads the tree (td / d_index still alive here)
        self.qnet.eval()
        return value_loss

    def _optimizer_and_target(self, update_value, update_target):
        """optimizer step + target-network update of one learner update (dqn_per.py:141-160)."""
        from ...model import FusedAdam
        if (self.fuse_optimizer_and_target and update_value and update_target and self.update_rate is not None
                and isinstance(self.qnet_optim, FusedAdam) and not np.isfinite(self.grad_max)):
            pairs = {id(p): tp for p, tp in zip(self.qnet.parameters(), self.qnet_target.parameters())}
            self.qnet_optim.step_with_soft_update(pairs, self.update_rate)
            return
        if update_value:
            if np.isfinite(self.grad_max):   # with max_norm = inf the reference's clip multiplies by exactly 1.0
                nn.utils.clip_grad_norm_(self.qnet.parameters(), self.grad_max)
            self.qnet_optim.step()
        if update_target:
            if self.update_rate is not None:
                soft_update(self.qnet_target, self.qnet, self.update_rate)
            else:
                self._update_counter += 1
                if self._update_counter % self.update_steps == 0:
                    hard_update(self.qnet_target, self.qnet)

    def _sync_grads(self):
        """Hook: gradient averaging across learner processes (see DQNApex)."""

    def update(self, update_value=True, update_target=True, concatenate_samples=True, **__):
        return self.update_device(update_value, update_target).item()

    def update_lr_scheduler(self):
        if hasattr(self, "qnet_lr_sch"):
            self.qnet_lr_sch.step()

    def load(self, model_dir, network_map=None, version=-1):
        super().load(model_dir, network_map, version)
        with t.no_grad():
            hard_update(self.qnet, self.qnet_target)
