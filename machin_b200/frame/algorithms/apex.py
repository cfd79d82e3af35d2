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

from ... import ops
from ..buffers import DistributedPrioritizedBuffer
from .dqn_per import DQNPer
from .utils import capture_graph, capture_safe, soft_update


def _flat_grad_all_reduce(params, group, flat=None):
    """Average gradients over the learner group: one NCCL all-reduce of one flat buffer (``flat``: the
    buffer the gradients already live in as views -- then there is nothing to pack or unpack)."""
    if group is None or group.size() == 1:
        return
    if flat is not None:
        group.all_reduce_sum(flat)
        flat /= group.size()
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = t.cat([g.reshape(-1) for g in grads])
    group.all_reduce_sum(flat)
    flat /= group.size()
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


def _snapshot_optimizer(optim):
    """Tensor state of an optimizer, cloned: what the warm-up updates of a graph capture will change."""
    return {id(p): {k: (v.clone() if t.is_tensor(v) else v) for k, v in st.items()} for p, st in optim.state.items()}


def _restore_optimizer(optim, saved, was_fresh):
    """Undo the warm-up in place (the graph holds these addresses): a fresh optimizer goes back to zeros, a used one
    to the snapshot -- enabling a graph must not change the trajectory."""
    with t.no_grad():
        for p, st in optim.state.items():
            for k, v in st.items():
                if t.is_tensor(v):
                    if was_fresh:
                        v.zero_()
                    elif id(p) in saved and k in saved[id(p)]:
                        v.copy_(saved[id(p)][k])


class _FixedShapeExchange:
    """The fixed-shape shard exchange shared by DQNApex and DDPGApex (mixed into both): every (shard, requester) pair
    always moves ``batch_size`` wire rows and every kernel reads the apportioned counts from device memory, so the whole
    sharded update -- shard sums, apportioning, draws, row exchange, the learner's math, gradient averaging, priority
    exchange, masked tree update -- is ONE CUDA graph replay.  Transport: peer memory over NVLink (csrc/peer.cu: the
    producing kernel stores into the consumer's buffer, one barrier kernel per exchange) or NCCL collectives captured in
    the graph.  Follows prioritized_buffer_d.py:184-303; state lives in ``self._static``."""

    SITE_SUMS, SITE_ROWS, SITE_PRIORITIES, SITE_GRADS, SITE_GRADS_2 = 0, 1, 2, 3, 4
    # Folding the barrier into the producing / consuming kernels (signal from the producer's last CTA after a system-scope
    # fence in EVERY CTA, wait at the top of every consumer CTA) removes a launch per exchange but measured SLOWER than
    # the separate one-warp barrier kernel (N = 2: 170.9 vs 152.6 us / update): a kernel boundary publishes the peer stores
    # for free, `__threadfence_system()` in hundreds of CTAs does not.  Kept as an option for the experiment log.
    fold_barriers = False

    PEER_CHECK_EVERY, PEER_CHECK_SECONDS = 1024, 30.0

    def _check_peers_now_and_then(self):
        """A peer barrier that gave up on a member (parallel/peer.py: 20 s) lets the update run on with stale rows; the
        flag it leaves is read here every PEER_CHECK_EVERY host-side updates, or PEER_CHECK_SECONDS after the last read (one
        small device-to-host read)."""
        st = getattr(self, "_static", None)
        if st is None or st.get("peer") is None:
            return
        import time
        now = time.monotonic()
        self._peer_checks = getattr(self, "_peer_checks", 0) + 1
        # (also whenever updates have become slow: with a member gone every barrier of every update waits out its timeout)
        if self._peer_checks % self.PEER_CHECK_EVERY == 0 or now - getattr(self, "_peer_checked_at", now) > self.PEER_CHECK_SECONDS:
            st["peer"].check()
            self._peer_checked_at = now
        elif not hasattr(self, "_peer_checked_at"):
            self._peer_checked_at = now

    def _static_setup(self, device_apportion: bool, peer_exchange: bool, grad_floats: int):
        """Collective.  Allocates every buffer of the exchange (peer-mapped ones if ``peer_exchange`` works here) and
        stores the state dict in ``self._static``.  ``grad_floats``: size of the largest flat gradient buffer that will be
        averaged through :meth:`_static_all_reduce`."""
        buf, grp = self.replay_buffer, self.apex_group
        dev, P, B = buf.buffer_device, grp.size(), self.batch_size
        if B > 1024 or list(buf._exchange().requesters) != list(range(P)):
            raise RuntimeError("static exchange needs batch <= 1024 and every member learning")
        pitch = buf.storage.pitch
        peer = None
        if peer_exchange and P > 1:
            try:
                from ...parallel.peer import PeerMemory
                peer = PeerMemory.create(grp)
            except Exception as exc:      # noqa: BLE001 -- every member fails alike (same torch, same topology)
                import warnings
                warnings.warn(f"peer exchange unavailable ({exc!r}); the fixed-shape update keeps its NCCL collectives")
                peer = None
        # wire row = packed row + (index, version, is_weight); padded to 16 bytes for the peer stores
        W, cap, capB = pitch + (32 if peer is not None else 24), B, B + P
        z = lambda shape, dt: t.zeros(shape, dtype=dt, device=dev)
        st = {"P": P, "cap": cap, "capB": capB, "W": W, "peer": peer,
              "counts": z(P, t.int64), "counts_pin": t.zeros(P, dtype=t.int64).pin_memory(),
              "params": z((P, 40), t.uint8), "params_pin": t.zeros((P, 40), dtype=t.uint8).pin_memory(),
              "uni": z((P, cap), t.float64), "uni_pin": t.zeros((P, cap), dtype=t.float64).pin_memory(),
              "scale": z(1, t.float64), "scale_pin": t.zeros(1, dtype=t.float64).pin_memory(),
              "send": z((P * cap, W), t.uint8), "recv": z((P * cap, W), t.uint8),
              "merged": z((capB, W), t.uint8), "idx": z(capB, t.int64), "ver": z(capB, t.int64), "isw": z(capB, t.float64),
              "send2": z((P * cap, 24), t.uint8), "recv2": z((P * cap, 24), t.uint8),
              "u_idx": z(P * cap, t.int64), "u_pri": z(P * cap, t.float32), "u_valid": z(P * cap, t.uint8),
              "d_idx": z((P, cap), t.int64), "d_isw": z((P, cap), t.float64),       # row r: the draw for requester r
              "device_apportion": bool(device_apportion),
              "sums": z(P, t.float64), "beta": z(1, t.float64), "nstored": z(1, t.int64),
              # device apportioning lets the host run ahead of the device, so what it writes per update goes through a
              # small ring of pinned buffers guarded by events (copied in FRONT of the replay, not inside the graph)
              "ring_uni": [t.zeros((P, cap), dtype=t.float64).pin_memory() for _ in range(4)],
              "ring_ns": [t.zeros(1, dtype=t.int64).pin_memory() for _ in range(4)],
              "ring_ev": [t.cuda.Event() for _ in range(4)], "ring_i": 0}
        if peer is not None:
            pb = {"recv": peer.alloc((P * cap, W), t.uint8), "recv2": peer.alloc((P * cap, 24), t.uint8),
                  "sums": peer.alloc((P,), t.float64), "grads": peer.alloc((P, int(grad_floats)), t.float32)}
            st["peer_buffers"] = pb
            st["recv"], st["recv2"], st["sums"] = pb["recv"].local, pb["recv2"].local, pb["sums"].local
        self._static = st
        return st

    def _static_sample(self, dense: bool = True):
        """Capturable: shard sums -> apportioning -> this shard's draws for every requester -> row exchange ->
        compaction.  Leaves ``st["idx"]`` / ``st["ver"]`` / ``st["isw"]`` (IS weights already scaled so that a mean over
        the padded rows is the mean over the real ones) and returns the merged, padded batch as dense field tensors
        (``dense``; over the peer transport compaction and unpacking are ONE launch) -- or, ``dense=False``, leaves it as
        wire rows in ``st["merged"]`` and returns None."""
        import torch.distributed as dist
        from ..._lib import lib, ptr, check, stream_ptr
        st, buf, grp = self._static, self.replay_buffer, self.apex_group
        P, me, B, cap, capB, W, peer = st["P"], grp.index, self.batch_size, st["cap"], st["capB"], st["W"], st["peer"]
        pitch, tree, table = buf.storage.pitch, buf.wt_tree.handle, buf.buffer_version_table
        pb = st.get("peer_buffers")
        if st["device_apportion"]:
            if peer is not None:
                # all-gather of the shard sums + barrier + apportioning: ONE one-warp kernel
                from ...parallel.peer import BARRIER_TIMEOUT_NS
                flags, _, state = peer.site(self.SITE_SUMS)
                check(lib.mb_peer_sums_apportion(flags, pb["sums"].ptrs, ptr(buf.weight_sum_tensor()), P, me, state,
                                                 BARRIER_TIMEOUT_NS, B, capB, ptr(st["nstored"]), ptr(st["beta"]),
                                                 float(buf.beta_increment_per_sampling), ptr(st["counts"]), ptr(st["scale"]),
                                                 ptr(st["params"]), stream_ptr()))
            else:
                if P > 1:
                    dist.all_gather_into_tensor(st["sums"], buf.weight_sum_tensor(), group=grp.pg)
                else:
                    st["sums"].copy_(buf.weight_sum_tensor())
                check(lib.mb_wire_apportion(ptr(st["sums"]), P, B, me, capB, ptr(st["nstored"]), ptr(st["beta"]),
                                            float(buf.beta_increment_per_sampling), ptr(st["counts"]), ptr(st["scale"]),
                                            ptr(st["params"]), stream_ptr()))
        else:
            for name in ("counts", "params", "uni", "scale"):
                st[name].copy_(st[name + "_pin"], non_blocking=True)
        if getattr(self, "batched_draw", False):     # opt-in: every requester's draw in ONE launch (one cluster each)
            check(lib.mb_per_sample_multi(tree, ptr(st["uni"]), cap, P, ptr(st["params"]), ptr(st["d_idx"]),
                                          ptr(st["d_isw"]), stream_ptr()))
        else:
            for r in range(P):                       # this shard serves every requester, member order
                check(lib.mb_per_sample(tree, ptr(st["uni"][r]), cap, 0, 0.0, 0.0, ptr(st["params"][r]),
                                        ptr(st["d_idx"][r]), ptr(st["d_isw"][r]), stream_ptr()))
        if peer is not None:
            # gather + metadata + all-to-all as ONE kernel: every row goes straight into its requester's buffer; its last
            # CTA signals the site, and the compaction waits at its top for every member's signal (no barrier launch)
            flags, ticket, state = peer.site(self.SITE_ROWS) if self.fold_barriers else (None, None, None)
            check(lib.mb_wire_push_rows(ptr(buf.storage.data), pitch, buf.storage.max_size, ptr(st["d_idx"]),
                                        ptr(st["d_isw"]), ptr(table), P, cap, me, W, pb["recv"].ptrs, flags, ticket, state,
                                        stream_ptr()))
            if not self.fold_barriers:
                peer.barrier(self.SITE_ROWS)
            if dense:
                fl = st.get("dense")
                if fl is None:
                    fields = buf.storage.fields
                    fl = st["dense"] = {k: t.zeros((capB,) + f.shape, dtype=f.dtype, device=buf.buffer_device) for k, f in fields.items()}
                    st["dense_arr"] = buf.storage._field_array(list(fields), fl)
                check(lib.mb_wire_compact_fields(ptr(st["recv"]), P, cap, W, pitch, ptr(st["counts"]), st["dense_arr"], len(fl),
                                                 capB, ptr(st["idx"]), ptr(st["ver"]), ptr(st["isw"]), ptr(st["scale"]), flags, me,
                                                 state, stream_ptr()))
                return fl
            check(lib.mb_wire_compact_wait(ptr(st["recv"]), P, cap, W, pitch, ptr(st["counts"]), ptr(st["merged"]), capB,
                                           ptr(st["idx"]), ptr(st["ver"]), ptr(st["isw"]), ptr(st["scale"]), flags, me, state,
                                           stream_ptr()))
            return None
        # the P wire blocks are contiguous in requester order: ONE gather and ONE metadata launch for all of them
        buf.storage.gather_into_wire(st["d_idx"].view(-1), st["send"])
        check(lib.mb_rows_pack_meta(ptr(st["send"]), W, pitch, P * cap, ptr(st["d_idx"]), None, ptr(table),
                                    ptr(st["d_isw"]), None, stream_ptr()))
        if P > 1:
            dist.all_to_all_single(st["recv"], st["send"], group=grp.pg)
        else:
            st["recv"].copy_(st["send"])
        check(lib.mb_wire_compact(ptr(st["recv"]), P, cap, W, pitch, ptr(st["counts"]), ptr(st["merged"]), capB,
                                  ptr(st["idx"]), ptr(st["ver"]), ptr(st["isw"]), ptr(st["scale"]), stream_ptr()))
        return buf.storage.unpack(st["merged"]) if dense else None

    def _static_write_back(self, abs_error: t.Tensor):
        """Capturable: route (index, version, |delta|) of every merged row back to the shard that served it, filter by
        version, update this shard's tree (prioritized_buffer_d.py:184-202, :282-291)."""
        import torch.distributed as dist
        from ..._lib import lib, ptr, check, stream_ptr
        from ..buffers import PrioritizedBuffer
        st, buf, grp = self._static, self.replay_buffer, self.apex_group
        P, me, cap, capB, peer, table = st["P"], grp.index, st["cap"], st["capB"], st["peer"], self.replay_buffer.buffer_version_table
        if peer is not None:
            check(lib.mb_wire_push_update(P, cap, me, ptr(st["counts"]), ptr(st["idx"]), ptr(st["ver"]), ptr(abs_error), capB,
                                          st["peer_buffers"]["recv2"].ptrs, stream_ptr()))
            peer.barrier(self.SITE_PRIORITIES)
        else:
            check(lib.mb_wire_scatter_update(ptr(st["send2"]), P, cap, ptr(st["counts"]), ptr(st["idx"]), ptr(st["ver"]),
                                             ptr(abs_error), capB, stream_ptr()))
            if P > 1:
                dist.all_to_all_single(st["recv2"], st["send2"], group=grp.pg)
            else:
                st["recv2"].copy_(st["send2"])
        check(lib.mb_wire_update_mask(ptr(st["recv2"]), P, cap, ptr(st["counts"]), me, ptr(table), table.numel(),
                                      ptr(st["u_idx"]), ptr(st["u_pri"]), ptr(st["u_valid"]), stream_ptr()))
        PrioritizedBuffer.update_priority(buf, st["u_pri"], st["u_idx"], valid=st["u_valid"])

    def _static_write_back_maybe_forked(self, abs_error: t.Tensor):
        """The write-back only feeds the NEXT draw: with ``overlap_priority_update`` it runs on a side stream (a parallel
        branch of the captured graph) beside backward, gradient averaging and the optimizers.  Every member issues its
        exchanges in the same program order.  Returns the (main, side) stream pair to join, or None."""
        if not getattr(self, "overlap_priority_update", False):
            return None
        dev = self.replay_buffer.buffer_device
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = t.cuda.Stream(device=dev)
        forked = (t.cuda.current_stream(dev), self._side_stream)
        forked[1].wait_stream(forked[0])
        with t.cuda.stream(forked[1]):
            self._static_write_back(abs_error)
        return forked

    def _static_all_reduce(self, flat: t.Tensor, site: int = 3):
        """Capturable: average a flat float32 gradient buffer over the members, in place.  Peer transport: my buffer
        into slot ``me`` of every member, barrier, then every member adds the P slots in member order -- bit-identical
        averaged gradients everywhere."""
        import torch.distributed as dist
        from ..._lib import lib, ptr, check, stream_ptr
        st, grp = self._static, self.apex_group
        P, me, peer = st["P"], grp.index, st["peer"]
        n = int(flat.numel())
        if peer is not None:
            slots = st["peer_buffers"]["grads"]
            stride = int(slots.local.shape[1]) * 4
            flags, ticket, state = peer.site(site) if self.fold_barriers else (None, None, None)
            # my buffer into slot `me` of every member, barrier, then the ordered sum
            check(lib.mb_peer_push(ptr(flat), 4 * n, P, me, stride, slots.ptrs, flags, ticket, state, stream_ptr()))
            if not self.fold_barriers:
                peer.barrier(site)
            check(lib.mb_peer_sum(ptr(slots.local), P, n, int(slots.local.shape[1]), float(P), ptr(flat), flags, me, state,
                                  stream_ptr()))
        elif P > 1:
            red = self.learner_group if getattr(self, "learner_group", None) is not None else grp
            dist.all_reduce(flat, group=red.pg)
            flat.div_(red.size())

    def _static_capture(self, body, warmup, nets, optims):
        """torch's capture protocol around ``body`` (warm-up on a side stream, capture, undo everything the warm-up
        touched in place) -> (graph, what body returned during capture)."""
        from ..._lib import lib
        st, buf, grp = self._static, self.replay_buffer, self.apex_group
        dev = buf.buffer_device
        tensors_of = lambda m: list(m.parameters()) + list(m.buffers())
        saved = [[x.detach().clone() for x in tensors_of(m)] for m in nets]
        fresh = [len(o.state) == 0 for o in optims]
        saved_optim = [_snapshot_optimizer(o) for o in optims]
        tree_snap, beta0 = buf.wt_tree.snapshot(), buf.curr_beta
        np_state = np.random.get_state()
        for m in nets:
            m.train()
        st["beta"].fill_(float(beta0))
        self._static_prepare()                           # collective: real counts / uniforms for the warm-up
        cur, side = t.cuda.current_stream(dev), t.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with t.cuda.stream(side):
            for _ in range(warmup):
                body()
        cur.wait_stream(side)
        t.cuda.synchronize(dev)
        graph = t.cuda.CUDAGraph()
        launched = lib.mb_launch_count()
        with capture_safe(), capture_graph(graph, side, capture_error_mode="thread_local"):
            out = body()
        st["graph"] = graph
        self.graph_launches = lib.mb_launch_count() - launched     # libmachin_b200 kernels per replayed update
        with t.no_grad():
            for m, sv in zip(nets, saved):
                for x, v in zip(tensors_of(m), sv):
                    x.copy_(v)
        for o, sv, fr in zip(optims, saved_optim, fresh):
            _restore_optimizer(o, sv, fr)
        buf.wt_tree.restore(tree_snap)
        buf.curr_beta = beta0
        st["beta"].fill_(float(beta0))
        np.random.set_state(np_state)
        t.cuda.synchronize(dev)
        if st["P"] > 1:
            grp.barrier()
        if st["peer"] is not None:
            st["peer"].check()
        return graph, out

    def _static_prepare(self):
        """Host part of a fixed-shape update: read the shard sums (one sync), apportion, draw this shard's
        uniforms from np.random in the reference's order, refresh the pinned parameter blocks."""
        from ...parallel.exchange import shard_sample_sizes
        st, buf, grp = self._static, self.replay_buffer, self.apex_group
        P, me, cap = st["P"], grp.index, st["cap"]
        if st["device_apportion"]:
            # nothing here depends on the device: fixed-size uniform blocks, the stored-row count, and a host mirror
            # of beta (it advances once per requester served, :425-427; the device holds the authoritative value)
            i = st["ring_i"] = (st["ring_i"] + 1) % len(st["ring_ev"])
            st["ring_ev"][i].synchronize()                  # the copy that last used this slot has finished
            st["ring_uni"][i].numpy()[:] = np.random.uniform(size=(P, cap))
            st["ring_ns"][i].numpy()[0] = len(buf.storage)
            st["uni"].copy_(st["ring_uni"][i], non_blocking=True)
            st["nstored"].copy_(st["ring_ns"][i], non_blocking=True)
            st["ring_ev"][i].record()
            for _ in range(P):
                buf.curr_beta = np.float64(min(1.0, buf.curr_beta + buf.beta_increment_per_sampling))
            return self.batch_size
        sums = grp.all_gather_f64(buf.weight_sum_tensor()).view(-1).tolist()
        ssize, total = shard_sample_sizes(sums, self.batch_size)
        n, n_mine = int(sum(ssize)), int(ssize[me])
        views = st.get("np")
        if views is None:                                 # NumPy views of the pinned blocks, taken once
            rec = np.dtype([("beta", "<f8"), ("total", "<f8"), ("n_stored", "<i8"), ("step", "<i8"), ("n_draw", "<i8")])
            views = st["np"] = {"counts": st["counts_pin"].numpy(), "scale": st["scale_pin"].numpy(),
                                "params": st["params_pin"].numpy().view(rec).reshape(P), "uni": st["uni_pin"].numpy()}
        views["counts"][:] = ssize
        views["scale"][0] = st["capB"] / n
        beta, betas = buf.curr_beta, []
        for _ in range(P):                                # one draw per requester; beta advances per draw (:425-427)
            betas.append(beta)
            if n_mine > 0:
                beta = np.float64(min(1.0, beta + buf.beta_increment_per_sampling))
        prm = views["params"]
        prm["beta"], prm["total"], prm["n_stored"], prm["step"], prm["n_draw"] = betas, total, len(buf.storage), 0, n_mine
        if n_mine > 0:
            # P consecutive draws of n_mine uniforms are one draw of P * n_mine from the same MT19937 stream (:411)
            views["uni"][:, :n_mine] = np.random.uniform(size=P * n_mine).reshape(P, n_mine)
        buf.curr_beta = beta
        return n



class DQNApex(_FixedShapeExchange, DQNPer):
    _config_frame, _config_kind = "DQNApex", "apex"

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
        _flat_grad_all_reduce(list(self.qnet.parameters()), self.learner_group, getattr(self, "_flat_grad", None))

    def serve_update(self):
        """For members of ``apex_group`` that hold a replay shard but do not learn (the samplers of
        apex.py's role split): take part in the collective sample / priority write-back of one update."""
        self.replay_buffer.sample_device(self.batch_size)
        self.replay_buffer.update_priority(None, None)
        if self.qnet_model_server is not None:
            self.qnet_model_server.pull(self.qnet)

    # ---- CUDA graphs around the collectives: [unpack, forward x3, TD, backward] and [optimizer, soft update]
    # are captured; the row exchange, the gradient all-reduce, the priority write-back and the weight
    # broadcast stay eager NCCL calls between the replays.  The merged batch has a data-dependent size
    # n in [B, B + P - 1] (ceil apportioning), so the graphs run on B + P padded rows whose IS weight is 0.
    def enable_cuda_graph(self, warmup: int = 3, static_exchange: bool = False, device_apportion: bool = False,
                          peer_exchange: bool = False):
        """``static_exchange``: fixed-shape variant -- every (shard, requester) pair always moves ``batch_size``
        wire rows and the kernels read the apportioned counts from device memory, so draw, both all-to-alls,
        compute, the gradient all-reduce, the optimizer and the tree write-back are ONE graph replay per
        update; the host only reads the shard sums (the reference's own synchronisation point,
        prioritized_buffer_d.py:221-234), draws the uniforms and refreshes ~100 bytes of parameters.
        ``device_apportion`` (with ``static_exchange``): the all-gather of the shard sums is captured too and the
        apportioning (counts, per-draw beta) runs in a one-thread kernel, so the host never waits for the device: per
        update it draws ``members x batch_size`` uniforms and replays.  Throughput mode: np.random is consumed in
        fixed-size blocks instead of the reference's apportioned sizes (identical on a group of one).
        ``peer_exchange`` (with ``static_exchange``): no NCCL inside the update -- sampled rows, priority write-backs,
        shard sums and gradients are stored straight into the other GPUs' buffers over NVLink by this library's
        kernels (csrc/peer.cu), ordered by one small barrier kernel per exchange; falls back to NCCL (with a warning)
        if the buffers cannot be peer-mapped."""
        if static_exchange:
            return self._enable_static_graph(warmup, device_apportion, peer_exchange)
        buf = self.replay_buffer
        dev = buf.buffer_device
        buf._exchange()                                   # collective: agree on the row layout
        cap = self.batch_size + self.apex_group.size()
        params = [p for p in self.qnet.parameters()]
        self._flat_grad = t.zeros(sum(p.numel() for p in params), dtype=t.float32, device=dev)
        o = 0
        for p in params:
            p.grad = self._flat_grad[o:o + p.numel()].view_as(p)
            o += p.numel()
        for group in self.qnet_optim.param_groups:
            if "capturable" in group:
                group["capturable"] = True
        st = {"cap": cap, "rows": t.zeros((cap, buf.storage.pitch + 24), dtype=t.uint8, device=dev),   # wire rows
              "isw": t.zeros(cap, dtype=t.float64, device=dev)}
        fields = buf.storage.fields
        grp = lambda fl, attr: {f.key: fl[f.name] for f in fields.values() if f.attr == attr and f.kind == "major"}

        def compute():
            fl = buf.storage.unpack(st["rows"])
            state, action, next_state = grp(fl, "state"), grp(fl, "action"), grp(fl, "next_state")
            with t.no_grad():
                nq = self._criticize(next_state)
                tq = self._criticize(next_state, True)
            q = self._criticize(state)
            td = ops.td_dqn(q.detach(), nq, tq, self.action_get_function(action), fl["reward"], fl["terminal"],
                            self.discount, is_weight=st["isw"], mode=self.mode, want_grad=True)
            self._flat_grad.zero_()
            q.backward(td["grad_q"])
            return td["abs_error"], td["loss_elem"].mean()

        def apply():
            if np.isfinite(self.grad_max):
                nn.utils.clip_grad_norm_(self.qnet.parameters(), self.grad_max)
            self.qnet_optim.step()
            if self.update_rate is not None:
                soft_update(self.qnet_target, self.qnet, self.update_rate)

        if not self._fused_mse or self.update_rate is None:
            raise RuntimeError("DQNApex.enable_cuda_graph needs an MSELoss criterion and soft target updates")
        saved = [[x.detach().clone() for x in list(m.parameters()) + list(m.buffers())] for m in (self.qnet, self.qnet_target)]
        fresh = len(self.qnet_optim.state) == 0
        saved_optim = _snapshot_optimizer(self.qnet_optim)
        self.qnet.train()
        cur, side = t.cuda.current_stream(dev), t.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with t.cuda.stream(side):
            for _ in range(warmup):
                compute()
                apply()
        cur.wait_stream(side)
        t.cuda.synchronize(dev)
        from ..._lib import lib
        g1, g2 = t.cuda.CUDAGraph(), t.cuda.CUDAGraph()
        launched = lib.mb_launch_count()
        with capture_graph(g1, side):
            st["abs"], st["loss"] = compute()
        with capture_graph(g2, side):
            apply()
        # libmachin_b200 kernels per update that run from the two graphs (replays bypass the C entry points,
        # which count the eager launches of the exchange themselves)
        self.graph_launches = lib.mb_launch_count() - launched
        with t.no_grad():
            for m, sv in zip((self.qnet, self.qnet_target), saved):
                for x, v in zip(list(m.parameters()) + list(m.buffers()), sv):
                    x.copy_(v)
        _restore_optimizer(self.qnet_optim, saved_optim, fresh)
        st["g1"], st["g2"] = g1, g2
        self._apex_graph = st
        t.cuda.synchronize(dev)

    # ------------------------------------------------------------------ fixed-shape, single-graph update
    def _enable_static_graph(self, warmup: int = 3, device_apportion: bool = False, peer_exchange: bool = False):
        from ..._lib import lib, ptr, check, stream_ptr
        buf, grp = self.replay_buffer, self.apex_group
        dev, P = buf.buffer_device, grp.size()
        if not self._fused_mse or self.update_rate is None:
            raise RuntimeError("static exchange needs MSELoss and soft target updates")
        params = list(self.qnet.parameters())
        self._flat_grad = t.zeros(sum(p.numel() for p in params), dtype=t.float32, device=dev)
        o = 0
        for p in params:
            p.grad = self._flat_grad[o:o + p.numel()].view_as(p)
            o += p.numel()
        for group in self.qnet_optim.param_groups:
            if "capturable" in group:
                group["capturable"] = True
        st = self._static_setup(device_apportion, peer_exchange, self._flat_grad.numel())
        capB = st["capB"]
        fields = buf.storage.fields
        grp_f = lambda fl, attr: {f.key: fl[f.name] for f in fields.values() if f.attr == attr and f.kind == "major"}
        prog = self._static_program(st, capB) if self.use_fused_program else None
        from ...model import FusedMLP
        # FusedMLP: the backward kernels write dW / db straight into the flat buffer (no zero-fill, no accumulate launches)
        sinks = isinstance(self.qnet, FusedMLP) and prog is None
        if sinks:
            self.qnet.set_grad_sinks([p.grad for p in params])

        def body():
            fl = self._static_sample(dense=prog is None)
            if prog is not None:
                # fused value math (csrc/mlp.cu): forward x3 + TD + backward straight into the flat gradient buffer
                buf.storage.unpack(st["merged"], names=prog["names"], out=prog["fields"])
                check(lib.mb_mlp_program_run(prog["handle"], stream_ptr()))
                td = {"abs_error": prog["abs"], "loss_elem": prog["loss"]}
            else:
                state, action, next_state = grp_f(fl, "state"), grp_f(fl, "action"), grp_f(fl, "next_state")
                if self.parallel_forwards:
                    # the two no-grad forwards on side streams: parallel branches of the captured graph (as DQNPer does)
                    if self._fwd_streams is None:
                        self._fwd_streams = (t.cuda.Stream(device=dev), t.cuda.Stream(device=dev))
                    cur = t.cuda.current_stream(dev)
                    outs = []
                    for side, use_target in zip(self._fwd_streams, (False, True)):
                        side.wait_stream(cur)
                        with t.cuda.stream(side), t.no_grad():
                            outs.append(self._criticize(next_state, use_target))
                    q = self._criticize(state)
                    for side in self._fwd_streams:
                        cur.wait_stream(side)
                    nq, tq = outs
                else:
                    with t.no_grad():
                        nq = self._criticize(next_state)
                        tq = self._criticize(next_state, True)
                    q = self._criticize(state)
                td = ops.td_dqn(q.detach(), nq, tq, self.action_get_function(action), fl["reward"], fl["terminal"],
                                self.discount, is_weight=st["isw"], mode=self.mode, want_grad=True)
            forked = self._static_write_back_maybe_forked(td["abs_error"])
            loss = None
            if forked is not None:
                with t.cuda.stream(forked[1]):              # the scalar loss only goes to the host: off the critical path too
                    loss = td["loss_elem"].mean()
            if prog is None:
                if not sinks:
                    self._flat_grad.zero_()
                q.backward(td["grad_q"])
            self._static_all_reduce(self._flat_grad, self.SITE_GRADS)
            self._optimizer_and_target(True, True)        # clip (if finite), optimizer step, soft target update
            if forked is None:
                self._static_write_back(td["abs_error"])
                loss = td["loss_elem"].mean()
            else:
                forked[0].wait_stream(forked[1])            # join before the next draw reads the tree
            return loss

        _, st["loss"] = self._static_capture(body, warmup, (self.qnet, self.qnet_target), (self.qnet_optim,))
        if sinks:
            # the captured kernels keep writing into the flat buffer; eager updates (update_value=False, a released graph)
            # go through ordinary autograd accumulation again
            self.qnet.set_grad_sinks(None)

    def _static_program(self, st, rows):
        """Opt-in (``use_fused_program``): the fused value-math program of DQNPer over the merged, padded batch of the
        fixed-shape exchange -- inputs are the unpacked merged rows, the IS weights are the compacted ones (padding
        rows weigh 0), gradients land in the flat buffer the all-reduce runs on."""
        import ctypes as C
        from ...model import FusedMLP
        from ..._lib import lib, ptr, check
        from .dqn_per import DQNPer
        buf = self.replay_buffer
        fields = buf.storage.fields
        sk = [f for f in fields.values() if f.attr == "state" and f.kind == "major"]
        nk = [f for f in fields.values() if f.attr == "next_state" and f.kind == "major"]
        ak = [f for f in fields.values() if f.attr == "action" and f.kind == "major"]
        ok = (isinstance(self.qnet, FusedMLP) and isinstance(self.qnet_target, FusedMLP) and self._fused_mse
              and self.qnet.sizes == self.qnet_target.sizes and len(sk) == len(nk) == len(ak) == 1
              and sk[0].dtype == t.float32 and sk[0].shape == (self.qnet.sizes[0],) and ak[0].dtype == t.int64
              and fields["reward"].dtype == t.float32 and fields["terminal"].dtype == t.bool
              and type(self)._criticize is DQNPer._criticize)
        if not ok:
            raise RuntimeError("use_fused_program needs FusedMLP Q-nets, MSELoss and one float32 state / int64 action key")
        dev, sizes = buf.buffer_device, self.qnet.sizes
        L = len(sizes) - 1
        names = [sk[0].name, ak[0].name, "reward", nk[0].name, "terminal"]
        prog = {"names": names,
                "fields": {n: t.zeros((rows,) + fields[n].shape, dtype=fields[n].dtype, device=dev) for n in names},
                "y": t.empty((rows, 1), dtype=t.float32, device=dev), "av": t.empty((rows, 1), dtype=t.float32, device=dev),
                "abs": t.empty(rows, dtype=t.float32, device=dev), "loss": t.empty((rows, 1), dtype=t.float64, device=dev)}
        plist = lambda net: [p for layer in net.layers for p in (layer.weight, layer.bias)]
        arr = lambda xs: (C.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
        handle = C.c_void_p()
        with t.cuda.device(dev):
            create = lib.mb_dqn_mlp_rows_create if self.fused_kind == "rows" else lib.mb_dqn_mlp_program_create
            check(create(dev.index, L, (C.c_int * (L + 1))(*sizes), rows, arr(plist(self.qnet)), arr(plist(self.qnet_target)),
                         arr([p.grad for p in plist(self.qnet)]), ptr(prog["fields"][names[0]]), ptr(prog["fields"][names[3]]),
                         ptr(prog["fields"][names[1]]), ptr(prog["fields"]["reward"]), ptr(prog["fields"]["terminal"]),
                         ptr(st["isw"]), float(self.discount), 0 if self.mode == "double" else 1, ptr(prog["y"]),
                         ptr(prog["av"]), ptr(prog["abs"]), ptr(prog["loss"]), C.byref(handle)))
        prog["handle"] = handle
        import weakref
        weakref.finalize(self, lib.mb_mlp_program_destroy, handle)
        return prog

    def release_graphs(self):
        """Collective-safe teardown helper: drop every captured graph of this learner (the fixed-shape one holds NCCL
        collectives; torch.distributed's teardown was seen to block while it is alive, tools/nccl_graph_repro.py).
        Later updates run the eager exchange.  Synchronise the device and pass a barrier first."""
        for holder, keys in ((getattr(self, "_static", None), ("graph",)), (getattr(self, "_apex_graph", None), ("g1", "g2"))):
            for key in keys if holder else ():
                graph = holder.pop(key, None)
                if graph is not None:
                    graph.reset()
        self._apex_graph = None
        if hasattr(self.qnet, "set_grad_sinks"):
            self.qnet.set_grad_sinks(None)             # back to ordinary autograd accumulation for the eager path
        super().release_graphs()

    def update_device(self, update_value=True, update_target=True, uniform=None):
        if getattr(self, "_static", None) is not None and "graph" in self._static and update_value and update_target:
            self._static_prepare()
            self._static["graph"].replay()
            return self._static["loss"]
        st = getattr(self, "_apex_graph", None)
        if st is None or not (update_value and update_target):
            return super().update_device(update_value, update_target, uniform)
        buf = self.replay_buffer
        rows, index, isw, _ = buf._exchange().sample(self.batch_size)
        n = int(rows.shape[0])
        if n > st["cap"]:
            raise RuntimeError("merged batch larger than the captured capacity")
        st["rows"][:n].copy_(rows)
        st["isw"].zero_()
        st["isw"][:n].copy_(isw * (st["cap"] / n))   # graphs average over cap rows; padded rows weigh 0
        st["g1"].replay()
        self._sync_grads()
        st["g2"].replay()
        buf.update_priority(st["abs"][:n], index)
        return st["loss"]

    def update(self, update_value=True, update_target=True, concatenate_samples=True, **__):
        result = super().update(update_value, update_target, concatenate_samples)
        self._check_peers_now_and_then()
        self.push_weights()
        return result

    def push_weights(self):
        """apex.py:141-150: publish the online net after every update (one broadcast; the lowest learner
        pushes, every other member pulls)."""
        srv = self.qnet_model_server
        if srv is None:
            return
        if self.learner_group is not None and self.learner_group.size() == self.apex_group.size():
            # every member is a learner applying the same averaged gradients: the weights are already
            # identical everywhere and nobody would consume the broadcast (no actor-only members)
            return
        net = self.qnet.module if self._is_using_DP_or_DDP else self.qnet
        if self.apex_group.index == srv.src_index:
            srv.push(net, pull_on_fail=False)
        else:
            srv.pull(net)


from .ddpg_per import DDPGPer  # noqa: E402


class DDPGApex(_FixedShapeExchange, DDPGPer):
    """apex.py:256-532: DDPGPer over the sharded buffer; pushes the ACTOR after every update."""
    _config_frame, _config_kind = "DDPGApex", "apex"

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

    # ---- the sharded DDPG update as ONE CUDA graph (BASELINE C4): the fixed-shape exchange of _FixedShapeExchange around
    # ddpg_per.py:92-185 -- every member a learner with its own replay shard, critic and actor gradients averaged
    # through two flat buffers.  The merged batch has B + P rows of which sum(counts) are real: the critic loss is
    # weighted by the (scaled) IS weights, which are 0 on padding rows; the policy loss is a mean over the real rows.
    def enable_cuda_graph(self, warmup: int = 3, device_apportion: bool = True, peer_exchange: bool = True):
        buf = self.replay_buffer
        dev = buf.buffer_device
        if self.update_rate is None:
            raise RuntimeError("DDPGApex.enable_cuda_graph needs soft target updates (update_rate)")
        flats = []
        for net, optim in ((self.critic, self.critic_optim), (self.actor, self.actor_optim)):
            params = list(net.parameters())
            flat = t.zeros(sum(p.numel() for p in params), dtype=t.float32, device=dev)
            o = 0
            for p in params:
                p.grad = flat[o:o + p.numel()].view_as(p)
                o += p.numel()
            for group in optim.param_groups:
                if "capturable" in group:
                    group["capturable"] = True
            flats.append(flat)
        self._flat_critic, self._flat_actor = flats
        st = self._static_setup(device_apportion, peer_exchange, max(f.numel() for f in flats))
        capB = st["capB"]
        fields = buf.storage.fields
        grp_f = lambda fl, attr: {f.key: fl[f.name] for f in fields.values() if f.attr == attr and f.kind == "major"}

        def body():
            fl = self._static_sample()
            state, action, next_state = grp_f(fl, "state"), grp_f(fl, "action"), grp_f(fl, "next_state")
            reward, terminal, isw = fl["reward"], fl["terminal"], st["isw"]
            real = (st["ver"] >= 0).to(t.float32)                 # compaction marks padding rows with version -1
            others = {}
            with t.no_grad():
                next_action = self.action_transform_function(self._act(next_state, True), next_state, others)
                next_value = self._criticize(next_state, next_action, True).view(capB, -1)
                y_i = ops.td_target(next_value, reward, terminal, self.discount)["y"]
            cur_value = self._criticize(state, action)
            value_loss = self.criterion(cur_value, y_i.type_as(cur_value))
            value_loss = (value_loss * isw.view(capB, 1).type_as(value_loss)).mean()
            self._flat_critic.zero_()
            self._backward(value_loss)
            self._static_all_reduce(self._flat_critic, self.SITE_GRADS)
            self._clip(self.critic)
            self.critic_optim.step()
            cur_action = self.action_transform_function(self._act(state), state, others)
            act_value = self._criticize(state, cur_action)
            act_policy_loss = -(act_value.view(-1) * real).sum() / real.sum()
            abs_error = ops.td_target(next_value, reward, terminal, self.discount, value=act_value.detach())["abs_error"]
            forked = self._static_write_back_maybe_forked(abs_error)
            self._flat_actor.zero_()
            self._backward(act_policy_loss)
            self._static_all_reduce(self._flat_actor, self.SITE_GRADS_2)
            self._clip(self.actor)
            self.actor_optim.step()
            soft_update(self.actor_target, self.actor, self.update_rate)
            soft_update(self.critic_target, self.critic, self.update_rate)
            if forked is None:
                self._static_write_back(abs_error)
            else:
                forked[0].wait_stream(forked[1])            # join before the next draw reads the tree
            return -act_policy_loss, value_loss

        nets = (self.actor, self.actor_target, self.critic, self.critic_target)
        _, st["losses"] = self._static_capture(body, warmup, nets, (self.actor_optim, self.critic_optim))
        self.actor.eval()
        self.critic.eval()

    def release_graphs(self):
        st = getattr(self, "_static", None)
        graph = st.pop("graph", None) if st else None
        if graph is not None:
            graph.reset()
        super().release_graphs()

    def update_device(self, update_value=True, update_policy=True, update_target=True, uniform=None):
        st = getattr(self, "_static", None)
        if st is not None and "graph" in st and update_value and update_policy and update_target:
            self._static_prepare()
            st["graph"].replay()
            return st["losses"]
        return super().update_device(update_value, update_policy, update_target, uniform)

    def update(self, update_value=True, update_policy=True, update_target=True, concatenate_samples=True, **__):
        result = super().update(update_value, update_policy, update_target, concatenate_samples)
        self._check_peers_now_and_then()
        srv = self.actor_model_server
        if srv is not None:
            net = self.actor.module if self._is_using_DP_or_DDP else self.actor
            if self.apex_group.index == srv.src_index:
                srv.push(net, pull_on_fail=False)
            else:
                srv.pull(net)
        return result
