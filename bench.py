#!/usr/bin/env python
"""bench.py -- APEX-DQN learner hot path on B200 (BASELINE.json configs[1], "C2").

Workload (synthetic, seeded): prioritized replay of 10 M transitions (state f32[84], action i64 in
[0,6), reward f32, terminal bool), batch 512, alpha 0.6, beta 0.4 -> 1; Q-net MLP 84-256-256-6, Adam,
MSE, soft target update.  One "step" = one learner update = PER draw -> row gather -> Q forward x3 ->
TD target / |delta| / IS-weighted loss -> priority write-back -> backward + optimizer -> soft update.

    python bench.py --gpus 1 --steps K --warmup W          # our arm   (one JSON line)
    python bench.py --impl reference --steps K --warmup W  # reference's CPU algorithm (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # N > 1

Numbers printed:
  value     transitions sampled (and learned from) per second, whole job, inputs resident in HBM
            (uniform table pre-staged on the device, loss left on the device)
  e2e       the same through the public DQNPer.update() call: uniforms drawn with np.random on the
            host and copied in every step, loss read back every step
  roofline  the slowest of this library's kernels in the step: algorithmic bytes per launch / CUDA-event
            duration per launch, against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline   the oracle port of the reference learner (oracle/learner.py) timed on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STATE_DIM, N_ACTIONS, HIDDEN = 84, 6, 256
ROW_BYTES = STATE_DIM * 4 * 2 + 8 + 4 + 1  # 685 B of payload per transition
METRIC, UNIT = "transitions_sampled_per_sec", "transitions/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2000)
    p.add_argument("--warmup", type=int, default=50)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                   help="BASELINE.json configuration: c2 APEX-DQN PER 10 M (default: the metric's configuration), c3 IMPALA "
                        "1024 x T=20 on one GPU, c4 APEX-DDPG sharded learner (12.5 M rows per shard), c5 IMPALA actors -> "
                        "learner over NCCL (bench_workloads.py)")
    p.add_argument("--replay-rows", type=int, default=10_000_000)
    p.add_argument("--batch", type=int, default=512)
    p.add_argument("--cpu-storage-rows", type=int, default=50_000)
    p.add_argument("--cpu-steps", type=int, default=1500)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-port", action="store_true", help="CPU arm: the oracle's restatement even if the vendored reference is there")
    p.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per update")
    p.add_argument("--net", default="fused", choices=["torch", "fused", "fused-program", "fused-rows"],
                   help="Q-net implementation: FusedMLP (default since round 2: one csrc/mlp.cu kernel per layer and direction, "
                        "127 vs 135 us/update at N = 1), a plain nn.Module over cuBLAS + eager torch kernels, or FusedMLP + one "
                        "of the fused forward/TD/backward programs (slower, opt-in)")
    p.add_argument("--dynamic-exchange", action="store_true",
                   help="N > 1: exactly-sized eager NCCL exchanges between two CUDA graphs instead of the default "
                        "fixed-shape exchange captured with the rest of the update in ONE graph")
    p.add_argument("--nccl-exchange", action="store_true",
                   help="N > 1, fixed-shape exchange: NCCL collectives captured in the graph instead of the default since round "
                        "2 -- rows, priorities, shard sums and gradients stored into the other GPUs' buffers over NVLink by "
                        "csrc/peer.cu, one barrier kernel per exchange (measured at N = 2: 200 vs 220 us/update; falls back to "
                        "NCCL by itself if the buffers cannot be peer-mapped)")
    p.add_argument("--host-apportion", action="store_true",
                   help="N > 1, fixed-shape exchange: round 1's variant -- the host reads the all-gathered shard sums every "
                        "update and apportions; one draw launch per requester (measured at N = 2: 278 vs 220 us/update)")
    p.add_argument("--watchdog-s", type=float, default=420.0,
                   help="N > 1: leave with exit code 5 if the run has not finished after this many seconds "
                        "(a stuck collective must not hold the box)")
    p.add_argument("--fuse-forwards", action="store_true", help="one online forward over [s; s'] instead of two passes like the "
                   "reference (measured: no gain, 0.204 vs 0.205 ms/update)")
    p.add_argument("--parallel-forwards", action="store_true",
                   help="N = 1, opt-in: the two no-grad forwards on side streams (parallel graph branches)")
    p.add_argument("--fuse-optimizer", action="store_true",
                   help="N = 1, opt-in: FusedAdam step and soft target update as one launch")
    p.add_argument("--overlap-update", action="store_true",
                   help="N = 1, opt-in: priority write-back on a parallel graph branch beside backward + optimizer")
    p.add_argument("--adam", default="fused", choices=["fused", "torch"], help="FusedAdam (csrc/mlp.cu) or torch.optim.Adam(fused=True)")
    p.add_argument("--serial", action="store_true",
                   help="switch the three stream-level options off (round 1's launch order: one chain of kernels)")
    a = p.parse_args()
    a.static_exchange = not a.dynamic_exchange
    # default since round 2 (profiles/r2_n2_variants.md): shard sums all-gathered and apportioned inside the graph (no host
    # read per update), the draws for all requesters in one launch
    a.device_apportion = a.batched_draw = not a.host_apportion
    a.peer_exchange = not a.nccl_exchange
    if not a.serial and a.net in ("torch", "fused"):
        # default since round 2 (measured at N = 1: 198 -> 138 us/update, profiles/r2_ab_variants.md): the two no-grad
        # forwards and the priority write-back are parallel branches of the captured graph, Adam + soft update one launch
        a.parallel_forwards = a.overlap_update = a.fuse_optimizer = True
    return a


def workload_name(a):
    return (f"C2 APEX-DQN learner: PER {a.replay_rows} x (state {STATE_DIM}, act {N_ACTIONS}), batch {a.batch}, "
            f"MLP {STATE_DIM}-{HIDDEN}-{HIDDEN}-{N_ACTIONS}")


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_learner_rate(a, steps, warmup, budget_s=25.0):
    """Oracle port of the reference learner on the host cores -> (transitions/s, ms/step, threads, steps).

    The MLP is tiny, so "all cores" is not the fastest setting (128 threads is ~500x slower than 8 on
    the GPU box): the thread count is calibrated over {1,2,4,8,16,32,all} and the best one is timed.
    The run stops after ``steps`` updates or ``budget_s`` seconds, whichever comes first."""
    import torch as t
    np.random.seed(0)
    learner, kind = make_cpu_learner(a)
    ncpu = os.cpu_count() or 1
    best, best_dt = 1, float("inf")
    for th in [c for c in (1, 2, 4, 8, 16, 32) if c <= ncpu] + ([ncpu] if ncpu <= 64 else []):
        t.set_num_threads(th)
        learner.step()
        t0 = time.perf_counter()
        for _ in range(3):
            learner.step()
        dt = time.perf_counter() - t0
        if dt < best_dt:
            best, best_dt = th, dt
    t.set_num_threads(best)
    for _ in range(warmup):
        learner.step()
    t0 = time.perf_counter()
    done = 0
    while done < steps and time.perf_counter() - t0 < budget_s:
        learner.step()
        done += 1
    dt = time.perf_counter() - t0
    return a.batch * done / dt, dt / done * 1e3, best, done, kind


def make_cpu_learner(a):
    """The CPU arm: the UNMODIFIED reference (machin's own DQNPer over its own PrioritizedBuffer, vendored by
    oracle/vendor_ref.py into oracle/_ref, or mounted at /root/reference) when it is there -- kind "reference" -- else the
    oracle's restatement of the same update (oracle/learner.py) -- kind "port"."""
    from oracle import ref_shim
    if not a.cpu_port and ref_shim.available():
        try:
            from oracle.ref_learner import ReferenceDqnPerLearner
            return ReferenceDqnPerLearner(a.replay_rows, a.cpu_storage_rows, STATE_DIM, N_ACTIONS, a.batch, hidden=HIDDEN), "reference"
        except Exception as exc:   # noqa: BLE001 -- e.g. an import the shim does not cover on this box: say so, use the port
            sys.stderr.write(f"bench.py: the vendored reference could not be set up ({exc!r}); timing the oracle port instead\n")
    from oracle.learner import CpuDqnPerLearner
    return CpuDqnPerLearner(a.replay_rows, a.cpu_storage_rows, STATE_DIM, N_ACTIONS, a.batch, hidden=HIDDEN), "port"


def cpu_sample_text(kind, done, threads, a):
    what = ("machin's own DQNPer.update() over its own PrioritizedBuffer (unmodified reference, oracle/_ref): NumPy float64 "
            "sum tree at full size" if kind == "reference" else
            "oracle/learner.py (restatement of DQNPer.update(): NumPy float64 tree at full size")
    return (f"{done} full learner updates of {what} ({a.replay_rows} leaves), Python-object storage capped at "
            f"{a.cpu_storage_rows} transitions, CPU torch MLP; best of thread counts 1..32 -> {threads} threads, "
            f"{os.cpu_count()} cores present")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, threads, done, kind = cpu_learner_rate(a, a.steps, max(a.warmup, 3), budget_s=120.0)
    sample = cpu_sample_text(kind, done, threads, a)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": done,
        "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 tree + f32 TD", "data": "synthetic",
        "learner_updates_per_sec": value / a.batch,
        "config": {"workload": workload_name(a), "batch": a.batch, "replay_rows": a.replay_rows,
                   "optimizer": "Adam", "device": "host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v == "Active":
                    reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------- our arm
def build_learner(a, device, seed, world=1):
    import torch as t
    from machin_b200.frame.algorithms import DQNApex, DQNPer
    from machin_b200.frame.buffers import PrioritizedBuffer

    class Q(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1 = t.nn.Linear(STATE_DIM, HIDDEN)
            self.fc2 = t.nn.Linear(HIDDEN, HIDDEN)
            self.fc3 = t.nn.Linear(HIDDEN, N_ACTIONS)

        def forward(self, state):
            return self.fc3(t.relu(self.fc2(t.relu(self.fc1(state)))))

    from machin_b200.model import FusedAdam, FusedMLP
    if a.net != "torch":
        Q = lambda: FusedMLP([STATE_DIM, HIDDEN, HIDDEN, N_ACTIONS])  # noqa: E731  same layers, fused kernels
    t.manual_seed(seed)
    qnet, qtgt = Q().to(device), Q().to(device)
    if world > 1:
        # every process = one replay shard + one learner (APEX with the nets data-parallel over the
        # learner group): global PER draw across shards, gradient all-reduce, weight broadcast per update
        from machin_b200.parallel import ShardGroup
        from machin_b200.parallel.server import PushPullModelServer
        group = ShardGroup("apex", device=device)
        t.manual_seed(0)   # identical initial weights on every learner
        qnet, qtgt = Q().to(device), Q().to(device)
        learner = DQNApex(qnet, qtgt, FusedAdam if a.adam == "fused" else t.optim.Adam, t.nn.MSELoss(reduction="none"), group,
                          (PushPullModelServer("qnet", group, 0, lockstep=True),), batch_size=a.batch, learning_rate=1e-3,
                          discount=0.99, update_rate=0.005, replay_size=a.replay_rows, learner_group=group)
        buf = learner.replay_buffer
        learner.batched_draw = a.batched_draw
        learner.parallel_forwards = a.parallel_forwards
        learner.overlap_priority_update = a.overlap_update
        learner.fuse_optimizer_and_target = a.fuse_optimizer and a.adam == "fused"
        learner.use_fused_program = a.net in ("fused-program", "fused-rows")     # honoured by the fixed-shape exchange
        learner.fused_kind = "rows" if a.net == "fused-rows" else "program"
    else:
        buf = PrioritizedBuffer(buffer_size=a.replay_rows, buffer_device=device)
        import functools
        adam = FusedAdam if a.adam == "fused" else functools.partial(t.optim.Adam, fused=True)
        learner = DQNPer(qnet, qtgt, adam, t.nn.MSELoss(reduction="none"), batch_size=a.batch,
                         learning_rate=1e-3, discount=0.99, update_rate=0.005, replay_buffer=buf)
        learner.use_fused_program = a.net in ("fused-program", "fused-rows")
        learner.fused_kind = "rows" if a.net == "fused-rows" else "program"
        learner.fuse_online_forwards = a.fuse_forwards
        learner.overlap_priority_update = a.overlap_update and world == 1
        learner.parallel_forwards = a.parallel_forwards and world == 1
        learner.fuse_optimizer_and_target = a.fuse_optimizer and a.adam == "fused"
    gen = t.Generator(device=device).manual_seed(seed)
    chunk = 1_000_000
    done = 0
    while done < a.replay_rows:  # fill the ring: synthetic transitions + Exp(1) priorities
        n = min(chunk, a.replay_rows - done)
        buf.store_episode_tensors(
            {"state": {"state": t.randn(n, STATE_DIM, device=device, generator=gen)},
             "action": {"action": t.randint(0, N_ACTIONS, (n, 1), device=device, generator=gen)},
             "next_state": {"state": t.randn(n, STATE_DIM, device=device, generator=gen)}},
            {"reward": t.randn(n, 1, device=device, generator=gen),
             "terminal": t.rand(n, 1, device=device, generator=gen) < 0.01},
            priorities=t.empty(n, device=device, dtype=t.float64).exponential_(1.0, generator=gen))
        done += n
    buf.wt_tree.check()
    t.cuda.synchronize(device)
    return learner


KERNELS = {  # C entry point -> (label, algorithmic bytes per launch as f(batch, n_params))
    "mb_per_sample": ("per_sample (stratify + descent + IS weight)", lambda B, P, D: (8 * (D - 1) + 32) * B),
    "mb_rows_gather": ("rows_gather (batch assembly)", lambda B, P, D: 2 * ROW_BYTES * B),
    "mb_td_dqn": ("td_dqn (target + |delta| + loss + grad)", lambda B, P, D: (3 * N_ACTIONS * 4 + 8 + 4 + 1 + 8 + 12 + N_ACTIONS * 4) * B),
    "mb_tree_update_priority": ("tree_update_part (normalise + one CTA per subtree, one launch)", lambda B, P, D: (8 + 4 + 8 + 24 * (D - 1)) * B),
    "mb_soft_update": ("soft_update", lambda B, P, D: 12 * P),
    # fused value math (--net fused-rows / fused-program): what must cross HBM at least once -- both input batches,
    # both nets' parameters, the saved activations and output gradients (written by part 0, read by part 1), the
    # parameter gradients.  The kernels themselves are FFMA / L2-latency bound, not HBM bound.
    "mb_mlp_program_part0": ("fused forward x3 + TD + backward-input",
                             lambda B, P, D: 2 * B * STATE_DIM * 4 + 2 * P * 4 + 2 * B * (2 * HIDDEN + N_ACTIONS) * 4 + 40 * B),
    "mb_mlp_program_part1": ("fused dW / db", lambda B, P, D: B * (STATE_DIM + 4 * HIDDEN + N_ACTIONS) * 4 + P * 4),
}


PER_STEP = {}    # launches per update of a timed kernel where it is not 1 (filled by time_kernels)


def time_kernels(learner, table, reps=20, per_graph=16, capture_mode=None):
    """Average device time per launch of each libmachin_b200 kernel of the step.

    Eager launches are CPU-launch-bound (the kernels take microseconds), so each kernel is captured
    ``per_graph`` times into a CUDA graph -- with a different, really sampled index set per launch -- and
    the graph is replayed ``reps`` times between CUDA events on the launching stream."""
    import torch as t
    from machin_b200 import _lib, ops
    from machin_b200._lib import lib

    buf = learner.replay_buffer
    dev, B = buf.buffer_device, learner.batch_size
    tree = buf.wt_tree.handle
    n_stored = len(buf.storage)
    idx = [t.empty(B, dtype=t.int64, device=dev) for _ in range(per_graph)]
    isw = [t.empty(B, dtype=t.float64, device=dev) for _ in range(per_graph)]
    err = [t.rand(B, device=dev) for _ in range(per_graph)]
    q = [t.randn(B, N_ACTIONS, device=dev) for _ in range(3)]
    sp = lambda: _lib.stream_ptr()

    def k_sample(i):
        _lib.check(lib.mb_per_sample(tree, _lib.ptr(table[i % table.shape[0]]), B, n_stored, 0.0, 0.4, None, _lib.ptr(idx[i]),
                                     _lib.ptr(isw[i]), sp()))

    for i in range(per_graph):
        k_sample(i)
    fields = buf.storage.gather(idx[0])
    names = list(buf.storage.fields)
    arr = buf.storage._field_array(names, fields)

    def k_gather(i):
        _lib.check(lib.mb_rows_gather(_lib.ptr(buf.storage.data), buf.storage.pitch, buf.storage.max_size,
                                      _lib.ptr(idx[i]), B, arr, len(names), sp()))

    def k_td(i):
        ops.td_dqn(q[0], q[1], q[2], fields["action/action"], fields["reward"], fields["terminal"], 0.99,
                   is_weight=isw[i], want_grad=True)

    def k_update(i):
        _lib.check(lib.mb_tree_update_priority(tree, _lib.ptr(idx[i]), _lib.ptr(err[i]), _lib.MB_F32, None, B,
                                               buf.epsilon, buf.alpha, sp()))

    def k_soft(i):
        ops.soft_update(learner.qnet_target, learner.qnet, 0.0)  # rate 0: same traffic, weights unchanged

    timed = [("mb_per_sample", k_sample), ("mb_rows_gather", k_gather), ("mb_td_dqn", k_td),
             ("mb_tree_update_priority", k_update), ("mb_soft_update", k_soft)]
    from machin_b200.model import FusedMLP
    if isinstance(learner.qnet, FusedMLP) and not getattr(learner, "_program", None):
        # FusedMLP Q-net: every layer of every pass is one csrc/mlp.cu GEMM launch of this library.  Per update: three
        # forwards (online s, online s', target s'), the backward of the online one (weight gradient per layer, input
        # gradient for all but the first layer).
        sizes = learner.qnet.sizes
        ws = learner.qnet._workspace(dev)
        for l in range(len(sizes) - 1):
            K, N = sizes[l], sizes[l + 1]
            x, w, b = t.randn(B, K, device=dev), t.randn(N, K, device=dev), t.randn(N, device=dev)
            y, dy, dx = t.empty(B, N, device=dev), t.randn(B, N, device=dev), t.empty(B, K, device=dev)
            dw, db = t.empty(N, K, device=dev), t.empty(N, device=dev)
            relu = int(l < len(sizes) - 2)

            def k_fwd(i, x=x, w=w, b=b, y=y, N=N, K=K, relu=relu):
                _lib.check(lib.mb_mlp_linear_forward(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), B, N, K, relu, *ws.args(), sp()))

            def k_dw(i, x=x, dy=dy, dw=dw, db=db, N=N, K=K):
                _lib.check(lib.mb_mlp_linear_backward_weight(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(dw), _lib.ptr(db), B, N, K, *ws.args(), sp()))

            def k_dx(i, x=x, w=w, dy=dy, dx=dx, N=N, K=K):
                _lib.check(lib.mb_mlp_linear_backward_input(_lib.ptr(dy), _lib.ptr(w), _lib.ptr(x), _lib.ptr(dx), B, N, K, *ws.args(), sp()))

            KERNELS[f"gemm_fwd_{l}"] = (f"gemm forward {B}x{N}x{K} (+bias{', ReLU' if relu else ''})", lambda B_, P_, D_, N=N, K=K: 4 * (B_ * K + N * K + N + B_ * N))
            KERNELS[f"gemm_dw_{l}"] = (f"gemm weight+bias gradient {N}x{K + 1}x{B}", lambda B_, P_, D_, N=N, K=K: 4 * (B_ * N + B_ * K + N * K + N))
            timed += [(f"gemm_fwd_{l}", k_fwd), (f"gemm_dw_{l}", k_dw)]
            PER_STEP[f"gemm_fwd_{l}"], PER_STEP[f"gemm_dw_{l}"] = 3.0, 1.0
            if l > 0:
                KERNELS[f"gemm_dx_{l}"] = (f"gemm input gradient {B}x{K}x{N} (ReLU mask)", lambda B_, P_, D_, N=N, K=K: 4 * (B_ * N + N * K + 2 * B_ * K))
                timed.append((f"gemm_dx_{l}", k_dx))
                PER_STEP[f"gemm_dx_{l}"] = 1.0
    prog = getattr(learner, "_program", None)
    if prog:                                   # the fused program replaces the stand-alone TD kernel in the step
        def k_part(part):
            return lambda i: _lib.check(lib.mb_mlp_program_run_part(prog["handle"], part, sp()))
        timed = [x for x in timed if x[0] != "mb_td_dqn"] + [("mb_mlp_program_part0", k_part(1))]
        if learner.fused_kind == "rows":       # (the cooperative program does everything in part 1 of the C call)
            timed.append(("mb_mlp_program_part1", k_part(2)))
    out = {}
    for name, fn in timed:
        side = t.cuda.Stream(device=dev)
        side.wait_stream(t.cuda.current_stream(dev))
        with t.cuda.stream(side):
            fn(0)
        t.cuda.current_stream(dev).wait_stream(side)
        t.cuda.synchronize(dev)
        g = t.cuda.CUDAGraph()
        with (t.cuda.graph(g, capture_error_mode=capture_mode) if capture_mode else t.cuda.graph(g)):
            for i in range(per_graph):
                fn(i)
        g.replay()
        t.cuda.synchronize(dev)
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        t.cuda.synchronize(dev)
        out[name] = {"avg_us": e0.elapsed_time(e1) * 1e3 / (reps * per_graph), "launches_timed": reps * per_graph,
                     "per_step": PER_STEP.get(name, 1.0)}
    return out


def cuda_device(index):
    import torch as t
    return t.device("cuda", index)


def run_b200(a):
    import torch as t
    import torch.distributed as dist
    from machin_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries ONE JSON line: keep a private handle to it and point fd 1 at stderr, so that nothing a
    # library writes to the C-level stdout (NCCL's version banner showed up there in one 2-GPU run) can precede it
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not t.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    t.cuda.set_device(local)
    device = cuda_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=device)
        import threading

        def watchdog():
            sys.stderr.write(f"bench.py: rank {rank} still running after {a.watchdog_s:.0f} s -- giving up (exit 5)\n")
            sys.stderr.flush()
            os._exit(5)
        wd = threading.Timer(a.watchdog_s, watchdog)
        wd.daemon = True
        wd.start()
    learner = build_learner(a, device, seed=rank, world=world)
    graph = not a.no_graph
    if graph and world > 1:
        # N > 1 default: fixed-shape exchange, NCCL collectives captured too -- one graph per update (measured at
        # N = 2: 0.373 ms/update against 0.703 ms for the dynamic path; equivalent to it draw for draw on a group
        # of one, tests/test_gpu_algorithms.py).  --dynamic-exchange: two compute graphs around eager, exactly
        # sized NCCL exchanges (measured at N = 2, 4).
        try:
            learner.enable_cuda_graph(static_exchange=a.static_exchange,
                                      device_apportion=a.device_apportion and a.static_exchange,
                                      peer_exchange=a.peer_exchange and a.static_exchange)
        except Exception as exc:   # noqa: BLE001 -- a capture problem is deterministic: every rank lands here together
            if not a.static_exchange:
                raise
            sys.stderr.write(f"bench.py: rank {rank}: fixed-shape exchange could not be captured ({exc!r}); "
                             "falling back to the dynamic exchange\n")
            t.cuda.synchronize(device)
            a.static_exchange = False
            a.skip_teardown = True      # a half-captured NCCL graph may linger: leave like the fixed-shape run does
            del learner
            t.cuda.empty_cache()
            learner = build_learner(a, device, seed=rank, world=world)
            learner.enable_cuda_graph()
    elif graph:
        learner.enable_cuda_graph()
    K, W, B = a.steps, max(a.warmup, 3), a.batch
    depth = learner.replay_buffer.wt_tree.depth
    n_params = sum(p.numel() for p in learner.qnet.parameters())

    def barrier():
        t.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        t.cuda.synchronize(device)

    np.random.seed(1234 + rank)
    table = t.from_numpy(np.random.uniform(size=(K + W, B))).to(device)  # resident in HBM before timing
    sampler = ClockSampler(local)
    # ---- value: device-resident inputs, no host round trip inside the timed region
    def dev_step(s):
        if world == 1:
            return learner.update_device(uniform=table[s])
        loss = learner.update_device()
        learner.push_weights()
        return loss

    for s in range(W):
        dev_step(s)
    barrier()
    sampler.start()
    launches0 = _lib.lib.mb_launch_count()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(K):
        loss = dev_step(W + s)
    e1.record()
    barrier()
    launches = _lib.lib.mb_launch_count() - launches0
    if graph:   # graph replays do not pass through the C entry points: add the captured kernels per replay
        # (N = 1 and the fixed-shape exchange: everything is captured; dynamic exchange: the draws, gathers and
        # tree updates around the two graphs were counted eagerly above)
        launches += K * int(getattr(learner, "graph_launches", 0))
    ms_total = e0.elapsed_time(e1)
    # ---- e2e: public API, host uniforms in, loss out, every step
    for _ in range(W):
        learner.update()
    barrier()
    t0 = time.perf_counter()
    f0, f1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(K):
        last = learner.update()
    f1.record()
    barrier()
    e2e_ms = max(f0.elapsed_time(f1), (time.perf_counter() - t0) * 1e3)
    # ---- per-kernel event timing inside the step
    kern = {}
    if rank == 0:
        try:                                   # (local kernels of this rank's shard; no collective inside)
            # (with a process group alive, NCCL's watchdog thread may touch CUDA while this thread captures)
            kern = time_kernels(learner, table, capture_mode="thread_local" if world > 1 else None)
        except Exception as exc:               # noqa: BLE001 -- at N > 1 the other ranks wait for this one: never raise here
            if world == 1:
                raise
            sys.stderr.write(f"bench.py: per-kernel timing skipped at N = {world}: {exc!r}\n")
    clocks = sampler.stop()
    tt = t.tensor([ms_total, e2e_ms], device=device, dtype=t.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)  # max over ranks
    ms_total, e2e_ms = float(tt[0]), float(tt[1])
    def leave():
        if world > 1:
            # NCCL collectives captured in a CUDA graph: destroy_process_group() blocks while such a graph is alive, and
            # returns once it has been released (tools/nccl_graph_repro.py ... release; profiles/r2_nccl_graph_release.log)
            t.cuda.synchronize(device)
            dist.barrier()
            if hasattr(learner, "release_graphs"):
                learner.release_graphs()
            t.cuda.synchronize(device)
            if getattr(a, "skip_teardown", False):      # a capture that failed half-way may still hold NCCL work
                sys.stdout.flush()
                sys.stderr.flush()
                dist.barrier()
                os._exit(0)
            dist.destroy_process_group()

    if rank != 0:
        leave()
        return
    value = world * B * K / (ms_total * 1e-3)
    e2e = world * B * K / (e2e_ms * 1e-3)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dom = max(kern, key=lambda k: kern[k]["avg_us"] * kern[k]["per_step"]) if kern else None
    roofline = None
    if dom:
        label, fbytes = KERNELS[dom]
        bytes_per_launch = fbytes(B, n_params, depth)
        achieved = bytes_per_launch / (kern[dom]["avg_us"] * 1e-6) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(dom)
        # `dom` = the kernel with the largest share of the update's device time (launches per update x time per launch);
        # the single slowest launch is named beside it
        slow = max(kern, key=lambda k: kern[k]["avg_us"])
        roofline = {"bound": "hbm", "kernel": label, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_us": kern[dom]["avg_us"],
                    "launches_per_step": kern[dom]["per_step"],
                    "slowest_single_launch": {"kernel": KERNELS[slow][0], "avg_launch_us": kern[slow]["avg_us"],
                                              "algorithmic_bytes_per_launch": KERNELS[slow][1](B, n_params, depth),
                                              "frac": KERNELS[slow][1](B, n_params, depth) / (kern[slow]["avg_us"] * 1e-6) / 1e9 / peak,
                                              "traffic": (json.load(open(tp)).get(slow) if os.path.exists(tp) else None)},
                    "note": "B=512 touches <1 MB per launch: the step is launch/latency-bound, not bandwidth-bound "
                            "(DESIGN.md section 5 reports the saturated-throughput figures per kernel)"}
    cpu = None
    if not a.no_cpu_baseline and world == 1:
        v, ms, threads, done, kind = cpu_learner_rate(a, a.cpu_steps, 5)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "ms_per_step": ms,
               "sample": cpu_sample_text(kind, done, threads, a)}
    # host <-> device bytes per update and rank, counted from the tensors that are copied: the uniforms (and, for the
    # fixed-shape exchange, the apportioned counts, the per-requester parameter blocks and the loss scale) go in; the
    # loss comes out -- and at N > 1 the all-gathered shard sums the host apportions with (prioritized_buffer_d.py:221-234)
    st = getattr(learner, "_static", None) if (world > 1 and a.static_exchange) else None
    if st is not None and not st["device_apportion"]:
        h2d_bytes = sum(st[k].numel() * st[k].element_size() for k in ("uni", "counts", "params", "scale"))
    elif st is not None:
        h2d_bytes = st["uni"].numel() * 8 + 8
    else:
        h2d_bytes = 8 * B
    d2h_bytes = 8 + (8 * world if (world > 1 and not (st is not None and st["device_apportion"])) else 0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 tree/IS-weights + f32 TD", "data": "synthetic",
        "learner_updates_per_sec": K / (ms_total * 1e-3),
        "config": {"workload": workload_name(a), "batch": B, "replay_rows": a.replay_rows, "optimizer": "Adam",
                   "parallelism": f"{world} learner(s), one replay shard each",
                   "launch": ("one CUDA graph per update" if (world == 1 or a.static_exchange) else
                              "two CUDA graphs per update around eager NCCL exchanges") if graph else "eager",
                   "net": a.net, "adam": a.adam,
                   "optimizer_and_target": "one launch (FusedAdam step + soft update)" if (a.fuse_optimizer and a.adam == "fused") else "two launches", "priority_update": "parallel graph branch" if (a.overlap_update and (world == 1 or a.static_exchange)) else "in line",
                   "step_options_off": os.environ.get("MB_STEP_OFF", "") or None,
                   "online_forwards": "one pass over [s; s']" if a.fuse_forwards else ("split, no-grad ones on parallel branches" if (a.parallel_forwards and (world == 1 or a.static_exchange)) else "split (as the reference)"),
                   "transport": None if world == 1 else (
                       "peer memory over NVLink: this library's kernels store into the other GPUs' buffers, one barrier kernel "
                       "per exchange (csrc/peer.cu)" if (getattr(learner, "_static", None) or {}).get("peer") is not None
                       else "NCCL (torch.distributed), collectives captured in the update's CUDA graph" if a.static_exchange
                       else "NCCL (torch.distributed), eager"),
                   "exchange": None if world == 1 else "per update: all-gather of shard sums, all-to-all of sampled rows, "
                               "all-to-all of priority write-backs, all-reduce of gradients (NCCL); no weight broadcast: every "
                               "member learns and applies the same averaged gradients (actor-only members: --workload c5)",
                   "l2": "inputs larger than L2: 268 MB tree + 6.9 GB ring, randomly addressed every step"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": e2e_ms / K, "last_loss": last},
        "gpu_launches": int(launches), "gpu_launches_per_step": launches / K,
        "kernels": {KERNELS[k][0]: v for k, v in kern.items()},
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
    }
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    leave()


if __name__ == "__main__":
    args = parse()
    if args.workload != "c2":
        import bench_workloads as bw
        if args.impl == "reference":
            bw.run_reference(args)
        else:
            {"c3": bw.run_c3, "c4": bw.run_c4, "c5": bw.run_c5}[args.workload](args, ClockSampler)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
