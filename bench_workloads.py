"""bench.py --workload c3 | c4 | c5: the other BASELINE.json configurations under the same JSON contract as C2.

  c3  IMPALA learner, 1024 episodes x T = 20, state 256, one GPU (BASELINE configs[2]).  One step = IMPALA.update():
      episode selection (Python's random, impala.py:58) -> row gather -> pi(a|s), V(s), V(s') -> V-trace kernel ->
      both losses, both backward passes, both Adam steps.  roofline = the V-trace kernel (37 B / step).
  c4  APEX-DDPG learner (state 376, action 17, batch 512) over a sharded prioritized buffer, 12.5 M rows per shard
      (50 M over 4 GPUs, BASELINE configs[3]); every process = one shard + one learner (weak scaling), NCCL row /
      priority exchange, gradient all-reduce.  roofline = the slowest of this library's kernels in the step.
  c5  IMPALA, N - 1 actor GPUs -> 1 learner GPU, 4096 episodes x T = 20 per batch (BASELINE configs[4]): the
      trajectory push is the sharded episode buffer's all-to-all (rows + episode lengths), the weight broadcast the
      versioned model server; `link` reports the bytes the learner receives per second against the NVLink peer-copy
      figure of B200_PROFILING.md.

The metric stays BASELINE's: transitions (steps) sampled and learned from per second, whole job.  `cpu_baseline` /
`--impl reference` time the oracle ports of the reference's CPU learners (oracle/impala_learner.py,
oracle/ddpg_learner.py) on the host cores for a bounded number of updates.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
METRIC, UNIT = "transitions_sampled_per_sec", "transitions/s"
C3 = dict(S=256, A=6, T=20, EP=1024, H=256)
C4 = dict(S=376, A=17, H=256, B=512, ROWS=12_500_000)
C5 = dict(S=256, A=6, T=20, EP=4096, H=256)
NVLINK_PEER_GBS = 770.0            # measured peer copy per direction (B200_PROFILING.md); nominal 900


def _peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def _graph_time(fn, per_graph=8, reps=20):
    """Average device time of one fn() launch: per_graph launches captured in a graph, replayed between CUDA events."""
    import torch as t
    side = t.cuda.Stream()
    side.wait_stream(t.cuda.current_stream())
    with t.cuda.stream(side):
        fn(0)
    t.cuda.current_stream().wait_stream(side)
    t.cuda.synchronize()
    g = t.cuda.CUDAGraph()
    with t.cuda.graph(g, capture_error_mode="thread_local"):
        for i in range(per_graph):
            fn(i)
    g.replay()
    t.cuda.synchronize()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    t.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * per_graph)


def _nets_impala(cfg):
    import torch as t
    S, A, H = cfg["S"], cfg["A"], cfg["H"]

    class Actor(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2, self.fc3 = t.nn.Linear(S, H), t.nn.Linear(H, H), t.nn.Linear(H, A)

        def forward(self, state, action=None):
            logp_all = t.log_softmax(self.fc3(t.relu(self.fc2(t.relu(self.fc1(state))))), dim=1)
            act = action if action is not None else t.multinomial(logp_all.exp(), 1)
            return act.view(-1, 1), logp_all.gather(1, act.view(-1, 1).long()), -(logp_all.exp() * logp_all).sum(1, keepdim=True)

    class Critic(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2, self.fc3 = t.nn.Linear(S, H), t.nn.Linear(H, H), t.nn.Linear(H, 1)

        def forward(self, state):
            return self.fc3(t.relu(self.fc2(t.relu(self.fc1(state)))))

    return Actor, Critic


def _store_episodes(buf, n_ep, cfg, dev, gen):
    """n_ep synthetic episodes of T steps, written 256 episodes per call (each call = one episode per T rows)."""
    import torch as t
    S, A, T = cfg["S"], cfg["A"], cfg["T"]
    term = t.zeros(T, 1, dtype=t.bool)
    term[-1] = True
    el = t.zeros(T, 1, dtype=t.int64)
    el[0] = T
    for _ in range(n_ep):
        buf.store_episode_tensors(
            {"state": {"state": t.randn(T, S, device=dev, generator=gen)},
             "action": {"action": t.randint(0, A, (T, 1), device=dev, generator=gen)},
             "next_state": {"state": t.randn(T, S, device=dev, generator=gen)}},
            {"reward": t.randn(T, 1, device=dev, generator=gen), "terminal": term},
            {"action_log_prob": -t.rand(T, 1, device=dev, generator=gen), "episode_length": el})


# =========================================================================================== CPU arms
def cpu_impala_rate(cfg, budget_s, max_updates):
    """oracle/impala_learner.py on the host cores -> (steps/s, ms/update, threads, updates, sample text)."""
    import torch as t
    from oracle.impala_learner import CpuImpalaLearner
    import random
    random.seed(0)
    learner = CpuImpalaLearner(cfg["EP"], cfg["T"], cfg["S"], cfg["A"], cfg["EP"], hidden=cfg["H"])
    ncpu = os.cpu_count() or 1
    best, best_dt = 1, float("inf")
    for th in [c for c in (1, 4, 16, 32) if c <= ncpu]:
        t.set_num_threads(th)
        t0 = time.perf_counter()
        learner.step(vtrace_episode_limit=8)
        dt = time.perf_counter() - t0
        if dt < best_dt:
            best, best_dt = th, dt
    t.set_num_threads(best)
    t0, done = time.perf_counter(), 0
    while done < max_updates and (done == 0 or time.perf_counter() - t0 < budget_s):
        learner.step()
        done += 1
    dt = time.perf_counter() - t0
    steps = cfg["EP"] * cfg["T"]
    sample = (f"{done} full IMPALA updates of oracle/impala_learner.py ({cfg['EP']} episodes x T={cfg['T']}: Python-object episode "
              f"sampling + cat, CPU torch actor/critic, the reference's per-step V-trace loop, Adam); best of thread counts "
              f"1/4/16/32 -> {best} threads, {ncpu} cores present")
    return steps * done / dt, dt / done * 1e3, best, done, sample


def cpu_ddpg_rate(cfg, rows, budget_s, max_updates, storage_rows=50_000):
    import torch as t
    from oracle.ddpg_learner import CpuDdpgPerLearner
    np.random.seed(0)
    learner = CpuDdpgPerLearner(rows, storage_rows, cfg["S"], cfg["A"], cfg["B"], hidden=cfg["H"])
    ncpu = os.cpu_count() or 1
    best, best_dt = 1, float("inf")
    for th in [c for c in (1, 4, 8, 16, 32) if c <= ncpu]:
        t.set_num_threads(th)
        learner.step()
        t0 = time.perf_counter()
        for _ in range(3):
            learner.step()
        dt = time.perf_counter() - t0
        if dt < best_dt:
            best, best_dt = th, dt
    t.set_num_threads(best)
    t0, done = time.perf_counter(), 0
    while done < max_updates and time.perf_counter() - t0 < budget_s:
        learner.step()
        done += 1
    dt = time.perf_counter() - t0
    sample = (f"{done} full DDPG-PER updates of oracle/ddpg_learner.py (NumPy float64 tree with {rows} leaves, Python-object "
              f"storage capped at {storage_rows} rows, CPU torch actor/critic); best of thread counts 1..32 -> {best} threads, "
              f"{ncpu} cores present")
    return cfg["B"] * done / dt, dt / done * 1e3, best, done, sample


def run_reference(a):
    rank, world, _ = _env()
    if rank != 0:
        return
    if a.workload in ("c3", "c5"):
        cfg = C3 if a.workload == "c3" else C5
        value, ms, threads, done, sample = cpu_impala_rate(cfg, budget_s=90.0, max_updates=max(1, min(a.steps, 12)))
        name = _c3_name() if a.workload == "c3" else _c5_name(a.gpus)
        dtype = "f32"
    else:
        rows = a.replay_rows if a.replay_rows != 10_000_000 else C4["ROWS"]
        value, ms, threads, done, sample = cpu_ddpg_rate(C4, rows, budget_s=90.0, max_updates=a.steps, storage_rows=a.cpu_storage_rows)
        name, dtype = _c4_name(rows, 1), "f64 tree/IS-weights + f32 TD"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": done,
            "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "learner_updates_per_sec": 1e3 / ms,
            "config": {"workload": name, "device": "host CPU"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def _c3_name():
    return f"C3 IMPALA learner: {C3['EP']} episodes x T={C3['T']} ({C3['EP'] * C3['T']} steps), state {C3['S']}, {C3['A']} actions, V-trace"


def _c4_name(rows, world):
    return (f"C4 APEX-DDPG learner: PER {rows} rows per shard x {world} shard(s) (state {C4['S']}, act {C4['A']}), "
            f"batch {C4['B']}, actor/critic MLPs {C4['H']}-{C4['H']}")


def _c5_name(world):
    return (f"C5 IMPALA: {max(world - 1, 1)} actor GPU(s) -> 1 learner GPU, {C5['EP']} episodes x T={C5['T']} per batch, "
            f"state {C5['S']}, NCCL trajectory push + weight broadcast")


# =========================================================================================== shared GPU plumbing
class _Run:
    def __init__(self, a):
        import torch as t
        import torch.distributed as dist
        self.t, self.dist = t, dist
        self.rank, self.world, self.local = _env()
        sys.stdout.flush()
        self.json_fd = os.dup(1)
        os.dup2(2, 1)                       # nothing a library prints may precede the one JSON line
        if not t.cuda.is_available():
            raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        t.cuda.set_device(self.local)
        self.dev = t.device("cuda", self.local)
        if self.world > 1:
            if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", device_id=self.dev)
            import threading

            def watchdog():
                sys.stderr.write(f"bench.py: rank {self.rank} still running after {a.watchdog_s:.0f} s -- giving up (exit 5)\n")
                sys.stderr.flush()
                os._exit(5)
            wd = threading.Timer(a.watchdog_s, watchdog)
            wd.daemon = True
            wd.start()

    def barrier(self):
        self.t.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.t.cuda.synchronize(self.dev)

    def max_over_ranks(self, values):
        tt = self.t.tensor(values, device=self.dev, dtype=self.t.float64)
        if self.world > 1:
            self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in tt]

    def emit(self, line):
        if self.rank == 0:
            os.write(self.json_fd, (json.dumps(line) + "\n").encode())

    def leave(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def _timed(run, K, step, sampler=None):
    """K steps between a barrier + synchronize on both sides -> (device ms, wall ms) of this rank."""
    t = run.t
    run.barrier()
    if sampler is not None:
        sampler.start()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out = None
    for s in range(K):
        out = step(s)
    e1.record()
    run.barrier()
    return e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3, out


# =========================================================================================== C3
def run_c3(a, ClockSampler):
    run = _Run(a)
    t, dev = run.t, run.dev
    from machin_b200 import _lib, ops
    from machin_b200.frame.algorithms import IMPALA
    from machin_b200.model import FusedAdam
    from machin_b200.parallel import LocalGroup
    if run.world != 1:
        raise SystemExit("--workload c3 is the single-GPU IMPALA learner; use c5 for actors -> learner over NCCL")
    cfg = C3
    S, A, T, EP = cfg["S"], cfg["A"], cfg["T"], cfg["EP"]
    Actor, Critic = _nets_impala(cfg)
    t.manual_seed(0)
    import functools
    adam = FusedAdam if a.adam == "fused" else functools.partial(t.optim.Adam, fused=True)
    imp = IMPALA(Actor().to(dev), Critic().to(dev), adam, t.nn.MSELoss(reduction="sum"), LocalGroup(device=dev), (),
                 batch_size=EP, replay_size=4 * EP * T)
    gen = t.Generator(device=dev).manual_seed(0)
    _store_episodes(imp.replay_buffer, 2 * EP, cfg, dev, gen)       # 2048 stored episodes, 1024 sampled per update
    graph = not a.no_graph
    if graph:
        imp.enable_cuda_graph(episode_length=T)
    K, W = a.steps, max(a.warmup, 3)
    import random
    random.seed(1234)
    sampler = ClockSampler(run.local)
    # value: update_device() -- the host picks the episodes (Python's random, as impala.py:58) and copies their slot
    # list in; the losses stay on the device
    for s in range(W):
        imp.update_device()
    launches0 = _lib.lib.mb_launch_count()
    ms_dev, ms_wall, _ = _timed(run, K, lambda s: imp.update_device(), sampler)
    launches = _lib.lib.mb_launch_count() - launches0 + (K * int(getattr(imp, "graph_launches", 0)) if graph else 0)
    ms_total = max(ms_dev, ms_wall) if not graph else ms_dev
    # e2e: the public update(): the same + both losses read back every step
    for _ in range(W):
        imp.update()
    f_dev, f_wall, last = _timed(run, K, lambda s: imp.update())
    e2e_ms = max(f_dev, f_wall)
    # V-trace kernel alone, at the learner batch and saturated
    n = EP * T

    def vt_inputs(steps, seed):
        g = t.Generator(device=dev).manual_seed(seed)
        v, nv, r = (t.randn(steps, 1, device=dev, generator=g) for _ in range(3))
        lp, lm = -t.rand(steps, 1, device=dev, generator=g), -t.rand(steps, 1, device=dev, generator=g)
        te = t.zeros(steps, 1, dtype=t.bool, device=dev)
        te[T - 1::T] = True
        return v, nv, r, te, lp, lm
    sets = [vt_inputs(n, i) for i in range(8)]
    us_batch = _graph_time(lambda i: ops.vtrace(*sets[i % 8], 0.99, n_episodes=EP, max_len=T))
    big_ep = 1 << 21                                    # 42 M steps: 1.55 GB of traffic per launch, beyond L2
    big = vt_inputs(big_ep * T, 99)
    us_big = _graph_time(lambda i: ops.vtrace(*big, 0.99, n_episodes=big_ep, max_len=T), per_graph=2, reps=5)
    clocks = sampler.stop()
    peak, peak_src = _peak()
    alg = 37 * n
    roofline = {"bound": "hbm", "kernel": "vtrace_kernel (fused reverse scan, shared-memory staged)", "achieved": alg / (us_batch * 1e-6) / 1e9,
                "peak": peak, "unit": "GB/s", "frac": alg / (us_batch * 1e-6) / 1e9 / peak, "traffic": _traffic("mb_vtrace"),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "avg_launch_us": us_batch,
                "saturated": {"steps_per_launch": big_ep * T, "avg_launch_us": us_big,
                              "achieved": 37 * big_ep * T / (us_big * 1e-6) / 1e9,
                              "frac": 37 * big_ep * T / (us_big * 1e-6) / 1e9 / peak},
                "note": "20 480 steps move 0.76 MB: one launch is latency-bound; `saturated` is the same kernel on 42 M steps"}
    cpu = None
    if not a.no_cpu_baseline:
        v, ms, threads, done, sample = cpu_impala_rate(cfg, budget_s=25.0, max_updates=6)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_step": ms, "sample": sample}
    line = {"metric": METRIC, "value": n * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "learner_updates_per_sec": K / (ms_total * 1e-3),
            "config": {"workload": _c3_name(), "episodes_per_update": EP, "episode_length": T, "stored_episodes": 2 * EP,
                       "optimizer": "Adam", "launch": "one CUDA graph per update (gather -> forwards -> V-trace -> backward -> Adam)"
                       if graph else "eager", "adam": a.adam,
                       "l2": "the batch is re-gathered from the ring by fresh random episode picks every step; the ring "
                             "(2048 episodes x 20 x 2.1 KB = 86 MB) is smaller than L2 -- stated, not hidden"},
            "e2e": {"value": n * K / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8,
                    "ms_per_step": e2e_ms / K, "last": [float(x) for x in last]},
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / K,
            "kernels": {"vtrace_kernel": {"avg_us": us_batch, "per_step": 1.0}},
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks}
    run.emit(line)
    run.leave()


def _traffic(key):
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        return json.load(open(tp)).get(key)
    return None


# =========================================================================================== C4
def run_c4(a, ClockSampler):
    run = _Run(a)
    t, dev, world, rank = run.t, run.dev, run.world, run.rank
    from machin_b200 import _lib, ops
    from machin_b200.frame.algorithms import DDPGApex, DDPGPer
    from machin_b200.model import FusedAdam
    from machin_b200.parallel import LocalGroup, ShardGroup
    cfg = C4
    S, A, H, B = cfg["S"], cfg["A"], cfg["H"], a.batch
    rows = a.replay_rows if a.replay_rows != 10_000_000 else cfg["ROWS"]

    class Actor(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2, self.fc3 = t.nn.Linear(S, H), t.nn.Linear(H, H), t.nn.Linear(H, A)

        def forward(self, state):
            return t.tanh(self.fc3(t.relu(self.fc2(t.relu(self.fc1(state))))))

    class Critic(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2, self.fc3 = t.nn.Linear(S + A, H), t.nn.Linear(H, H), t.nn.Linear(H, 1)

        def forward(self, state, action):
            return self.fc3(t.relu(self.fc2(t.relu(self.fc1(t.cat([state, action], dim=1))))))

    if a.net == "fused":                                # the same layers as one csrc/mlp.cu kernel per layer and direction
        from machin_b200.model import FusedMLP

        class Actor(t.nn.Module):                       # noqa: F811
            def __init__(self):
                super().__init__()
                self.net = FusedMLP([S, H, H, A])

            def forward(self, state):
                return t.tanh(self.net(state))

        Critic = lambda: FusedMLP([S + A, H, H, 1])     # noqa: E731  forward(state, action) concatenates
    t.manual_seed(0)                                    # identical initial weights on every learner
    nets = [Actor().to(dev), Actor().to(dev), Critic().to(dev), Critic().to(dev)]
    graph = not a.no_graph
    if world > 1:
        group = ShardGroup("apex", device=dev)
        learner = DDPGApex(*nets, FusedAdam, t.nn.MSELoss(reduction="none"), group, (), batch_size=B, replay_size=rows,
                           learner_group=group)
        learner.batched_draw = learner.overlap_priority_update = not a.serial
    else:
        learner = DDPGPer(*nets, FusedAdam, t.nn.MSELoss(reduction="none"), batch_size=B, replay_size=rows)
    buf = learner.replay_buffer
    gen = t.Generator(device=dev).manual_seed(rank)
    done = 0
    while done < rows:
        n = min(500_000, rows - done)
        buf.store_episode_tensors(
            {"state": {"state": t.randn(n, S, device=dev, generator=gen)},
             "action": {"action": t.rand(n, A, device=dev, generator=gen) * 2 - 1},
             "next_state": {"state": t.randn(n, S, device=dev, generator=gen)}},
            {"reward": t.randn(n, 1, device=dev, generator=gen), "terminal": t.rand(n, 1, device=dev, generator=gen) < 0.01},
            priorities=t.empty(n, device=dev, dtype=t.float64).exponential_(1.0, generator=gen))
        done += n
    buf.wt_tree.check()
    if graph and world > 1:
        learner.enable_cuda_graph(device_apportion=not a.host_apportion, peer_exchange=not a.nccl_exchange)
    elif graph:
        learner.enable_cuda_graph()
    K, W = a.steps, max(a.warmup, 3)
    np.random.seed(1234 + rank)
    sampler = ClockSampler(run.local)
    for _ in range(W):
        learner.update_device()
    launches0 = _lib.lib.mb_launch_count()
    ms_dev, ms_wall, _ = _timed(run, K, lambda s: learner.update_device(), sampler)
    launches = _lib.lib.mb_launch_count() - launches0 + (K * int(getattr(learner, "graph_launches", 0)) if graph else 0)
    for _ in range(W):
        learner.update()
    f_dev, f_wall, last = _timed(run, K, lambda s: learner.update())
    # at N > 1 the exchange reads the shard sums on the host every update: wall time is the honest clock there
    ms_total, e2e_ms = run.max_over_ranks([ms_dev if graph else max(ms_dev, ms_wall), max(f_dev, f_wall)])
    if hasattr(learner, "release_graphs"):
        t.cuda.synchronize(dev)
        learner.release_graphs()
    # this library's kernels on this rank's shard
    kern = {}
    if rank == 0:
        tree, n_stored, depth = buf.wt_tree.handle, len(buf.storage), buf.wt_tree.depth
        idx = [t.empty(B, dtype=t.int64, device=dev) for _ in range(8)]
        isw = [t.empty(B, dtype=t.float64, device=dev) for _ in range(8)]
        uni = t.rand(8, B, dtype=t.float64, device=dev)
        sp = _lib.stream_ptr

        def k_sample(i):
            _lib.check(_lib.lib.mb_per_sample(tree, _lib.ptr(uni[i % 8]), B, n_stored, 0.0, 0.4, None, _lib.ptr(idx[i % 8]),
                                              _lib.ptr(isw[i % 8]), sp()))
        for i in range(8):
            k_sample(i)
        pitch = buf.storage.pitch
        payload = sum(f.nbytes for f in buf.storage.fields.values()) if hasattr(next(iter(buf.storage.fields.values())), "nbytes") else pitch
        err = [t.rand(B, device=dev) for _ in range(8)]
        nv, rw, te = t.randn(B, 1, device=dev), t.randn(B, 1, device=dev), t.zeros(B, 1, dtype=t.bool, device=dev)
        kern = {
            "per_sample (stratify + descent + IS weight)": (_graph_time(k_sample), (8 * (depth - 1) + 32) * B),
            "rows_gather (batch assembly, %d B rows)" % pitch: (_graph_time(lambda i: buf.storage.gather(idx[i % 8])), 2 * payload * B),
            "td_target (y, |delta|)": (_graph_time(lambda i: ops.td_target(nv, rw, te, 0.99)), 17 * B),
            "tree_update_part (priority write-back)": (_graph_time(lambda i: _lib.check(_lib.lib.mb_tree_update_priority(
                tree, _lib.ptr(idx[i % 8]), _lib.ptr(err[i % 8]), _lib.MB_F32, None, B, buf.epsilon, buf.alpha, sp()))),
                (20 + 24 * (depth - 1)) * B),
        }
    clocks = sampler.stop()
    peak, peak_src = _peak()
    roofline = None
    if kern:
        dom = max(kern, key=lambda k: kern[k][0])
        us, alg = kern[dom]
        roofline = {"bound": "hbm", "kernel": dom, "achieved": alg / (us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": alg / (us * 1e-6) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg, "avg_launch_us": us,
                    "note": "B=512: every launch moves < 4 MB -- latency-bound; saturated figures in DESIGN.md section 2"}
    cpu = None
    if not a.no_cpu_baseline and rank == 0 and world == 1:
        v, ms, threads, done_u, sample = cpu_ddpg_rate(cfg, rows, budget_s=25.0, max_updates=a.cpu_steps, storage_rows=a.cpu_storage_rows)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_step": ms, "sample": sample}
    h2d = 8 * B * (world if world > 1 else 1)
    line = {"metric": METRIC, "value": world * B * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 tree/IS-weights + f32 TD", "data": "synthetic", "learner_updates_per_sec": K / (ms_total * 1e-3),
            "config": {"workload": _c4_name(rows, world), "batch": B, "rows_per_shard": rows, "row_pitch_bytes": int(buf.storage.pitch),
                       "shard_bytes": int(buf.storage.pitch) * rows, "optimizer": "Adam (FusedAdam)", "net": a.net,
                       "parallelism": f"{world} learner(s), one replay shard each",
                       "launch": "one CUDA graph per update" if graph else "eager kernels around eager NCCL exchanges",
                       "transport": None if world == 1 else (
                           "peer memory over NVLink (csrc/peer.cu): push kernels + one barrier kernel per exchange"
                           if (getattr(learner, "_static", None) or {}).get("peer") is not None else "NCCL (torch.distributed)"),
                       "exchange": None if world == 1 else "per update: shard sums, sampled rows, priority write-backs, "
                                   "2 gradient averages (critic, actor)",
                       "l2": "inputs larger than L2: 268 MB tree + 38.8 GB ring per shard, randomly addressed every step"},
            "e2e": {"value": world * B * K / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 16 + (8 * world if world > 1 else 0), "ms_per_step": e2e_ms / K,
                    "last": [float(x) for x in last]},
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / K,
            "kernels": {k: {"avg_us": v[0], "algorithmic_bytes": v[1]} for k, v in kern.items()},
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks}
    run.emit(line)
    run.leave()


# =========================================================================================== C5
def run_c5(a, ClockSampler):
    run = _Run(a)
    t, dev, world, rank = run.t, run.dev, run.world, run.rank
    from machin_b200 import _lib, ops
    from machin_b200.frame.algorithms import IMPALA
    from machin_b200.model import FusedAdam
    from machin_b200.parallel import LocalGroup, ShardGroup
    from machin_b200.parallel.server import PushPullModelServer
    cfg = C5
    S, A, T, EP = cfg["S"], cfg["A"], cfg["T"], cfg["EP"]
    Actor, Critic = _nets_impala(cfg)
    t.manual_seed(0)
    group = ShardGroup("impala", device=dev) if world > 1 else LocalGroup(device=dev)
    server = (PushPullModelServer("actor", group, 0, lockstep=True),) if world > 1 else ()
    stored = 2048                                      # episodes held by every actor
    imp = IMPALA(Actor().to(dev), Critic().to(dev), FusedAdam, t.nn.MSELoss(reduction="sum"), group, server,
                 batch_size=EP, replay_size=2 * max(stored, EP) * T, requesters=[0] if world > 1 else None)
    gen = t.Generator(device=dev).manual_seed(rank)
    if world == 1:
        _store_episodes(imp.replay_buffer, 2 * EP, cfg, dev, gen)
    elif rank > 0:                                      # actors hold the trajectories; the learner stores nothing
        _store_episodes(imp.replay_buffer, stored, cfg, dev, gen)
    run.barrier()
    K, W = a.steps, max(a.warmup, 3)
    import random
    random.seed(1234 + rank)
    sampler = ClockSampler(run.local)
    for _ in range(W):
        imp.update_device()
    launches0 = _lib.lib.mb_launch_count()
    ms_dev, ms_wall, _ = _timed(run, K, lambda s: imp.update_device(), sampler)
    launches = _lib.lib.mb_launch_count() - launches0
    for _ in range(W):
        imp.update()
    f_dev, f_wall, last = _timed(run, K, lambda s: imp.update())
    ms_total, e2e_ms = run.max_over_ranks([max(ms_dev, ms_wall), max(f_dev, f_wall)])
    clocks = sampler.stop()
    actors = max(world - 1, 1)
    per_member = -(-EP // world) if world > 1 else EP   # buffer_d.py:143-144: ceil(batch / members) episodes per member
    n_steps = (actors * per_member * T) if world > 1 else EP * T
    pitch = int(imp.replay_buffer.storage.pitch) if imp.replay_buffer.storage.data is not None else None
    pitch = run.max_over_ranks([float(pitch or 0)])[0]
    wire_bytes = n_steps * (pitch + 8)                  # packed rows + the episode lengths that ride along
    n_params = sum(p.numel() for p in imp.actor.parameters())
    peak, peak_src = _peak()
    us_vt = None
    if rank == 0:
        g = t.Generator(device=dev).manual_seed(5)
        v, nv, r = (t.randn(n_steps, 1, device=dev, generator=g) for _ in range(3))
        lp, lm = -t.rand(n_steps, 1, device=dev, generator=g), -t.rand(n_steps, 1, device=dev, generator=g)
        te = t.zeros(n_steps, 1, dtype=t.bool, device=dev)
        te[T - 1::T] = True
        us_vt = _graph_time(lambda i: ops.vtrace(v, nv, r, te, lp, lm, 0.99, n_episodes=n_steps // T, max_len=T))
    alg = 37 * n_steps
    roofline = None if us_vt is None else {
        "bound": "hbm", "kernel": "vtrace_kernel (fused reverse scan, shared-memory staged)", "achieved": alg / (us_vt * 1e-6) / 1e9,
        "peak": peak, "unit": "GB/s", "frac": alg / (us_vt * 1e-6) / 1e9 / peak, "traffic": _traffic("mb_vtrace"),
        "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "avg_launch_us": us_vt}
    line = {"metric": METRIC, "value": n_steps * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "learner_updates_per_sec": K / (ms_total * 1e-3),
            "config": {"workload": _c5_name(world), "episodes_per_update": actors * per_member if world > 1 else EP,
                       "episodes_per_member": per_member, "episode_length": T, "row_pitch_bytes": pitch,
                       "parallelism": f"1 learner + {world - 1} actor-only member(s)" if world > 1 else "1 GPU (learner holds the episodes)",
                       "launch": "eager (variable-shape trajectory exchange)",
                       "exchange": None if world == 1 else "per update: all-to-all of trajectory rows + episode lengths (actors -> learner), "
                                   "broadcast of the actor's weights with version (learner -> actors)",
                       "l2": "168 MB of trajectory rows per update cross NVLink into fresh buffers (larger than L2)"},
            "link": None if world == 1 else {"bytes_into_learner_per_update": wire_bytes, "weight_bytes_broadcast_per_update": 4 * n_params + 16,
                                             "achieved_GBps": wire_bytes / (ms_total / K * 1e-3) / 1e9,
                                             "nvlink_peer_copy_GBps": NVLINK_PEER_GBS,
                                             "frac_of_peer_copy": wire_bytes / (ms_total / K * 1e-3) / 1e9 / NVLINK_PEER_GBS,
                                             "note": "whole-update time in the denominator (exchange + learner compute are not overlapped)"},
            "e2e": {"value": n_steps * K / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * per_member * T,
                    "d2h_bytes_per_step": 8, "ms_per_step": e2e_ms / K, "last": None if last is None else [float(x) for x in last]},
            "gpu_launches": int(launches), "gpu_launches_per_step": launches / K,
            "roofline": roofline, "cpu_baseline": None, "clocks": clocks}
    run.emit(line)
    run.leave()
