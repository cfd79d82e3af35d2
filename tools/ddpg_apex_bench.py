#!/usr/bin/env python
"""BASELINE config C4: APEX-DDPG learner over a sharded prioritized buffer (state 376, action 17, batch 512), one
replay shard + one learner per GPU, eager NCCL exchange.  Run under torchrun; rank 0 prints one JSON object.

    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29577 \\
        tools/ddpg_apex_bench.py --rows 12500000 --steps 200

(12.5 M rows x 3104 B = 38.8 GB per shard, 50 M rows over 4 GPUs; --rows scales it down for a first run.)"""
import argparse
import json
import os
import sys
import time

import torch as t
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

S, A, H = 376, 17, 256


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


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--rows", type=int, default=1_000_000, help="transitions per shard (C4: 12.5 M)")
    p.add_argument("--batch", type=int, default=512)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--graph", action="store_true", help="one GPU only: DDPGPer over a single-process prioritized buffer with "
                   "the whole update captured as one CUDA graph (DDPGPer.enable_cuda_graph)")
    a = p.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    t.cuda.set_device(local)
    dev = t.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    from machin_b200.frame.algorithms import DDPGApex
    from machin_b200.model import FusedAdam
    from machin_b200.parallel import LocalGroup, ShardGroup
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = ShardGroup("apex", device=dev)
    else:
        group = LocalGroup(device=dev)
    t.manual_seed(0)                                    # identical initial weights on every learner
    nets = [Actor().to(dev), Actor().to(dev), Critic().to(dev), Critic().to(dev)]
    if a.graph:
        if world != 1:
            raise SystemExit("--graph is the single-GPU variant")
        from machin_b200.frame.algorithms import DDPGPer
        learner = DDPGPer(*nets, FusedAdam, t.nn.MSELoss(reduction="none"), batch_size=a.batch, replay_size=a.rows)
    else:
        learner = DDPGApex(*nets, FusedAdam, t.nn.MSELoss(reduction="none"), group, (), batch_size=a.batch, replay_size=a.rows,
                           learner_group=group if world > 1 else None)
    gen = t.Generator(device=dev).manual_seed(rank)
    done = 0
    while done < a.rows:
        n = min(500_000, a.rows - done)
        learner.replay_buffer.store_episode_tensors(
            {"state": {"state": t.randn(n, S, device=dev, generator=gen)},
             "action": {"action": t.rand(n, A, device=dev, generator=gen) * 2 - 1},
             "next_state": {"state": t.randn(n, S, device=dev, generator=gen)}},
            {"reward": t.randn(n, 1, device=dev, generator=gen), "terminal": t.rand(n, 1, device=dev, generator=gen) < 0.01},
            priorities=t.empty(n, device=dev, dtype=t.float64).exponential_(1.0, generator=gen))
        done += n
    learner.replay_buffer.wt_tree.check()

    def barrier():
        t.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t.cuda.synchronize(dev)

    if a.graph:
        learner.enable_cuda_graph()
    for _ in range(max(a.warmup, 3)):
        learner.update()
    barrier()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(a.steps):
        out = learner.update()
    e1.record()
    barrier()
    ms = t.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)], dtype=t.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms_step = float(ms) / a.steps
        print(json.dumps({"workload": f"C4 APEX-DDPG learner: PER {a.rows} rows per shard x (state {S}, act {A}), batch {a.batch}",
                          "n_gpus": world, "ms_per_update": ms_step, "updates_per_sec": 1e3 / ms_step,
                          "transitions_sampled_per_sec": world * a.batch * 1e3 / ms_step, "last": [float(x) for x in out],
                          "launch": ("one CUDA graph per update" if a.graph else "eager") + " (host-synchronous update(): loss read back every step)",
                          "steps": a.steps}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
