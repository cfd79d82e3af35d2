#!/usr/bin/env python
"""BASELINE config C5: IMPALA with N - 1 actor GPUs feeding one learner GPU over NCCL (trajectory push = rows + episode
lengths through the sharded episode buffer, weight broadcast through the model server), batch 4096 episodes of T = 20,
state 256.  Run under torchrun; rank 0 (the learner) prints one JSON object.

    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29578 \\
        tools/impala_dist_bench.py --steps 50
"""
import argparse
import json
import os
import sys
import time

import torch as t
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

S, A, T = 256, 6, 20


class Actor(t.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = t.nn.Linear(S, 256), t.nn.Linear(256, 256), t.nn.Linear(256, A)

    def forward(self, state, action=None):
        logp_all = t.log_softmax(self.fc3(t.relu(self.fc2(t.relu(self.fc1(state))))), dim=1)
        act = action if action is not None else t.multinomial(logp_all.exp(), 1)
        return act.view(-1, 1), logp_all.gather(1, act.view(-1, 1).long()), -(logp_all.exp() * logp_all).sum(1, keepdim=True)


class Critic(t.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = t.nn.Linear(S, 256), t.nn.Linear(256, 256), t.nn.Linear(256, 1)

    def forward(self, state):
        return self.fc3(t.relu(self.fc2(t.relu(self.fc1(state)))))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--episodes", type=int, default=4096, help="episodes per learner batch (C5: 4096)")
    p.add_argument("--stored", type=int, default=4096, help="episodes held by every actor")
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=3)
    a = p.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    t.cuda.set_device(local)
    dev = t.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    from machin_b200.frame.algorithms import IMPALA
    from machin_b200.parallel import ShardGroup
    from machin_b200.parallel.server import PushPullModelServer
    group = ShardGroup("impala", device=dev)
    t.manual_seed(0)
    imp = IMPALA(Actor().to(dev), Critic().to(dev), t.optim.Adam, t.nn.MSELoss(reduction="sum"), group,
                 (PushPullModelServer("actor", group, 0),), batch_size=a.episodes, replay_size=2 * a.stored * T, requesters=[0])
    if rank > 0:                                         # actors hold the trajectories
        for _ in range(a.stored):
            term = t.zeros(T, 1, dtype=t.bool)
            term[-1] = True
            el = t.zeros(T, 1, dtype=t.int64)
            el[0] = T
            imp.replay_buffer.store_episode_tensors(
                {"state": {"state": t.randn(T, S, device=dev)}, "action": {"action": t.randint(0, A, (T, 1), device=dev)},
                 "next_state": {"state": t.randn(T, S, device=dev)}}, {"reward": t.randn(T, 1, device=dev), "terminal": term},
                {"action_log_prob": -t.rand(T, 1, device=dev), "episode_length": el})
    dist.barrier()
    for _ in range(a.warmup):
        imp.update()
    t.cuda.synchronize(dev)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        out = imp.update()
    t.cuda.synchronize(dev)
    dist.barrier()
    ms = (time.perf_counter() - t0) * 1e3 / a.steps
    if rank == 0:
        per_member = -(-a.episodes // world)
        print(json.dumps({"workload": f"C5 IMPALA: {world - 1} actor GPU(s) -> 1 learner, {per_member} episodes x T={T} per member "
                                      f"per batch, state {S}", "n_gpus": world, "ms_per_update": ms, "updates_per_sec": 1e3 / ms,
                          "steps_per_sec": (world - 1) * per_member * T * 1e3 / ms, "last": [float(x) for x in out],
                          "steps": a.steps}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
