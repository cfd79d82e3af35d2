#!/usr/bin/env python
"""IMPALA learner update at BASELINE config C3 (1024 episodes x T=20, state 256) on one GPU, next to the
oracle port of the reference's V-trace block on the host.  Prints one JSON object."""
import json
import os
import sys
import time

import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from machin_b200.frame.algorithms import IMPALA  # noqa: E402
from machin_b200.parallel import LocalGroup  # noqa: E402
from oracle import vtrace as ovt  # noqa: E402

S, A, T, EP = 256, 6, 20, 1024


class Actor(t.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = t.nn.Linear(S, 256), t.nn.Linear(256, 256), t.nn.Linear(256, A)

    def forward(self, state, action=None):
        logits = self.fc3(t.relu(self.fc2(t.relu(self.fc1(state)))))
        logp_all = t.log_softmax(logits, dim=1)
        act = action if action is not None else t.multinomial(logp_all.exp(), 1)
        logp = logp_all.gather(1, act.view(-1, 1).long())
        ent = -(logp_all.exp() * logp_all).sum(1, keepdim=True)
        return act.view(-1, 1), logp, ent


class Critic(t.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = t.nn.Linear(S, 256), t.nn.Linear(256, 256), t.nn.Linear(256, 1)

    def forward(self, state):
        return self.fc3(t.relu(self.fc2(t.relu(self.fc1(state)))))


def main():
    dev = t.device("cuda", 0)
    t.manual_seed(0)
    imp = IMPALA(Actor().to(dev), Critic().to(dev), t.optim.Adam, t.nn.MSELoss(reduction="sum"), LocalGroup(device=dev), (),
                 batch_size=EP, replay_size=2 * EP * T)
    buf = imp.replay_buffer
    for e in range(EP):
        term = t.zeros(T, 1, dtype=t.bool)
        term[-1] = True
        el = t.zeros(T, 1, dtype=t.int64)
        el[0] = T
        buf.store_episode_tensors(
            {"state": {"state": t.randn(T, S, device=dev)}, "action": {"action": t.randint(0, A, (T, 1), device=dev)},
             "next_state": {"state": t.randn(T, S, device=dev)}},
            {"reward": t.randn(T, 1, device=dev), "terminal": term},
            {"action_log_prob": -t.rand(T, 1, device=dev), "episode_length": el})
    graph = "--graph" in sys.argv[1:]          # the whole update of the fixed-length batch as one CUDA graph replay
    if graph:
        imp.enable_cuda_graph(episode_length=T)
    for _ in range(3):
        imp.update()
    t.cuda.synchronize()
    n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        imp.update()
    t.cuda.synchronize()
    gpu_ms = (time.perf_counter() - t0) / n * 1e3
    # oracle: the reference's V-trace block (per-step Python loop) on the same shapes, host cores
    steps = EP * T
    g = t.Generator().manual_seed(0)
    v, nv, r = (t.randn(steps, 1, generator=g) for _ in range(3))
    lp, lm = -t.rand(steps, 1, generator=g), -t.rand(steps, 1, generator=g)
    te = t.zeros(steps, 1, dtype=t.bool)
    te[T - 1::T] = True
    c, rho, delta = ovt.vtrace_prepare(lp, lm, v, nv, r, te, 0.99, 1.0, 1.0)
    sub = min(64, EP)                          # bounded sample: 64 episodes of the loop, scaled up
    t0 = time.perf_counter()
    ovt.vtrace_loop(v[:sub * T], delta[:sub * T], c[:sub * T], [T] * sub, 0.99)
    loop_ms = (time.perf_counter() - t0) * 1e3 * EP / sub
    print(json.dumps({"config": f"C3 IMPALA {EP} episodes x T={T}, state {S}", "launch": "one CUDA graph per update" if graph else "eager",
                      "gpu_update_ms": gpu_ms,
                      "gpu_steps_per_s": steps / gpu_ms * 1e3,
                      "cpu_vtrace_loop_ms_scaled": loop_ms, "cpu_us_per_step": loop_ms * 1e3 / steps,
                      "note": "gpu: whole IMPALA.update() (episode sampling, gather, actor/critic forward+backward, "
                              "V-trace kernel, Adam); cpu: ONLY the reference's V-trace recurrence loop (impala.py:340-362)"}))


if __name__ == "__main__":
    main()
