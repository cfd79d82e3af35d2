"""torchrun worker for tests/test_gpu_zz_static_exchange.py: the fixed-shape single-graph APEX update
(NCCL collectives captured) against the dynamic exchange, every member a learner, same seeds."""
import os
import sys

import numpy as np
import torch as t
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class QNet(t.nn.Module):
    def __init__(self, s, a, h):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = t.nn.Linear(s, h), t.nn.Linear(h, h), t.nn.Linear(h, a)

    def forward(self, state):
        return self.fc3(t.relu(self.fc2(t.relu(self.fc1(state)))))


def flat(m):
    return t.cat([p.detach().reshape(-1) for p in m.parameters()]).cpu().numpy()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    t.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=t.device("cuda", local))
    from machin_b200.frame.algorithms import DQNApex
    from machin_b200.parallel import ShardGroup
    S, A, B = 12, 5, 64
    N = min(3000 + 250 * rank, 4000)                      # (a shard never holds more than its ring of 4096)
    group = ShardGroup("group")

    def make():
        t.manual_seed(0)                                  # identical initial weights on every member
        q, qt = QNet(S, A, 32).cuda(), QNet(S, A, 32).cuda()
        learner = DQNApex(q, qt, t.optim.SGD, t.nn.MSELoss(reduction="none"), group, (), batch_size=B,
                          learning_rate=0.05, discount=0.9, update_rate=0.1, replay_size=4096, learner_group=group)
        rng = np.random.RandomState(3 + rank)             # a different shard on every member
        learner.replay_buffer.store_episode_tensors(
            {"state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))},
             "action": {"action": t.from_numpy(rng.randint(0, A, size=(N, 1)))},
             "next_state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))}},
            {"reward": t.from_numpy(rng.randn(N, 1).astype(np.float32)), "terminal": t.from_numpy(rng.rand(N, 1) < 0.1)},
            priorities=rng.exponential(size=N))
        return learner

    a, b = make(), make()
    b.enable_cuda_graph(static_exchange=True)
    assert np.array_equal(a.replay_buffer.wt_tree.weights, b.replay_buffer.wt_tree.weights)
    for step in range(6):
        np.random.seed(100 * rank + step)
        la = a.update_device().item()
        np.random.seed(100 * rank + step)
        lb = b.update_device().item()
        assert abs(la - lb) <= 1e-5 * max(1.0, abs(la)), (step, la, lb)
        np.testing.assert_allclose(b.replay_buffer.wt_tree.weights, a.replay_buffer.wt_tree.weights, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(flat(b.qnet), flat(a.qnet), rtol=1e-4, atol=1e-6)
        assert a.replay_buffer.curr_beta == b.replay_buffer.curr_beta
    # device-side apportioning (host never reads the shard sums): a valid trajectory -- finite losses, consistent
    # trees, identical weights on every member, beta advanced once per requester served
    c = make()
    c.batched_draw = True                                 # all requesters' draws in one launch
    c.enable_cuda_graph(static_exchange=True, device_apportion=True)
    beta0 = float(c.replay_buffer.curr_beta)
    for step in range(6):
        lc = c.update_device().item()
        assert np.isfinite(lc) and lc >= 0.0, (step, lc)
    t.cuda.synchronize()
    c.replay_buffer.wt_tree.check()
    assert abs(float(c._static["beta"].item()) - min(1.0, beta0 + 6 * world * 1e-3)) < 1e-12
    wc = t.from_numpy(flat(c.qnet)).cuda()
    wcs = [t.empty_like(wc) for _ in range(world)]
    dist.all_gather(wcs, wc)
    assert all(t.equal(wcs[0], x) for x in wcs)
    # ---- peer exchange (no NCCL inside the update: rows, priorities, shard sums and gradients are stored into the
    #      other GPUs' buffers over NVLink by csrc/peer.cu): the same trajectory as the NCCL fixed-shape update, with
    #      host apportioning (draw for draw against `b2`) and with device apportioning + the stream-level options
    b2, d = make(), make()
    b2.enable_cuda_graph(static_exchange=True)
    d.enable_cuda_graph(static_exchange=True, peer_exchange=True)
    peer_on = d._static["peer"] is not None
    for step in range(6):
        np.random.seed(500 * rank + step)
        lb2 = b2.update_device().item()
        np.random.seed(500 * rank + step)
        ld = d.update_device().item()
        assert abs(ld - lb2) <= 1e-5 * max(1.0, abs(lb2)), (step, ld, lb2)
        np.testing.assert_allclose(d.replay_buffer.wt_tree.weights, b2.replay_buffer.wt_tree.weights, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(flat(d.qnet), flat(b2.qnet), rtol=1e-4, atol=1e-6)
    if peer_on:
        d._static["peer"].check()
    e = make()
    e.batched_draw = e.parallel_forwards = e.overlap_priority_update = True
    e.enable_cuda_graph(static_exchange=True, device_apportion=True, peer_exchange=True)
    for step in range(8):
        le = e.update_device().item()
        assert np.isfinite(le) and le >= 0.0, (step, le)
    t.cuda.synchronize()
    e.replay_buffer.wt_tree.check()
    if e._static["peer"] is not None:
        e._static["peer"].check()
    we = t.from_numpy(flat(e.qnet)).cuda()
    wes = [t.empty_like(we) for _ in range(world)]
    dist.all_gather(wes, we)
    assert all(t.equal(wes[0], x) for x in wes)           # bit-identical weights on every member (ordered peer sum)
    if rank == 0:
        print(f"peer exchange {'ok' if peer_on else 'UNAVAILABLE (NCCL fallback ran)'} (world {world})", flush=True)
    # ---- the sharded DDPG update (BASELINE C4) as one graph over the same exchange: against the dynamic, eager exchange
    #      draw for draw (host apportioning), NCCL transport and peer transport
    from machin_b200.frame.algorithms import DDPGApex
    AD = 3

    class Actor(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2 = t.nn.Linear(S, 16), t.nn.Linear(16, AD)

        def forward(self, state):
            return t.tanh(self.fc2(t.relu(self.fc1(state))))

    class Critic(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2 = t.nn.Linear(S + AD, 16), t.nn.Linear(16, 1)

        def forward(self, state, action):
            return self.fc2(t.relu(self.fc1(t.cat([state, action], dim=1))))

    def make_ddpg():
        t.manual_seed(7)
        nets = [Actor().cuda(), Actor().cuda(), Critic().cuda(), Critic().cuda()]
        learner = DDPGApex(*nets, t.optim.SGD, t.nn.MSELoss(reduction="none"), group, (), batch_size=B, actor_learning_rate=0.01,
                           critic_learning_rate=0.02, discount=0.9, update_rate=0.1, replay_size=4096, learner_group=group)
        rng = np.random.RandomState(11 + rank)
        learner.replay_buffer.store_episode_tensors(
            {"state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))},
             "action": {"action": t.from_numpy(rng.rand(N, AD).astype(np.float32) * 2 - 1)},
             "next_state": {"state": t.from_numpy(rng.randn(N, S).astype(np.float32))}},
            {"reward": t.from_numpy(rng.randn(N, 1).astype(np.float32)), "terminal": t.from_numpy(rng.rand(N, 1) < 0.1)},
            priorities=rng.exponential(size=N))
        return learner

    da, dn, dp = make_ddpg(), make_ddpg(), make_ddpg()
    dn.enable_cuda_graph(device_apportion=False, peer_exchange=False)
    dp.enable_cuda_graph(device_apportion=False, peer_exchange=True)
    for step in range(5):
        outs = []
        for learner in (da, dn, dp):
            np.random.seed(900 * rank + step)
            outs.append([x.item() for x in learner.update_device()])
        for other in outs[1:]:
            for x, y in zip(outs[0], other):
                assert abs(x - y) <= 1e-5 * max(1.0, abs(x)), (step, outs)
        for learner in (dn, dp):
            np.testing.assert_allclose(learner.replay_buffer.wt_tree.weights, da.replay_buffer.wt_tree.weights, rtol=1e-5, atol=1e-9)
            np.testing.assert_allclose(flat(learner.actor), flat(da.actor), rtol=1e-4, atol=1e-6)
            np.testing.assert_allclose(flat(learner.critic), flat(da.critic), rtol=1e-4, atol=1e-6)
    if dp._static["peer"] is not None:
        dp._static["peer"].check()
    if rank == 0:
        print(f"sharded DDPG fixed-shape update ok (world {world})", flush=True)
    # every member holds the same weights after averaged gradients
    w = t.from_numpy(flat(b.qnet)).cuda()
    ws = [t.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    assert all(t.equal(ws[0], x) for x in ws)
    t.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("static ok")
    sys.stdout.flush()
    # destroy_process_group() blocks while a graph holding captured NCCL work is alive: leave directly
    os._exit(0)


if __name__ == "__main__":
    main()
