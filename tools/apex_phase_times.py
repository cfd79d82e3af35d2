"""torchrun helper: host-side time per phase of the sharded APEX learner step (rank 0)."""
import argparse, os, sys, time
import torch as t, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
t.cuda.set_device(local); dev = t.device("cuda", local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=dev)
a = argparse.Namespace(gpus=world, steps=3, warmup=3, impl="b200", replay_rows=2_000_000, batch=512, cpu_storage_rows=1000,
                       cpu_steps=1, no_cpu_baseline=True, no_graph=False, net="torch", adam="fused", fuse_forwards=False)
L = bench.build_learner(a, dev, seed=rank, world=world)
L.enable_cuda_graph()
st, buf = L._apex_graph, L.replay_buffer
T = {}
def tick(name, t0, sync):
    if sync: t.cuda.synchronize()
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
for sync in (False, True):
    T.clear()
    for it in range(220):
        if it == 20: T.clear()
        t0 = time.perf_counter(); rows, index, isw, _ = buf._exchange().sample(L.batch_size); tick("1 sample exchange", t0, sync)
        n = int(rows.shape[0])
        t0 = time.perf_counter(); st["rows"][:n].copy_(rows); st["isw"].zero_(); st["isw"][:n].copy_(isw * (st["cap"] / n)); tick("2 stage", t0, sync)
        t0 = time.perf_counter(); st["g1"].replay(); tick("3 graph fwd/td/bwd", t0, sync)
        t0 = time.perf_counter(); L._sync_grads(); tick("4 grad all-reduce", t0, sync)
        t0 = time.perf_counter(); st["g2"].replay(); tick("5 graph adam/soft", t0, sync)
        t0 = time.perf_counter(); buf.update_priority(st["abs"][:n], index); tick("6 priority exchange", t0, sync)
        t0 = time.perf_counter(); L.push_weights(); tick("7 weight broadcast", t0, sync)
    t.cuda.synchronize()
    if rank == 0:
        print("sync" if sync else "host-only", {k: round(v / 200 * 1e3, 3) for k, v in T.items()}, "total ms", round(sum(T.values()) / 200 * 1e3, 3))
dist.destroy_process_group()
