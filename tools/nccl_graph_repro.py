"""Minimal repro: NCCL collectives captured in a CUDA graph, interleaved with an eager all-gather per replay.
torchrun --nproc-per-node 2 tools/nccl_graph_repro.py [thread_local|global|relaxed] [keep|release]

Round 1: every stage completes (0.082 ms per replay on 2 GPUs); destroy_process_group() then blocks while the graph
object is alive ("keep", the default).  "release" drops the graph (and synchronises) before the teardown -- whether
that is enough decides if the library can offer release_graphs() instead of bench.py leaving through os._exit().
Run it under `timeout`: a blocked teardown never returns."""
import os, sys, time
import torch as t, torch.distributed as dist
mode = sys.argv[1] if len(sys.argv) > 1 else "thread_local"
teardown = sys.argv[2] if len(sys.argv) > 2 else "keep"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
t.cuda.set_device(local); dev = t.device("cuda", local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=dev)
def say(*a):
    if rank == 0: print(*a, flush=True)
P, cap, W = world, 512, 712
send, recv = t.zeros((P * cap, W), dtype=t.uint8, device=dev), t.zeros((P * cap, W), dtype=t.uint8, device=dev)
grad = t.ones(100000, device=dev); root = t.ones(1, dtype=t.float64, device=dev) * (rank + 1)
sums = t.zeros(P, dtype=t.float64, device=dev)
def body():
    send.add_(1)
    dist.all_to_all_single(recv, send)
    grad.mul_(1.0001)
    dist.all_reduce(grad); grad.div_(P)
    dist.all_to_all_single(send, recv)
def prepare():
    dist.all_gather_into_tensor(sums, root)
    return sums.tolist()
prepare(); say("eager all_gather ok")
cur, side = t.cuda.current_stream(), t.cuda.Stream()
side.wait_stream(cur)
with t.cuda.stream(side):
    for _ in range(3): body()
cur.wait_stream(side); t.cuda.synchronize(); say("warm-up ok")
g = t.cuda.CUDAGraph()
with t.cuda.graph(g, capture_error_mode=mode):
    body()
t.cuda.synchronize(); say("capture ok", mode)
dist.barrier(); say("barrier ok")
t0 = time.time()
for i in range(50):
    prepare(); g.replay()
t.cuda.synchronize(); say("50 replays with interleaved eager all_gather ok", round((time.time() - t0) / 50 * 1e3, 3), "ms each")
if teardown == "release":
    g.reset()
    del g
    t.cuda.synchronize()
    dist.barrier()
    say("graph released")
dist.destroy_process_group()
say("process group destroyed (", teardown, ")")
