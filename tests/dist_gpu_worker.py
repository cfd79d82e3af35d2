"""torchrun worker for tests/test_gpu_distributed.py: the sharded prioritized buffer over NCCL, checked
against the oracle and against the reference's distributed golden vector."""
import json
import os
import sys

import numpy as np
import torch as t
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    t.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=t.device("cuda", local))
    from machin_b200.frame.buffers import DistributedPrioritizedBuffer
    from machin_b200.parallel import ShardGroup
    from oracle import per as oper
    from oracle.sumtree import OracleTree

    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "kat_prioritized_buffer.json")))["buffer"]
    group = ShardGroup("group")
    # ---- 1. the reference's controlled test: members 0..world-2 store the golden episode, the last one samples
    np.random.seed(0)
    buf = DistributedPrioritizedBuffer("buffer", group, 5, requesters=[world - 1])
    if rank < world - 1:
        ep = [dict(state={"state_1": t.zeros([1, 2])}, action={"action_1": t.zeros([1, 3])},
                   next_state={"next_state_1": t.zeros([1, 2])}, reward=1, terminal=True, index=i) for i in range(5)]
        buf.store_episode(ep, priorities=kat["sample"]["priorities"])
    n, sample, indexes, isw = buf.sample_batch(5 * (world - 1), sample_attrs=["index"])
    if rank == world - 1:
        assert n == 5 * (world - 1), n
        assert sample[0] == kat["sample"]["index"] * (world - 1), sample[0]
        assert list(indexes.keys()) == [str(r) for r in range(world - 1)]
        assert np.all(np.abs(isw - np.array(kat["sample"]["is_weight"] * (world - 1))) < 1e-6)
        buf.update_priority(isw, indexes)
    else:
        assert n == 0
        buf.update_priority(None, None)
    assert buf.all_size() == 5 * (world - 1)
    if rank == 0:
        print(f"reference distributed golden ok (world {world})", flush=True)
    # ---- 2. every member a learner: parity of each shard's draws / write-backs against the oracle
    np.random.seed(10 + rank)
    cap, S = 4096, 84
    buf = DistributedPrioritizedBuffer("buffer2", group, cap)
    rng = np.random.RandomState(rank)
    n_rows = min(3000 + 250 * rank, 4000)                 # (a shard never holds more than its ring of 4096)
    state = rng.randn(n_rows, S).astype(np.float32)
    pri = rng.exponential(size=n_rows)
    buf.store_episode_tensors({"state": {"state": t.from_numpy(state)}, "action": {"action": t.zeros(n_rows, 1, dtype=t.int64) + rank},
                               "next_state": {"state": t.from_numpy(state)}},
                              {"reward": t.from_numpy(pri.astype(np.float32)).view(-1, 1), "terminal": t.zeros(n_rows, 1, dtype=t.bool)},
                              priorities=pri)
    ora = OracleTree(cap)
    # the distributed store normalises ONE priority at a time (prioritized_buffer_d.py:132-136): NumPy's
    # scalar pow and its SIMD array pow can differ in the last ulp, so the oracle takes the scalar path too
    leaf0 = np.array([oper.normalize_priority(p) for p in pri])
    ora.update_leaf_batch(leaf0, np.arange(n_rows))
    assert np.array_equal(buf.wt_tree.weights, ora.weights)
    sums = group.all_gather_f64(t.tensor([ora.get_weight_sum()], dtype=t.float64, device=group.device)).view(-1).tolist()
    ssize, total = oper.shard_sample_sizes(sums, 512)
    state_np = np.random.get_state()
    expect_idx = []
    beta = 0.4
    for r in range(world):                      # this shard serves every requester in member order
        idx, _ = oper.sample_index_and_weight(ora, ssize[rank], n_rows, beta, all_weight_sum=total)
        beta = min(1.0, beta + 0.001)
        expect_idx.append(idx)
    np.random.set_state(state_np)
    n, fields, index, isw = buf.sample_device(512)
    assert n == sum(ssize)
    o = 0
    for m in range(world):
        assert t.all(fields["action/action"][o:o + ssize[m]] == m)
        o += ssize[m]
    # what I received from shard `rank` must be the draw that shard made for requester `rank`
    mine = index[str(rank)][0].cpu().numpy()
    assert np.array_equal(mine, expect_idx[rank]), "index draw differs from the oracle"
    o = sum(ssize[:rank])
    assert np.array_equal(fields["state/state"][o:o + ssize[rank]].cpu().numpy(), state[mine])
    # priority write-back with a tainted slot
    new_p = t.rand(n, device=group.device)
    victim = int(expect_idx[0][0])
    buf.buffer_version_table[victim] += 1
    buf.update_priority(new_p, index)
    all_p = [t.empty_like(new_p) for _ in range(world)]
    # every requester r sent me new_p_r[offset_of_my_rows]; rebuild the expected tree with the oracle
    gathered = [t.empty(n, device=group.device) for _ in range(world)]
    dist.all_gather(gathered, new_p)
    o = sum(ssize[:rank])
    for r in range(world):
        pr = gathered[r][o:o + ssize[rank]].cpu().numpy()
        ix = expect_idx[r]
        keep = ix != victim
        ora.update_leaf_batch(oper.normalize_priority(pr[keep]), ix[keep])
    got, want = buf.wt_tree.weights, ora.weights
    assert np.allclose(got, want, rtol=1e-6, atol=0), np.abs(got - want).max()
    assert got[victim] == leaf0[victim]
    if rank == 0:
        print(f"sharded draw / gather / version-filtered write-back ok (world {world}): tainted slot {victim} kept", flush=True)
    # ---- 3. weight server over NCCL: the reference's goldens 11/12/13, a pull that is not newer, a rejected push on
    #         a stale chain, two pushers in one round (param_server.py:36-91, ordered_server.py:112-128)
    from server_cases import grad_expect, grad_server_case, model_server_case
    dev = t.device("cuda", local)
    assert model_server_case(rank, world, device=dev)
    if rank == 0:
        print(f"model server ok (world {world}): goldens 11/12/13, stale pull, rejected push, racing pushers", flush=True)
    # ---- 4. A3C gradient server over NCCL (param_server.py:411-493; test_param_server.py:108-173)
    for method in ("mean", "sum"):
        assert grad_server_case(rank, world, method, device=dev)
        if rank == 0:
            print(f"grad server ok (world {world}): {method} -> {grad_expect(method, world)}", flush=True)
    dist.barrier()
    if rank == 0:
        print("dist ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
