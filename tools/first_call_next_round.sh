#!/bin/bash
# First GPU call of the next round (one GPU, ~30 minutes): everything written after round 1's GPU budget ran
# out -- run the pending tests, then A/B the opt-in variants against the default on the bench workload.
#   gpurun --timeout 2700 -- 'bash tools/first_call_next_round.sh'
# Results land in gpurun_out/next_*.  Nothing here changes clocks or kills by pattern.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rxXs -p no:cacheprovider > gpurun_out/next_pytest.txt 2>&1
tail -25 gpurun_out/next_pytest.txt
for v in "MB_X=0" "MB_SORTED_CLUSTER=1" "MB_SORTED_ASYNC=1" "MB_SORTED_CLUSTER=1 MB_SORTED_ASYNC=1"; do
    env $v timeout 120 python tools/ub_sorted_update.py 2>/dev/null | tee -a gpurun_out/next_sorted_update.jsonl
done
run() {   # name, env, flags
    env $2 timeout 240 python bench.py --steps 3000 --warmup 50 --no-cpu-baseline $3 \
        > gpurun_out/next_bench_$1.json 2> gpurun_out/next_bench_$1.err
    python - "$1" <<'PY'
import json, sys
name = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/next_bench_{name}.json").read().strip().splitlines()[-1])
    print(f"{name:28s} {j['ms_per_step']*1e3:8.1f} us/update  value {j['value']:.3e}  e2e {j['e2e']['value']:.3e}  "
          f"dominant {j['roofline'].get('kernel')} {j['roofline'].get('avg_launch_us')} us")
except Exception as e:
    print(name, "FAILED", e)
PY
}
run default            "MB_X=0"               ""
run overlap_update     "MB_X=0"               "--overlap-update"
run parallel_forwards  "MB_X=0"               "--parallel-forwards"
run both               "MB_X=0"               "--overlap-update --parallel-forwards"
run cluster_update     "MB_SORTED_CLUSTER=1"  ""
run async_update       "MB_SORTED_ASYNC=1"    ""
run cluster_async      "MB_SORTED_CLUSTER=1 MB_SORTED_ASYNC=1"  ""
run all                "MB_SORTED_CLUSTER=1"  "--overlap-update --parallel-forwards"
run fused_autograd     "MB_X=0"               "--net fused"
run fused_rows         "MB_X=0"               "--net fused-rows"
run fused_rows_cpt2    "MB_ROWS_CPT=2"        "--net fused-rows"
run fuse_optimizer     "MB_X=0"               "--fuse-optimizer"
run fused_rows_overlap "MB_X=0"               "--net fused-rows --overlap-update"
run fused_rows_all     "MB_SORTED_CLUSTER=1 MB_SORTED_ASYNC=1"  "--net fused-rows --overlap-update --fuse-optimizer"
# C4 (APEX-DDPG shapes) on one GPU: eager sharded learner on a group of one against DDPGPer with the update captured
for v in "" "--graph"; do
    timeout 200 python tools/ddpg_apex_bench.py --rows 2000000 --steps 300 $v 2> gpurun_out/next_ddpg.err | tee -a gpurun_out/next_ddpg.jsonl
done
timeout 300 python tools/microbench.py > gpurun_out/next_microbench.json 2> gpurun_out/next_microbench.err; tail -c 600 gpurun_out/next_microbench.json
# C3 (IMPALA, 1024 episodes x T = 20) on one GPU: the whole update, next to the reference's V-trace loop on the host
for v in "" "--graph"; do
    timeout 240 python tools/impala_bench.py $v 2>> gpurun_out/next_impala.err | tee -a gpurun_out/next_impala.jsonl
done
