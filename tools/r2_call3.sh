#!/bin/bash
# Round 2, GPU call 3 (1 GPU): A/B of the opt-in learner variants on the C2 bench; ncu --set full of the write-back kernel.
mkdir -p gpurun_out
run() {   # name, env, flags
    env $2 timeout 240 python bench.py --steps 3000 --warmup 50 --no-cpu-baseline $3 \
        > gpurun_out/r2_ab_$1.json 2> gpurun_out/r2_ab_$1.err
    python - "$1" <<'PY'
import json, sys
name = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/r2_ab_{name}.json").read().strip().splitlines()[-1])
    print(f"{name:28s} {j['ms_per_step']*1e3:8.1f} us/update  value {j['value']:.3e}  e2e {j['e2e']['value']:.3e}  "
          f"launches/step {j['gpu_launches_per_step']}  dominant {j['roofline'].get('kernel')} {j['roofline'].get('avg_launch_us'):.1f} us")
except Exception as e:
    print(name, "FAILED", e)
PY
}
run default            "MB_X=0"  ""
run overlap_update     "MB_X=0"  "--overlap-update"
run parallel_forwards  "MB_X=0"  "--parallel-forwards"
run fuse_optimizer     "MB_X=0"  "--fuse-optimizer"
run all_streams        "MB_X=0"  "--overlap-update --parallel-forwards --fuse-optimizer"
run fused_autograd     "MB_X=0"  "--net fused"
run fused_rows         "MB_X=0"  "--net fused-rows"
run fused_rows_cpt2    "MB_ROWS_CPT=2"  "--net fused-rows"
run fused_rows_all     "MB_X=0"  "--net fused-rows --overlap-update --fuse-optimizer"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tree_update_part -s 2 -c 2 \
    -o gpurun_out/r2_part python tools/ub_sorted_update.py > gpurun_out/r2_ncu_part.log 2>&1
tail -3 gpurun_out/r2_ncu_part.log
timeout 200 python -m pytest tests/test_gpu_zz_impala_graph.py tests/test_gpu_zz_sorted_cluster.py -q -p no:cacheprovider 2>&1 | tail -3
