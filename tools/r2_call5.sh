#!/bin/bash
# Round 2, GPU call 5 (1 GPU): fused-MLP A/B with the small-problem GEMM, microbench with a clock record, ncu --set full
# over every kernel of the product paths + the launch list of one step.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_algorithms.py -q -p no:cacheprovider -k "fused" 2>&1 | tail -3
run() {
    env $2 timeout 240 python bench.py --steps 3000 --warmup 50 --no-cpu-baseline $3 > gpurun_out/r2_net_$1.json 2> gpurun_out/r2_net_$1.err
    python - "$1" <<'PY'
import json, sys
name = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/r2_net_{name}.json").read().strip().splitlines()[-1])
    print(f"{name:24s} {j['ms_per_step']*1e3:8.1f} us/update  value {j['value']:.3e}  e2e {j['e2e']['value']:.3e}  launches/step {j['gpu_launches_per_step']}")
except Exception as e:
    print(name, "FAILED", e); print(open(f"gpurun_out/r2_net_{name}.err").read()[-800:])
PY
}
run torch        "MB_X=0"            ""
run fused        "MB_X=0"            "--net fused"
run fused_legacy "MB_GEMM_LEGACY=1"  "--net fused"
run fused_serial "MB_X=0"            "--net fused --serial"
run fused_ff     "MB_X=0"            "--net fused --fuse-forwards"
timeout 400 python tools/microbench.py > gpurun_out/r2_microbench.json 2> gpurun_out/r2_microbench.err; tail -c 400 gpurun_out/r2_microbench.json
MB_ROWS_LEGACY=1 timeout 400 python tools/microbench.py > gpurun_out/r2_microbench_rows_legacy.json 2> gpurun_out/r2_microbench_rows_legacy.err
K='per_sample|rows_gather_bulk|rows_move|gemm_small|gemm_kernel|td_dqn|td_target|td_dist|tree_update_part|tree_build|tree_find|tree_leaf|adam|soft_update|rows_pack|rows_unpack|wire_|vtrace|gae_kernel'
NCU_ROWS=10000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -c 120 \
    -o gpurun_out/r2_all python tools/ncu_targets_r2.py > gpurun_out/r2_ncu_all.log 2>&1
tail -2 gpurun_out/r2_ncu_all.log
ncu -i gpurun_out/r2_all.ncu-rep --page raw --csv > gpurun_out/r2_all_raw.csv 2>/dev/null
python tools/ncu_summarize.py gpurun_out/r2_all_raw.csv --traffic gpurun_out/r2_all_traffic.json > gpurun_out/r2_all_summary.md
head -60 gpurun_out/r2_all_summary.md
