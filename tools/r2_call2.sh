#!/bin/bash
# Round 2, GPU call 2 (2 GPUs): whole -m gpu suite incl. the world-2 NCCL tests, write-back kernel A/B, N=1 bench.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus2.txt
timeout 1200 python -m pytest tests -m gpu -q -rfEs --tb=short -p no:cacheprovider > gpurun_out/r2_pytest2.txt 2>&1
tail -30 gpurun_out/r2_pytest2.txt
rm -f gpurun_out/r2_update_ab.jsonl
for v in "MB_X=0" "MB_UPDATE_LEGACY=1" "MB_SORTED_CLUSTER=1"; do
    env $v timeout 120 python tools/ub_sorted_update.py 2>gpurun_out/r2_ub.err | tee -a gpurun_out/r2_update_ab.jsonl
done
timeout 300 python bench.py --steps 2000 --warmup 50 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err
tail -c 1500 gpurun_out/r2_bench_n1_a.json
