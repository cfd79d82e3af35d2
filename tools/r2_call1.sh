#!/bin/bash
# Round 2, GPU call 1 (1 GPU): the whole -m gpu suite with the xfail mask removed (no -x: every failure is
# listed with its traceback), then the priority write-back variants side by side.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus.txt
timeout 900 python -m pytest tests -m gpu -q -rfEs --tb=short -p no:cacheprovider > gpurun_out/r2_pytest1.txt 2>&1
tail -40 gpurun_out/r2_pytest1.txt
for v in "MB_X=0" "MB_SORTED_CLUSTER=1" "MB_SORTED_ASYNC=1" "MB_SORTED_CLUSTER=1 MB_SORTED_ASYNC=1"; do
    env $v timeout 120 python tools/ub_sorted_update.py 2>gpurun_out/r2_ub.err | tee -a gpurun_out/r2_sorted_update.jsonl
done
