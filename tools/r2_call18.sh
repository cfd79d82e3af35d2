#!/bin/bash
# Round 2, GPU call (1 GPU): ncu launch list of ONE default update on the final code (after the stream-level changes).
mkdir -p gpurun_out
timeout 150 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final_step.csv python tools/step_launches.py > gpurun_out/r2_launches_final_step.log 2>&1
tail -3 gpurun_out/r2_launches_final_step.log; wc -l gpurun_out/r2_launches_final_step.csv
