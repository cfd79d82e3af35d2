#!/bin/bash
# ncu captures owed from round 1 plus the new kernels (one GPU, ~6 minutes; never a multi-rank command):
#   gpurun --timeout 900 -- 'bash tools/ncu_call_next_round.sh'
mkdir -p gpurun_out
K='tree_update_sorted|dqn_rows_kernel|mlp_program_kernel|transpose_multi|per_sample_cluster|wire_apportion|adam_|soft_update|gae_kernel|td_target|rows_move_kernel|rows_pack_meta|td_dist_kernel'
MB_SORTED_CLUSTER=1 MB_SORTED_ASYNC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -c 64 \
    -o gpurun_out/prof_next python tools/ncu_targets_next.py > gpurun_out/ncu_next.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tree_update_sorted" -c 4 \
    -o gpurun_out/prof_sorted_default python tools/ncu_targets.py > gpurun_out/ncu_sorted.log 2>&1
for f in prof_next prof_sorted_default; do
    ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
    python tools/ncu_summarize.py gpurun_out/${f}_raw.csv --traffic gpurun_out/${f}_traffic.json > gpurun_out/${f}_summary.md
    cat gpurun_out/${f}_summary.md
done
