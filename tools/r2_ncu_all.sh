#!/bin/bash
# ncu --set full over every kernel of the product paths, in three passes with focused kernel filters; only the raw CSV
# and the summary are kept (gpurun merges at most 64 MiB back).
mkdir -p gpurun_out
pass() {   # name, regex, count
    NCU_ROWS=4000000 timeout 600 ncu --set full --clock-control none -k regex:"$2" -c $3 -o gpurun_out/r2_ncu_$1 \
        python tools/ncu_targets_r2.py > gpurun_out/r2_ncu_$1.log 2>&1
    ncu -i gpurun_out/r2_ncu_$1.ncu-rep --page raw --csv > gpurun_out/r2_ncu_$1_raw.csv 2>/dev/null
    rm -f gpurun_out/r2_ncu_$1.ncu-rep
    python tools/ncu_summarize.py gpurun_out/r2_ncu_$1_raw.csv --traffic gpurun_out/r2_ncu_$1_traffic.json > gpurun_out/r2_ncu_$1_summary.md
    python - "$1" <<'PY'
import csv, sys
name = sys.argv[1]
path = f"gpurun_out/r2_ncu_{name}_raw.csv"
rows = list(csv.reader(open(path)))
hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name", "Block Size", "Grid Size") or any(k in h for k in (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__shared_mem_config_size", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__occupancy_limit", "smsp__cycles_active.avg", "l1tex__t_bytes.sum"))]
with open(path, "w", newline="") as f:          # keep the columns the summaries use: the full raw page is ~30 MB
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in keep if i < len(r)])
PY
    cat gpurun_out/r2_ncu_$1_summary.md
}
pass step   'per_sample_cluster$|rows_gather_bulk|gemm_wk|gemm_kernel|td_dqn|tree_update_part|adam|soft_update' 56
pass wire   'per_sample_cluster_multi|rows_move_kernel|rows_pack|rows_unpack|wire_|peer_' 44
pass scans  'vtrace|gae_kernel|td_target|td_dist|tree_find|tree_leaf|tree_build' 40
