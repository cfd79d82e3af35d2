#!/bin/bash
# Round 2, GPU call 10 (1 GPU): the whole -m gpu suite on the final code, microbench with a clock record (default and
# legacy gather), the launch list of one default update, ncu --set full over every kernel (three passes).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rfEs --tb=short -p no:cacheprovider > gpurun_out/r2_pytest10.txt 2>&1
tail -6 gpurun_out/r2_pytest10.txt
timeout 300 python bench.py --steps 3000 --warmup 50 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; tail -c 1800 gpurun_out/r2_bench_n1_final.json
timeout 300 python bench.py --impl reference --steps 200 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; tail -c 700 gpurun_out/r2_bench_reference_arm.json
timeout 400 python tools/microbench.py > gpurun_out/r2_microbench.json 2> gpurun_out/r2_microbench.err; python - <<'PY'
import json
j = json.load(open("gpurun_out/r2_microbench.json"))
for k, v in j["results"].items():
    print(f"{k:28s} {v.get('us', v.get('ms', 0) * 1e3):10.2f} us  frac {v.get('frac', 0):.3f}")
print(j["clocks"])
PY
for mode in "" "--serial --net torch"; do
  name=default; [ -n "$mode" ] && name=round1_order_torch_net
  timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_$name.csv python tools/step_launches.py $mode > gpurun_out/r2_launches_$name.log 2>&1
done
bash tools/r2_ncu_all.sh > gpurun_out/r2_ncu_all.log 2>&1; tail -5 gpurun_out/r2_ncu_all.log
du -sh gpurun_out
