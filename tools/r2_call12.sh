#!/bin/bash
# Round 2, GPU call 13 (1 GPU): FusedMLP backward with the weight-gradient GEMMs on a side branch -- parity + step time.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_algorithms.py tests/test_gpu_zz_overlap.py tests/test_gpu_zz_static_exchange.py tests/test_gpu_zz_ddpg_graph.py -q -p no:cacheprovider --tb=short -rf 2>&1 | tail -6
timeout 240 python bench.py --steps 3000 --warmup 50 --no-cpu-baseline > gpurun_out/r2_n1_sidegrads.json 2> gpurun_out/r2_n1_sidegrads.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2_n1_sidegrads.json").read().strip().splitlines()[-1])
print(f"default: {j['ms_per_step']*1e3:.1f} us/update value {j['value']:.3e} e2e {j['e2e']['value']:.3e} ({j['e2e']['ms_per_step']*1e3:.1f} us)")
PY
timeout 240 python bench.py --workload c4 --replay-rows 2000000 --steps 500 --warmup 10 --no-cpu-baseline --net fused > gpurun_out/r2_c4_n1_fused.json 2> gpurun_out/r2_c4_n1_fused.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2_c4_n1_fused.json").read().strip().splitlines()[-1])
    print(f"c4 N=1 fused nets: {j['ms_per_step']*1e3:.1f} us/update value {j['value']:.3e}")
except Exception as e:
    print("c4 fused FAILED", e, open("gpurun_out/r2_c4_n1_fused.err").read()[-800:])
PY
