#!/bin/bash
# Round 2, GPU call (1 GPU): the default update with MB_CARVEOUT=max (every kernel of the step prefers the largest shared-memory carve-out).
mkdir -p gpurun_out
MB_CARVEOUT=max timeout 75 python bench.py --steps 2000 --warmup 30 --no-cpu-baseline > gpurun_out/r2_n1_carveout.json 2> gpurun_out/r2_n1_carveout.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2_n1_carveout.json").read().strip().splitlines()[-1])
    print(f"carveout=max N=1 {j['ms_per_step']*1e3:8.1f} us/step value {j['value']:.3e} e2e {j['e2e']['value']:.3e}")
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/r2_n1_carveout.err").read()[-1200:])
PY
