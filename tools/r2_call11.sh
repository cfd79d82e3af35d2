#!/bin/bash
# Round 2, GPU call 11 (1 GPU): the whole-K cp.async GEMM -- parity, then the step with each GEMM variant.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_algorithms.py tests/test_gpu_zz_overlap.py tests/test_gpu_zz_rows_program.py -q -p no:cacheprovider --tb=short -rf 2>&1 | tail -15
for m in wk small legacy; do
  MB_GEMM=$m timeout 240 python bench.py --steps 3000 --warmup 50 --no-cpu-baseline > gpurun_out/r2_gemm_$m.json 2> gpurun_out/r2_gemm_$m.err
  python - $m <<'PY'
import json, sys
m = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/r2_gemm_{m}.json").read().strip().splitlines()[-1])
    print(f"MB_GEMM={m:7s} {j['ms_per_step']*1e3:8.1f} us/update value {j['value']:.3e} e2e {j['e2e']['value']:.3e}")
    for k, v in j["kernels"].items():
        if "gemm" in k: print(f"    {k:55s} {v['avg_us']:.2f} us x{v['per_step']}")
except Exception as e:
    print(m, "FAILED", e); print(open(f"gpurun_out/r2_gemm_{m}.err").read()[-1200:])
PY
done

