#!/bin/bash
# Round 2, GPU call 16 (2 GPUs): gradient sinks + side-branch weight gradients in the fixed-shape update -- parity, then timing.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_static_exchange.py tests/test_gpu_distributed.py tests/test_gpu_algorithms.py tests/test_gpu_zz_overlap.py -q -rfEs --tb=short -p no:cacheprovider > gpurun_out/r2_pytest13.txt 2>&1
tail -6 gpurun_out/r2_pytest13.txt
show() { python - "$1" <<'PY'
import json, sys
name = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/{name}.json").read().strip().splitlines()[-1])
    print(f"{name:26s} N={j['n_gpus']} {j['ms_per_step']*1e3:8.1f} us/step  value {j['value']:.3e}  e2e {j['e2e']['value']:.3e}  launches/step {j.get('gpu_launches_per_step')}")
except Exception as e:
    print(name, "FAILED", e)
    try: print(open(f"gpurun_out/{name}.err").read()[-1500:])
    except Exception: pass
PY
}
tr() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
tr 200 29751 bench.py --gpus 2 --steps 1000 --warmup 20 --watchdog-s 150 > gpurun_out/r2_n2_final.json 2> gpurun_out/r2_n2_final.err; show r2_n2_final
tr 240 29752 bench.py --gpus 2 --workload c4 --replay-rows 2000000 --steps 500 --warmup 10 --watchdog-s 200 --net fused > gpurun_out/r2_c4_n2_fused.json 2> gpurun_out/r2_c4_n2_fused.err; show r2_c4_n2_fused
timeout 200 python bench.py --steps 3000 --warmup 50 --no-cpu-baseline > gpurun_out/r2_n1_check.json 2> gpurun_out/r2_n1_check.err; show r2_n1_check
