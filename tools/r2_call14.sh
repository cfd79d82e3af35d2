#!/bin/bash
# Round 2, GPU call 19 (8 GPUs): final numbers -- C2 at N = 8 and N = 4, C4 at its full size on 4 GPUs, C5 on 8; world-4 tests.
mkdir -p gpurun_out
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
tr() { timeout $2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $3 "${@:4}"; }
tr 8 200 29761 bench.py --gpus 8 --steps 1000 --warmup 20 --watchdog-s 150 > gpurun_out/r2_n8_final.json 2> gpurun_out/r2_n8_final.err; show r2_n8_final
tr 4 200 29762 bench.py --gpus 4 --steps 1000 --warmup 20 --watchdog-s 150 > gpurun_out/r2_n4_final.json 2> gpurun_out/r2_n4_final.err; show r2_n4_final
tr 4 400 29763 bench.py --gpus 4 --workload c4 --steps 500 --warmup 10 --watchdog-s 350 > gpurun_out/r2_c4_n4_final.json 2> gpurun_out/r2_c4_n4_final.err; show r2_c4_n4_final
tr 8 300 29764 bench.py --gpus 8 --workload c5 --steps 30 --warmup 3 --watchdog-s 250 > gpurun_out/r2_c5_n8_final.json 2> gpurun_out/r2_c5_n8_final.err; show r2_c5_n8_final
timeout 400 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_zz_static_exchange.py tests/test_gpu_zz_impala_exchange.py -q -rA --tb=short -p no:cacheprovider -k "4" > gpurun_out/r2_pytest14.txt 2>&1
tail -6 gpurun_out/r2_pytest14.txt
