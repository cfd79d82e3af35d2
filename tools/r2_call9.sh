#!/bin/bash
# Round 2, GPU call 9 (8 GPUs): world-4 parity tests, C2 at N = 8 (peer and NCCL transport), C5 (7 actors -> 1 learner).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_zz_static_exchange.py tests/test_gpu_zz_impala_exchange.py -q -rA --tb=short -p no:cacheprovider -k "4" > gpurun_out/r2_pytest9.txt 2>&1
tail -8 gpurun_out/r2_pytest9.txt
show() { python - "$1" <<'PY'
import json, sys
name = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/{name}.json").read().strip().splitlines()[-1])
    print(f"{name:26s} N={j['n_gpus']} {j['ms_per_step']*1e3:8.1f} us/step  value {j['value']:.3e}  e2e {j['e2e']['value']:.3e}  launches/step {j.get('gpu_launches_per_step')}  {j['config'].get('transport')} {j.get('link')}")
except Exception as e:
    print(name, "FAILED", e)
    try: print(open(f"gpurun_out/{name}.err").read()[-1500:])
    except Exception: pass
PY
}
tr() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
tr 200 29741 bench.py --gpus 8 --steps 1000 --warmup 20 --watchdog-s 150 > gpurun_out/r2_n8_default.json 2> gpurun_out/r2_n8_default.err; show r2_n8_default
tr 200 29742 bench.py --gpus 8 --steps 1000 --warmup 20 --watchdog-s 150 --nccl-exchange > gpurun_out/r2_n8_nccl.json 2> gpurun_out/r2_n8_nccl.err; show r2_n8_nccl
tr 300 29743 bench.py --gpus 8 --workload c5 --steps 30 --warmup 3 --watchdog-s 250 > gpurun_out/r2_c5_n8.json 2> gpurun_out/r2_c5_n8.err; show r2_c5_n8
