#!/bin/bash
# Round 2, GPU call 8 (4 GPUs): the world-4 NCCL / peer parity tests, C2 at N = 4, C4 at its full size (50 M rows over 4 shards).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_zz_static_exchange.py tests/test_gpu_zz_impala_exchange.py -q -rA --tb=short -p no:cacheprovider > gpurun_out/r2_pytest8.txt 2>&1
tail -12 gpurun_out/r2_pytest8.txt
show() { python - "$1" <<'PY'
import json, sys
name = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/{name}.json").read().strip().splitlines()[-1])
    print(f"{name:26s} N={j['n_gpus']} {j['ms_per_step']*1e3:8.1f} us/step  value {j['value']:.3e}  e2e {j['e2e']['value']:.3e}  launches/step {j.get('gpu_launches_per_step')}  {j['config'].get('transport')}")
except Exception as e:
    print(name, "FAILED", e)
    try: print(open(f"gpurun_out/{name}.err").read()[-1500:])
    except Exception: pass
PY
}
tr() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
tr 200 29731 bench.py --gpus 4 --steps 1000 --warmup 20 --watchdog-s 150 > gpurun_out/r2_n4_default.json 2> gpurun_out/r2_n4_default.err; show r2_n4_default
tr 200 29732 bench.py --gpus 4 --steps 1000 --warmup 20 --watchdog-s 150 --nccl-exchange > gpurun_out/r2_n4_nccl.json 2> gpurun_out/r2_n4_nccl.err; show r2_n4_nccl
tr 400 29733 bench.py --gpus 4 --workload c4 --steps 500 --warmup 10 --watchdog-s 350 > gpurun_out/r2_c4_n4.json 2> gpurun_out/r2_c4_n4.err; show r2_c4_n4
tr 300 29734 bench.py --gpus 4 --workload c4 --replay-rows 2000000 --steps 500 --warmup 10 --watchdog-s 250 --net fused > gpurun_out/r2_c4_n4_fused_small.json 2> gpurun_out/r2_c4_n4_fused_small.err; show r2_c4_n4_fused_small
