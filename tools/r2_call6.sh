#!/bin/bash
# Round 2, GPU call 6 (2 GPUs): the peer-memory exchange over NVLink -- parity against the NCCL fixed-shape update, then timing.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_static_exchange.py tests/test_gpu_algorithms.py -q -rfEs --tb=short -p no:cacheprovider -k "static or fused" > gpurun_out/r2_pytest6.txt 2>&1
tail -25 gpurun_out/r2_pytest6.txt
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
tr() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
tr 200 29711 bench.py --gpus 2 --steps 1000 --warmup 20 --watchdog-s 150 > gpurun_out/r2_n2_nccl.json 2> gpurun_out/r2_n2_nccl.err; show r2_n2_nccl
tr 200 29712 bench.py --gpus 2 --steps 1000 --warmup 20 --watchdog-s 150 --peer-exchange > gpurun_out/r2_n2_peer.json 2> gpurun_out/r2_n2_peer.err; show r2_n2_peer
tr 200 29713 bench.py --gpus 2 --steps 1000 --warmup 20 --watchdog-s 150 --peer-exchange --net fused > gpurun_out/r2_n2_peer_fused.json 2> gpurun_out/r2_n2_peer_fused.err; show r2_n2_peer_fused
timeout 200 python bench.py --steps 3000 --warmup 50 --no-cpu-baseline --net fused > gpurun_out/r2_n1_fused.json 2> gpurun_out/r2_n1_fused.err; show r2_n1_fused
