#!/bin/bash
# Round 2, GPU call 7 (2 GPUs): sharded DDPG fixed-shape update (NCCL + peer transport) parity, C4 through bench.py.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_static_exchange.py tests/test_gpu_zz_ddpg_graph.py tests/test_gpu_algorithms.py -q -rfEs --tb=short -p no:cacheprovider > gpurun_out/r2_pytest7.txt 2>&1
tail -15 gpurun_out/r2_pytest7.txt
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
tr 240 29721 bench.py --gpus 2 --workload c4 --replay-rows 2000000 --steps 500 --warmup 10 --watchdog-s 200 > gpurun_out/r2_c4_n2_graph.json 2> gpurun_out/r2_c4_n2_graph.err; show r2_c4_n2_graph
tr 240 29722 bench.py --gpus 2 --workload c4 --replay-rows 2000000 --steps 500 --warmup 10 --watchdog-s 200 --nccl-exchange > gpurun_out/r2_c4_n2_graph_nccl.json 2> gpurun_out/r2_c4_n2_graph_nccl.err; show r2_c4_n2_graph_nccl
tr 200 29723 bench.py --gpus 2 --steps 1000 --warmup 20 --watchdog-s 150 > gpurun_out/r2_n2_default.json 2> gpurun_out/r2_n2_default.err; show r2_n2_default
timeout 240 python bench.py --workload c4 --replay-rows 2000000 --steps 500 --warmup 10 --no-cpu-baseline > gpurun_out/r2_c4_n1_small.json 2> gpurun_out/r2_c4_n1_small.err; show r2_c4_n1_small
