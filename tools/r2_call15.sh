#!/bin/bash
# Round 2, GPU call (2 GPUs): the whole -m gpu suite on the final code (incl. the world-2 tests), N = 1 and N = 2 bench lines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rfEs --tb=short -p no:cacheprovider > gpurun_out/r2_pytest15.txt 2>&1
tail -6 gpurun_out/r2_pytest15.txt
show() { python - "$1" <<'PY'
import json, sys
name = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/{name}.json").read().strip().splitlines()[-1])
    print(f"{name:26s} N={j['n_gpus']} {j['ms_per_step']*1e3:8.1f} us/step  value {j['value']:.3e}  e2e {j['e2e']['value']:.3e}  launches/step {j.get('gpu_launches_per_step')}")
    for k, v in j.get("kernels", {}).items():
        if "gemm forward" in k: print(f"    {k:55s} {v['avg_us']:.2f} us x{v['per_step']}")
except Exception as e:
    print(name, "FAILED", e)
    try: print(open(f"gpurun_out/{name}.err").read()[-1500:])
    except Exception: pass
PY
}
timeout 200 python bench.py --steps 3000 --warmup 50 --no-cpu-baseline > gpurun_out/r2_n1_last.json 2> gpurun_out/r2_n1_last.err; show r2_n1_last
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29771 bench.py --gpus 2 --steps 1000 --warmup 20 --watchdog-s 150 > gpurun_out/r2_n2_last.json 2> gpurun_out/r2_n2_last.err; show r2_n2_last
