#!/bin/bash
# Round 2, GPU call 4 (2 GPUs): suite incl. world-2 NCCL tests (new server cases, bulk gather), N=2 variants of the C2
# step, sanity runs of --workload c3/c4/c5, NCCL graph teardown with release.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rfEs --tb=short -p no:cacheprovider > gpurun_out/r2_pytest4.txt 2>&1
tail -12 gpurun_out/r2_pytest4.txt
show() { python - "$1" <<'PY'
import json, sys
name = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/{name}.json").read().strip().splitlines()[-1])
    r = j.get("roofline") or {}
    print(f"{name:30s} N={j['n_gpus']} {j['ms_per_step']*1e3:9.1f} us/step  value {j['value']:.3e}  e2e {j['e2e']['value']:.3e}  "
          f"launches/step {j.get('gpu_launches_per_step')}  dom {r.get('kernel')} {r.get('avg_launch_us')} frac {r.get('frac')}")
except Exception as e:
    print(name, "FAILED", e)
    try: print(open(f"gpurun_out/{name}.err").read()[-1500:])
    except Exception: pass
PY
}
tr() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
timeout 200 python bench.py --steps 3000 --warmup 50 --no-cpu-baseline > gpurun_out/r2_n1_default.json 2> gpurun_out/r2_n1_default.err; show r2_n1_default
tr 200 29701 bench.py --gpus 2 --steps 1000 --warmup 20 --watchdog-s 150 --serial > gpurun_out/r2_n2_serial.json 2> gpurun_out/r2_n2_serial.err; show r2_n2_serial
tr 200 29702 bench.py --gpus 2 --steps 1000 --warmup 20 --watchdog-s 150 > gpurun_out/r2_n2_streams.json 2> gpurun_out/r2_n2_streams.err; show r2_n2_streams
tr 200 29703 bench.py --gpus 2 --steps 1000 --warmup 20 --watchdog-s 150 --device-apportion --batched-draw > gpurun_out/r2_n2_devapp.json 2> gpurun_out/r2_n2_devapp.err; show r2_n2_devapp
timeout 300 python bench.py --workload c3 --steps 200 --warmup 5 > gpurun_out/r2_c3.json 2> gpurun_out/r2_c3.err; show r2_c3
tr 240 29704 bench.py --gpus 2 --workload c4 --replay-rows 2000000 --steps 200 --warmup 5 --watchdog-s 200 > gpurun_out/r2_c4_n2_small.json 2> gpurun_out/r2_c4_n2_small.err; show r2_c4_n2_small
tr 240 29705 bench.py --gpus 2 --workload c5 --steps 20 --warmup 3 --watchdog-s 200 > gpurun_out/r2_c5_n2.json 2> gpurun_out/r2_c5_n2.err; show r2_c5_n2
tr 90 29706 tools/nccl_graph_repro.py thread_local release > gpurun_out/r2_nccl_release.log 2>&1; echo "release rc=$?"; tail -4 gpurun_out/r2_nccl_release.log
