#!/bin/bash
# Round 2, GPU call (2 GPUs): the stream-level changes of the captured update (capture on the warm-up stream, last weight
# gradient on the main stream, beta advance + loss mean on a third branch) -- whole -m gpu suite (incl. world-2 tests),
# then N = 1 with the options on / all off, then N = 2.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -rfEs --tb=short -p no:cacheprovider --maxfail=6 > gpurun_out/r2_pytest17.txt 2>&1
tail -8 gpurun_out/r2_pytest17.txt
show() { python - "$1" <<'PY'
import json, sys
name = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/{name}.json").read().strip().splitlines()[-1])
    print(f"{name:26s} N={j['n_gpus']} {j['ms_per_step']*1e3:8.1f} us/step  value {j['value']:.3e}  e2e {j['e2e']['value']:.3e}  off={j['config'].get('step_options_off')}")
except Exception as e:
    print(name, "FAILED", e)
    try: print(open(f"gpurun_out/{name}.err").read()[-1500:])
    except Exception: pass
PY
}
timeout 150 python bench.py --steps 3000 --warmup 50 > gpurun_out/r2_n1_streams.json 2> gpurun_out/r2_n1_streams.err; show r2_n1_streams
MB_STEP_OFF=capture_stream,last_grad_main,third_branch timeout 120 python bench.py --steps 3000 --warmup 50 --no-cpu-baseline > gpurun_out/r2_n1_streams_off.json 2> gpurun_out/r2_n1_streams_off.err; show r2_n1_streams_off
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29781 bench.py --gpus 2 --steps 1000 --warmup 20 --watchdog-s 120 > gpurun_out/r2_n2_streams.json 2> gpurun_out/r2_n2_streams.err; show r2_n2_streams
