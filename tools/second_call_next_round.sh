#!/bin/bash
# Second GPU call of the next round (2 GPUs, charged 2x; ~10 minutes): the multi-GPU paths that have never run.
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/second_call_next_round.sh'
# Every torchrun sits under `timeout`; bench.py arms its own watchdog as well.
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_zz_static_exchange.py tests/test_gpu_zz_impala_exchange.py -m gpu -q -rxXs \
    -p no:cacheprovider > gpurun_out/next2_pytest.txt 2>&1
tail -15 gpurun_out/next2_pytest.txt
run() {   # name, flags
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
        bench.py --gpus 2 --steps 1000 --warmup 20 --watchdog-s 150 $2 > gpurun_out/next2_bench_$1.json 2> gpurun_out/next2_bench_$1.err
    python - "$1" <<'PY'
import json, sys
name = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/next2_bench_{name}.json").read().strip().splitlines()[-1])
    print(f"{name:20s} {j['ms_per_step']*1e3:8.1f} us/update  value {j['value']:.3e}  e2e {j['e2e']['value']:.3e}  launches/step {j['gpu_launches_per_step']}")
except Exception as e:
    print(name, "FAILED", e)
PY
}
run static            ""
run batched_draw      "--batched-draw"
run device_apportion  "--device-apportion --batched-draw"
run dynamic           "--dynamic-exchange"
run static_fused_rows "--net fused-rows"
run devapp_fused_rows "--net fused-rows --device-apportion"
# can the process group be torn down once the NCCL-capturing graph is released?  (blocked teardown -> timeout, rc 124)
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 \
    tools/nccl_graph_repro.py thread_local release > gpurun_out/next2_teardown.txt 2>&1; echo "teardown after release: rc=$?"; tail -3 gpurun_out/next2_teardown.txt

# C5's shape at the size of this call: 1 actor GPU -> 1 learner GPU (trajectory push + weight broadcast over NCCL)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 \
    tools/impala_dist_bench.py --steps 30 > gpurun_out/next2_impala.json 2> gpurun_out/next2_impala.err; tail -2 gpurun_out/next2_impala.json
# C4's shape at 2 shards (eager exchange)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29558 \
    tools/ddpg_apex_bench.py --rows 2000000 --steps 200 > gpurun_out/next2_ddpg.json 2> gpurun_out/next2_ddpg.err; tail -2 gpurun_out/next2_ddpg.json
