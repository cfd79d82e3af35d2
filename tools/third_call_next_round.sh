#!/bin/bash
# Third GPU call of the next round: the sizes BASELINE.json names that were never run -- C2 at 4 and 8 GPUs, C4 at 4
# shards, C5 at 8 GPUs.  Charged N x, so each run is short; pass the GPU count of the call as $1 (4 or 8):
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/third_call_next_round.sh 8'
#   gpurun --gpus 4 --timeout 600 -- 'bash tools/third_call_next_round.sh 4'
# Every torchrun sits under `timeout`; bench.py arms its own watchdog as well.
N=${1:-4}
mkdir -p gpurun_out
launch() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for variant in "" "--dynamic-exchange"; do
    name=static; [ -n "$variant" ] && name=dynamic
    launch 240 29561 bench.py --gpus $N --steps 1000 --warmup 20 --watchdog-s 200 $variant \
        > gpurun_out/next3_bench_n${N}_$name.json 2> gpurun_out/next3_bench_n${N}_$name.err
    tail -c 700 gpurun_out/next3_bench_n${N}_$name.json; echo
done
if [ "$N" = 4 ]; then      # C4: 50 M rows over 4 shards (12.5 M x 3104 B = 38.8 GB per GPU)
    launch 420 29562 tools/ddpg_apex_bench.py --rows 12500000 --steps 200 > gpurun_out/next3_ddpg_n4.json 2> gpurun_out/next3_ddpg_n4.err
    tail -2 gpurun_out/next3_ddpg_n4.json
fi
if [ "$N" = 8 ]; then      # C5: 7 actor GPUs -> 1 learner GPU, 4096 episodes per batch
    launch 300 29563 tools/impala_dist_bench.py --steps 30 > gpurun_out/next3_impala_n8.json 2> gpurun_out/next3_impala_n8.err
    tail -2 gpurun_out/next3_impala_n8.json
fi
