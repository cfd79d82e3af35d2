#!/bin/bash
# usage: tools/gpu_retry.sh [--gpus N] TIMEOUT 'command'   -- retries gpurun while the pod answers "busy" (exit 3)
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun $G --timeout $T -- "$@" > /tmp/gpurun_last.txt 2>&1
    rc=$?
    if grep -q "status=transient" /tmp/gpurun_last.txt || [ $rc -eq 3 ]; then sleep 60; continue; fi
    break
done
cat /tmp/gpurun_last.txt
