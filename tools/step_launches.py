#!/usr/bin/env python
"""One graph-replayed learner update between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum` (the launch list of ONE step)."""
import argparse
import os
import sys

import torch as t

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

sys.argv = [sys.argv[0], "--replay-rows", os.environ.get("ROWS", "2000000"), "--no-cpu-baseline", *sys.argv[1:]]
a = bench.parse()                      # bench.py's defaults (fused net, stream-level options); extra flags pass through
dev = t.device("cuda", 0)
learner = bench.build_learner(a, dev, seed=0)
learner.enable_cuda_graph()
for _ in range(5):
    learner.update()
t.cuda.synchronize()
t.cuda.cudart().cudaProfilerStart()
for _ in range(2):
    learner.update()
t.cuda.synchronize()
t.cuda.cudart().cudaProfilerStop()
print("done")
