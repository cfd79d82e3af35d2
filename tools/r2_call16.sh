#!/bin/bash
# Round 2, GPU call (1 GPU): the machin_b200.auto launcher tests on real buffers, and the launch command end to end.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_zz_auto.py -m gpu -q -rfEs --tb=short -p no:cacheprovider > gpurun_out/r2_pytest16.txt 2>&1
tail -40 gpurun_out/r2_pytest16.txt
timeout 120 python -m machin_b200.auto generate --algo DQNPer --env classic_control --output gpurun_out/r2_auto_cfg.json
python - <<'PY'
import json
c = json.load(open("gpurun_out/r2_auto_cfg.json"))
c["frame_config"].update(models=["machin_b200.model.FusedMLP"] * 2, model_kwargs=[{"sizes": [4, 64, 64, 2]}] * 2, batch_size=64, replay_size=100000)
c.update(root_dir="gpurun_out/r2_auto_trial", episode_per_epoch=5, max_episodes=15, cuda_graph=True)
json.dump(c, open("gpurun_out/r2_auto_cfg.json", "w"), indent=1)
PY
( time timeout 200 python -m machin_b200.auto launch --config gpurun_out/r2_auto_cfg.json ) > gpurun_out/r2_auto_launch.log 2>&1
tail -5 gpurun_out/r2_auto_launch.log
cat gpurun_out/r2_auto_trial/log/train_log/metrics.jsonl | tail -15
ls gpurun_out/r2_auto_trial/model
rm -rf gpurun_out/r2_auto_trial/model
