"""CPU: the parts of bench.py's contract that do not need a device -- the reference arm (`--impl reference`: the oracle
port of the reference learner on the host cores) prints exactly ONE JSON line with the agreed keys, ranks other than 0
leave without work, and the B200 arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run_bench(*flags, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True,
                          timeout=300, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--replay-rows", "50000", "--cpu-storage-rows", "2000")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 1 and j["steps"] == 2 and j["warmup"] >= 3
    for key in ("metric", "value", "unit", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["vs_baseline"] is None and j["higher_is_better"] is True and j["value"] > 0 and "workload" in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == j["value"]
    assert cb["unit"] == j["unit"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_of_the_sharded_ddpg_workload():
    """bench.py --workload c4 --impl reference: the oracle port of the reference's DDPG-PER update, same contract."""
    res = run_bench("--impl", "reference", "--workload", "c4", "--steps", "2", "--replay-rows", "20000", "--cpu-storage-rows", "1000")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "transitions_sampled_per_sec" and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and "C4" in j["config"]["workload"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_leave_without_work():
    res = run_bench("--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1",
                    env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == "", (res.stdout, res.stderr[-500:])


def test_b200_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        return                                            # (on the GPU box the arm runs: covered by the round-end bench)
    res = run_bench("--steps", "1", "--warmup", "1", "--replay-rows", "1000", "--no-cpu-baseline")
    assert res.returncode != 0 and res.stdout.strip() == ""
    assert "needs a CUDA device" in res.stderr
