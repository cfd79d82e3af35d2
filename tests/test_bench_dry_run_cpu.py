"""CPU: a DRY RUN of bench.py's B200 arm at N = 1 -- argument parsing, learner construction, the ring fill, graph
capture, the device-timed loop, the end-to-end loop, the per-kernel timing harness and the JSON line -- with CUDA
streams / events / graphs as no-op stand-ins, the ring and the tree as host-memory doubles and the library as the
signature-checking recorder (tests/test_learner_dry_run_cpu.py).  No number means anything; what is checked is that
every Python line of the arm executes and that the line carries the contract's keys, for the default flags and for
the opt-in variants the next GPU session will time."""
import contextlib
import importlib.util
import json
import os
import sys

import pytest
import torch as t

from conftest import ROOT
from fake_ring import make_cpu_buffer
from test_learner_dry_run_cpu import _Ring, _Tree, _dry_run_patches


class _BenchTree(_Tree):
    depth = property(lambda self: self.ora.depth)

    def check(self):
        pass


class _BenchRing(_Ring):
    """(bench.py's per-kernel timing harness builds the gather descriptor itself)"""
    from machin_b200.frame.buffers.device_ring import DeviceRing
    _field_array = DeviceRing._field_array
    del DeviceRing


class _TimedEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return 1.0


def _load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("flags", [[], ["--no-graph"], ["--net", "fused"], ["--net", "fused-rows", "--overlap-update", "--fuse-optimizer"],
                                   ["--parallel-forwards", "--overlap-update"], ["--adam", "torch"]])
def test_b200_arm_dry_run(flags, monkeypatch):
    import machin_b200.frame.buffers as buffers_mod
    bench = _load_bench()
    rec = _dry_run_patches(monkeypatch)
    monkeypatch.setattr(bench, "cuda_device", lambda index: t.device("cpu", index))
    monkeypatch.setattr(t.cuda, "is_available", lambda: True)
    monkeypatch.setattr(t.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(t.cuda, "Event", _TimedEvent)
    monkeypatch.setattr(t.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(t.cuda, "is_current_stream_capturing", lambda: False)
    monkeypatch.setattr(t.nn.Module, "cuda", lambda self, device=None: self)
    real_per = buffers_mod.PrioritizedBuffer

    def per_double(buffer_size, buffer_device):
        buf = make_cpu_buffer(real_per, buffer_size, epsilon=1e-2, alpha=0.6, beta=0.4, curr_beta=0.4,
                              beta_increment_per_sampling=1e-3, wt_tree=_BenchTree(buffer_size))
        buf.storage = _BenchRing(buffer_size)
        buf.buffer_device = buffer_device
        return buf

    monkeypatch.setattr(buffers_mod, "PrioritizedBuffer", per_double)
    written = []

    class OsProxy:                                       # bench.py keeps a private handle to fd 1 and points fd 1 at stderr:
        dup = staticmethod(lambda fd: 99)                # not under pytest's own capture
        dup2 = staticmethod(lambda a, b: None)
        write = staticmethod(lambda fd, data: written.append(data) if fd == 99 else os.write(fd, data))

        def __getattr__(self, name):
            return getattr(os, name)

    monkeypatch.setattr(bench, "os", OsProxy())
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "3", "--warmup", "1", "--replay-rows", "3000", "--batch", "16",
                                      "--no-cpu-baseline", *flags])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    bench.run_b200(bench.parse())
    assert len(written) == 1
    line = json.loads(written[0].decode())
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["warmup"] == 3 and line["vs_baseline"] is None
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["e2e"]["h2d_bytes_per_step"] == 8 * 16 and "workload" in line["config"]
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    called = {name for name, _ in rec.calls}
    assert "mb_per_sample" in called and "mb_tree_update_priority" in called


class _Left(Exception):
    """bench.py left through os._exit (what the fixed-shape run does instead of the NCCL teardown)."""


@pytest.mark.parametrize("case", ["static", "dynamic", "static_capture_fails"])
def test_b200_arm_dry_run_two_ranks(case, monkeypatch):
    """The N > 1 arm as rank 0 of a 2-process job: process-group calls are no-ops and the shard group is a group of
    one, so the exchange moves nothing -- but bench.py's own N > 1 code (learner construction, the fixed-shape default,
    the fall-back to the dynamic exchange when the capture raises, push_weights in the loop, the max-over-ranks
    reduction, leaving without the teardown) executes line by line."""
    import torch.distributed as dist
    import machin_b200.frame.algorithms.apex as apex_mod
    import machin_b200.parallel as parallel_mod
    from machin_b200.frame.buffers import DistributedPrioritizedBuffer
    from threading import RLock
    bench = _load_bench()
    rec = _dry_run_patches(monkeypatch)
    dev = t.device("cpu", 0)
    monkeypatch.setattr(bench, "cuda_device", lambda index: dev)
    monkeypatch.setattr(t.cuda, "is_available", lambda: True)
    monkeypatch.setattr(t.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(t.cuda, "Event", _TimedEvent)
    monkeypatch.setattr(t.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(t.cuda, "is_current_stream_capturing", lambda: False)
    monkeypatch.setattr(t.nn.Module, "cuda", lambda self, device=None: self)
    for name in ("init_process_group", "barrier", "destroy_process_group", "all_reduce"):
        monkeypatch.setattr(dist, name, lambda *a, **k: None)

    def group_of_one(name, device=None):
        g = parallel_mod.LocalGroup(name, device="cpu")
        g.device = dev
        return g

    monkeypatch.setattr(parallel_mod, "ShardGroup", group_of_one)

    def buffer_double(buffer_name, group, buffer_size, requesters=None, **kw):
        buf = make_cpu_buffer(DistributedPrioritizedBuffer, buffer_size, epsilon=1e-2, alpha=0.6, beta=0.4, curr_beta=0.4,
                              beta_increment_per_sampling=1e-3, wt_tree=_BenchTree(buffer_size), buffer_name=buffer_name,
                              group=group, wr_lock=RLock(), _requesters=requesters, _sampler=None,
                              buffer_version_table=t.zeros(buffer_size, dtype=t.int64))
        buf.storage = _BenchRing(buffer_size)
        buf.buffer_device = dev
        buf.weight_sum_tensor = lambda: t.tensor([max(buf.wt_tree.get_weight_sum(), 1.0)], dtype=t.float64)
        return buf

    monkeypatch.setattr(apex_mod, "DistributedPrioritizedBuffer", buffer_double)
    if case == "static_capture_fails":
        real_enable, calls = apex_mod.DQNApex.enable_cuda_graph, []

        def enable(self, warmup=3, static_exchange=False, device_apportion=False, peer_exchange=False):
            calls.append(static_exchange)
            if static_exchange:
                raise RuntimeError("capture refused (injected)")
            return real_enable(self, warmup)

        monkeypatch.setattr(apex_mod.DQNApex, "enable_cuda_graph", enable)
    written, left = [], []

    class OsProxy:
        dup = staticmethod(lambda fd: 99)
        dup2 = staticmethod(lambda a, b: None)
        write = staticmethod(lambda fd, data: written.append(data) if fd == 99 else os.write(fd, data))

        @staticmethod
        def _exit(code):
            left.append(code)
            raise _Left()

        def __getattr__(self, name):
            return getattr(os, name)

    monkeypatch.setattr(bench, "os", OsProxy())
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "2", "--steps", "3", "--warmup", "1", "--replay-rows", "3000",
                                      "--batch", "16", "--watchdog-s", "600",
                                      *(["--dynamic-exchange"] if case == "dynamic" else [])])
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setenv("WORLD_SIZE", "2")
    import threading
    timers = []
    real_timer = threading.Timer
    monkeypatch.setattr(threading, "Timer", lambda *a, **k: timers.append(real_timer(*a, **k)) or timers[-1])
    try:
        if case != "static_capture_fails":
            bench.run_b200(bench.parse())                       # releases its graphs, then destroy_process_group (a no-op here)
            assert left == []
        else:
            with pytest.raises(_Left):                          # a half-captured NCCL graph may linger: no teardown
                bench.run_b200(bench.parse())
            assert left == [0]
    finally:
        for tm in timers:
            tm.cancel()                                         # the watchdog must not outlive the test
    assert len(written) == 1
    line = json.loads(written[0].decode())
    assert line["n_gpus"] == 2 and line["cpu_baseline"] is None
    assert line["roofline"] is not None and line["kernels"]          # rank 0 times its shard's kernels at N > 1 too
    # the default fixed-shape update apportions on the device: only the loss comes back; the dynamic exchange (and the
    # fall-back to it) reads the P shard sums on the host
    assert line["e2e"]["d2h_bytes_per_step"] == (8 if case == "static" else 8 + 8 * 2)
    launch = line["config"]["launch"]
    if case == "static":
        assert launch == "one CUDA graph per update"
    else:
        assert launch.startswith("two CUDA graphs")
    if case == "static_capture_fails":
        assert calls == [True, False]
    called = {name for name, _ in rec.calls}
    assert ("mb_wire_compact" in called) == (case == "static")
