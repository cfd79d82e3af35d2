"""CPU: dry runs of the measurement tools the next GPU session starts with (tools/ddpg_apex_bench.py eager and --graph,
tools/impala_bench.py): their main() executes line by line over the buffer doubles and the library recorder, so a typo
or a stale call in a tool does not cost a paid-for GPU call.  Numbers mean nothing here."""
import importlib.util
import json
import os
import sys
from threading import RLock

import pytest
import torch as t

from conftest import ROOT
from fake_ring import make_cpu_buffer
from test_bench_dry_run_cpu import _BenchRing, _BenchTree, _TimedEvent
from test_learner_dry_run_cpu import _dry_run_patches


class _TorchProxy:
    """`torch` as a tool sees it: every device it names is the host."""

    def __getattr__(self, name):
        return getattr(t, name)

    @staticmethod
    def device(kind="cpu", index=None):
        return t.device("cpu", 0 if index is None else index)

    @staticmethod
    def Generator(device=None):
        return t.Generator()


def _load(name):
    spec = importlib.util.spec_from_file_location("tool_" + name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.t = _TorchProxy()
    return mod


@pytest.fixture
def device_doubles(monkeypatch):
    import machin_b200.frame.algorithms.apex as apex_mod
    import machin_b200.frame.algorithms.ddpg_per as ddpg_mod
    import machin_b200.frame.algorithms.impala as impala_mod
    import machin_b200.frame.buffers as buffers_mod
    import machin_b200.parallel.group as group_mod
    rec = _dry_run_patches(monkeypatch)
    monkeypatch.setattr(t.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(t.cuda, "Event", _TimedEvent)
    monkeypatch.setattr(t.cuda, "is_current_stream_capturing", lambda: False)
    monkeypatch.setattr(t.nn.Module, "to", lambda self, *a, **k: self)
    dev = t.device("cpu", 0)
    real_per, real_dper, real_imp = (buffers_mod.PrioritizedBuffer, buffers_mod.DistributedPrioritizedBuffer,
                                     buffers_mod.IMPALABuffer)

    def finish(buf, size):
        buf.storage, buf.buffer_device = _BenchRing(size), dev
        return buf

    def per(buffer_size=1000, buffer_device=None, **kw):
        return finish(make_cpu_buffer(real_per, buffer_size, epsilon=1e-2, alpha=0.6, beta=0.4, curr_beta=0.4,
                                      beta_increment_per_sampling=1e-3, wt_tree=_BenchTree(buffer_size)), buffer_size)

    def dper(buffer_name, group, buffer_size, requesters=None, **kw):
        buf = finish(make_cpu_buffer(real_dper, buffer_size, epsilon=1e-2, alpha=0.6, beta=0.4, curr_beta=0.4,
                                     beta_increment_per_sampling=1e-3, wt_tree=_BenchTree(buffer_size), buffer_name=buffer_name,
                                     group=group, wr_lock=RLock(), _requesters=requesters, _sampler=None,
                                     buffer_version_table=t.zeros(buffer_size, dtype=t.int64)), buffer_size)
        buf.weight_sum_tensor = lambda: t.tensor([max(buf.wt_tree.get_weight_sum(), 1.0)], dtype=t.float64)
        return buf

    def imp(buffer_name, group, buffer_size, requesters=None):
        return make_cpu_buffer(real_imp, buffer_size, buffer_name=buffer_name, group=group, wr_lock=RLock(),
                               requesters=requesters or [0], _schema_agreed=False)

    monkeypatch.setattr(ddpg_mod, "PrioritizedBuffer", per)
    monkeypatch.setattr(apex_mod, "DistributedPrioritizedBuffer", dper)
    monkeypatch.setattr(impala_mod, "IMPALABuffer", imp)
    real_local = group_mod.LocalGroup.__init__

    def local_init(self, name="group", device=None):
        real_local(self, name, device="cpu")
        self.device = dev

    monkeypatch.setattr(group_mod.LocalGroup, "__init__", local_init)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    return rec


@pytest.mark.parametrize("flags", [[], ["--graph"]])
def test_ddpg_bench_tool_dry_run(flags, device_doubles, monkeypatch, capsys):
    tool = _load("ddpg_apex_bench")
    monkeypatch.setattr(tool, "S", 6)
    monkeypatch.setattr(tool, "A", 2)
    monkeypatch.setattr(tool, "H", 8)
    monkeypatch.setattr(sys, "argv", ["ddpg_apex_bench.py", "--rows", "600", "--batch", "8", "--steps", "2", "--warmup", "1", *flags])
    tool.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["launch"].startswith("one CUDA graph" if flags else "eager")
    assert {"mb_per_sample", "mb_td_target", "mb_tree_update_priority"} <= {name for name, _ in device_doubles.calls}


@pytest.mark.parametrize("flags", [[], ["--graph"]])
def test_impala_bench_tool_dry_run(flags, device_doubles, monkeypatch, capsys):
    import machin_b200.frame.algorithms.impala as impala_mod
    tool = _load("impala_bench")
    for name, value in (("S", 6), ("A", 3), ("T", 4), ("EP", 8)):
        monkeypatch.setattr(tool, name, value)
    monkeypatch.setattr(sys, "argv", ["impala_bench.py", *flags])
    monkeypatch.setattr(tool.t.optim, "Adam", t.optim.SGD)       # (torch's capturable Adam refuses host tensors)
    real_double = impala_mod.IMPALABuffer

    def with_device(*a, **k):
        buf = real_double(*a, **k)
        buf.buffer_device = t.device("cpu", 0)
        return buf

    monkeypatch.setattr(impala_mod, "IMPALABuffer", with_device)
    tool.main()
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert "C3 IMPALA" in line["config"] and line["gpu_update_ms"] > 0 and line["cpu_us_per_step"] > 0
    assert line["launch"] == ("one CUDA graph per update" if flags else "eager")
    assert "mb_vtrace_len" in {name for name, _ in device_doubles.calls}
