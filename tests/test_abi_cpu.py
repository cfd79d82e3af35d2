"""CPU: the C-ABI library loads and exports every symbol declared in include/machin_b200.h; handle
constructors fail loudly without a GPU (no CPU fallback); the oracle is not imported by the product."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "machin_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from machin_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 35
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert set(_lib.SIGNATURES) == set(names), set(names) ^ set(_lib.SIGNATURES)
    assert lib.mb_abi_version() == 1
    lib.mb_status_string.restype = ctypes.c_char_p
    assert lib.mb_status_string(-2) == b"Index has elements out of boundary!"
    assert lib.mb_status_string(-3) == b"Dimension of weights and indexes doesn't match!"
    assert lib.mb_status_string(-4) == b"Weights size must match tree size!"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from machin_b200.frame.buffers import PrioritizedBuffer, WeightTree
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        WeightTree(5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PrioritizedBuffer(buffer_size=5)
    with pytest.raises(ValueError, match="device-resident"):
        PrioritizedBuffer(buffer_size=5, buffer_device="cpu")
    from machin_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.td_target(torch.zeros(2, 1), torch.zeros(2, 1), torch.zeros(2, 1, dtype=torch.bool), 0.9)


def test_product_never_imports_the_oracle():
    out = subprocess.run([sys.executable, "-c",
                          "import sys; import machin_b200, machin_b200.ops, machin_b200.frame.buffers, "
                          "machin_b200.frame.algorithms, machin_b200.parallel.server; "
                          "print(sorted(m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')))"],
                         cwd=ROOT, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "[]"
    for dirpath, _, files in os.walk(os.path.join(ROOT, "machin_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)


def test_host_logic_ring_layout_and_transition():
    import torch as t
    from machin_b200.frame.buffers.device_ring import Field, layout
    from machin_b200.frame.transition import Transition
    fields = [Field("state/state", "state", "state", t.float32, (84,), "major"),
              Field("action/action", "action", "action", t.int64, (1,), "major"),
              Field("next_state/state", "next_state", "state", t.float32, (84,), "major"),
              Field("reward", "reward", None, t.float32, (1,), "sub"),
              Field("terminal", "terminal", None, t.bool, (1,), "sub")]
    pitch = layout(fields)
    assert pitch == 688 and sum(f.nbytes for f in fields) == 685         # 685 B per C2 row
    assert all(f.offset % 16 == 0 for f in fields if f.nbytes >= 16)
    spans = sorted((f.offset, f.offset + f.nbytes) for f in fields)
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    tr = Transition(state={"s": t.zeros(1, 3)}, action={"a": t.zeros(1, 1)}, next_state={"s": t.zeros(1, 3)},
                    reward=1.0, terminal=False, zeta=1, alpha=2)
    assert tr.keys() == ["state", "action", "next_state", "reward", "terminal", "alpha", "zeta"]
    with pytest.raises(ValueError):
        Transition(state={"s": t.zeros(2, 3)}, action={"a": t.zeros(2, 1)}, next_state={"s": t.zeros(2, 3)},
                   reward=1.0, terminal=False)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_episode_book_vs_reference_buffer(tag):
    """Slot assignment, whole-episode eviction and sampling order of the host-side bookkeeping against the
    unmodified reference Buffer (tests/golden/ref_buffer.npz; buffer.py:46-109,224-252, storage.py:85-114)."""
    import random
    import numpy as np
    from conftest import GOLDEN
    from machin_b200.frame.buffers import EpisodeBook
    g = np.load(os.path.join(GOLDEN, "ref_buffer.npz"))
    cap = int(g[f"{tag}_cap"])
    book = EpisodeBook(cap)
    for i, n in enumerate(g[f"{tag}_lens"]):
        first, pos = book.assign(int(n))
        book.record(pos)
        keys = book.sampleable()
        want = [int(k) for k in g[f"{tag}_keys"][i] if k >= 0]
        assert keys == want, (i, keys, want)
        assert book.count == int(g[f"{tag}_size"][i])
        random.seed(5)
        got = random.sample(keys, k=min(len(keys), 4))
        assert got == [int(k) for k in g[f"{tag}_pick_unique"][i] if k >= 0]
        random.seed(6)
        got = random.choices(keys, k=min(len(keys), 4))
        assert got == [int(k) for k in g[f"{tag}_pick_random"][i] if k >= 0]
    with pytest.raises(ValueError, match="Failed to store episode"):
        book.assign(cap + 1)


def test_distributed_default_priorities_vs_reference():
    """Compounding default priorities + taint counters of the sharded buffer against the unmodified reference
    (tests/golden/ref_per_d.npz; prioritized_buffer_d.py:123-137, 282-291), replayed on the oracle tree."""
    import numpy as np
    from conftest import GOLDEN
    from machin_b200.frame.buffers import EpisodeBook
    from machin_b200.frame.buffers.prioritized_buffer_d import compounding_default_weights
    from oracle import per as oper
    from oracle.sumtree import OracleTree
    g = np.load(os.path.join(GOLDEN, "ref_per_d.npz"))
    cap = g["versions"].shape[1]
    tree, book, version = OracleTree(cap), EpisodeBook(cap), np.zeros(cap, dtype=np.int64)
    for step, n in enumerate(g["lens"]):
        _, pos = book.assign(int(n))
        book.record(pos)
        w = compounding_default_weights(tree.get_leaf_max(), int(n), 1e-2, 0.6)
        tree.update_leaf_batch(w, pos)
        version[pos] += 1
        assert np.array_equal(tree.weights, g["trees"][step]), step
        assert np.array_equal(version, g["versions"][step]), step
        if step == 1:
            tree.update_leaf_batch(oper.normalize_priority(g["wb_pri"]), g["wb_idx"])
            assert np.array_equal(tree.weights, g["tree_after_wb"])
            assert isinstance(tree.get_leaf_max(), np.float32)      # the float32 write-back now owns max_leaf


def test_entry_points_reject_bad_arguments_before_touching_the_device():
    """Argument validation happens before any CUDA call, so it can be exercised without a GPU: NULL handles,
    negative counts, unsupported dtypes -> MB_ERR_INVALID (mapped to RuntimeError by the shim)."""
    import ctypes as C
    from machin_b200 import _lib
    from machin_b200._lib import lib
    null = C.c_void_p(0)
    one = (C.c_int64 * 1)(0)
    cases = [
        lib.mb_tree_update_batch(null, one, one, _lib.MB_F64, null, 1, null),
        lib.mb_tree_update_priority(null, one, one, _lib.MB_F32, null, 1, 1e-2, 0.6, null),
        lib.mb_tree_find(null, one, 1, one, null),
        lib.mb_per_sample(null, one, 1, 1, 0.0, 0.4, null, one, one, null),
        lib.mb_rows_gather(null, 16, 1, one, 1, None, 0, null),
        lib.mb_wire_compact(null, 1, 1, 40, 16, one, one, 2, one, one, one, null, null),
        lib.mb_wire_apportion(null, 1, 1, 0, 2, one, one, 1e-3, one, one, one, null),
        lib.mb_wire_apportion(one, 2, 1, 5, 3, one, one, 1e-3, one, one, one, null),      # self_block out of range
        lib.mb_mlp_program_run(null, null),
        lib.mb_mlp_program_run_part(null, 0, null),
    ]
    assert all(c == _lib.MB_ERR_INVALID for c in cases), cases
    with pytest.raises(RuntimeError, match="libmachin_b200"):
        _lib.check(_lib.MB_ERR_INVALID)
    for status, message in ((_lib.MB_ERR_INDEX_OOB, "Index has elements out of boundary!"),
                            (_lib.MB_ERR_DIM_MISMATCH, "Dimension of weights and indexes doesn't match!"),
                            (_lib.MB_ERR_SIZE_MISMATCH, "Weights size must match tree size!")):
        with pytest.raises(ValueError, match=message):          # the reference's own messages
            _lib.check(status)


def test_ctypes_signatures_match_the_header_prototypes():
    """Argument count and argument classes (pointer / int / int64 / float / double) of every ctypes signature in
    machin_b200/_lib.py against the prototype in include/machin_b200.h: ctypes converts silently, a swapped or missing
    argument would only show up as garbage on the device."""
    import ctypes as C
    import re
    from machin_b200 import _lib
    header = open(os.path.join(ROOT, "include", "machin_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = dict(re.findall(r"\b(?:int|long long|const char \*|const double \*|double \*|void \*|int64_t)\s*(mb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;",
                             header, flags=re.S))

    def c_class(arg):
        arg = " ".join(arg.split())
        if arg in ("void", ""):
            return None
        if "*" in arg:
            return "ptr"
        if re.search(r"\bint64_t\b", arg):
            return "i64"
        if re.search(r"\bdouble\b", arg):
            return "f64"
        if re.search(r"\bfloat\b", arg):
            return "f32"
        if re.search(r"\b(int|unsigned|mb_dtype)\b", arg):
            return "i32"
        raise AssertionError(f"unclassified C argument: {arg!r}")

    def py_class(tp):
        if tp in (C.c_void_p, C.c_char_p) or (isinstance(tp, type) and issubclass(tp, C._Pointer)):
            return "ptr"
        return {C.c_int64: "i64", C.c_double: "f64", C.c_float: "f32", C.c_int: "i32", C.c_int32: "i32"}[tp]

    assert len(protos) >= 40
    bad = []
    for name, (_res, args) in _lib.SIGNATURES.items():
        assert name in protos, name
        want = [c for c in (c_class(a) for a in protos[name].split(",")) if c is not None]
        got = [py_class(a) for a in args]
        if want != got:
            bad.append((name, want, got))
    assert not bad, bad
