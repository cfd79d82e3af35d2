"""GPU parity: the bulk-copy (TMA, cp.async.bulk) gather against the one-warp-per-row gather and a plain torch
indexing reference -- byte-exact -- through the C ABI (mb_rows_gather), for the C2 and C4 row layouts, narrow and odd
fields, every tile size the launcher picks (n from 1 row to several tiles per CTA), out-of-range clamping and the
identity index."""
import ctypes as C

import numpy as np
import pytest
import torch as t

pytestmark = pytest.mark.gpu


def _gather(ring, pitch, cap, index, n, fields):
    """fields: list of (row_offset, nbytes, dense tensor u8[n, dense_pitch])."""
    from machin_b200 import _lib
    arr = (_lib.MbField * len(fields))()
    for a, (off, nb, dense) in zip(arr, fields):
        a.row_offset, a.nbytes, a.d_dense, a.dense_pitch = off, nb, dense.data_ptr(), dense.shape[1]
    _lib.check(_lib.lib.mb_rows_gather(_lib.ptr(ring), pitch, cap, _lib.ptr(index), n, arr, len(fields), _lib.stream_ptr()))
    t.cuda.synchronize()


LAYOUTS = {   # name -> (pitch, [(row_offset, nbytes, dense_pitch)])
    "c2": (688, [(0, 336, 336), (336, 336, 336), (672, 8, 8), (680, 4, 4), (684, 1, 1)]),
    "c4": (3104, [(0, 1504, 1504), (1504, 1504, 1504), (3008, 68, 68), (3076, 4, 4), (3080, 1, 1)]),
    "packed_row": (688, [(0, 688, 688)]),                       # gather_packed: one field = the whole row
    "wide_unaligned_dest": (688, [(0, 336, 344), (336, 336, 336), (672, 16, 16)]),   # 344 % 16 != 0: not bulk-able
    "tiny": (16, [(0, 16, 16)]),
    "impala": (2112, [(0, 1024, 1024), (1024, 1024, 1024), (2048, 8, 8), (2056, 8, 8), (2064, 4, 4), (2068, 4, 4), (2072, 1, 1)]),
}


@pytest.mark.parametrize("layout", list(LAYOUTS))
@pytest.mark.parametrize("n", [1, 7, 512, 4099, 100_000])
def test_bulk_gather_matches_legacy_and_reference(layout, n, monkeypatch):
    pitch, fields = LAYOUTS[layout]
    cap = 5000
    g = t.Generator(device="cuda").manual_seed(n)
    ring = t.randint(0, 256, (cap, pitch), dtype=t.uint8, device="cuda", generator=g)
    index = t.randint(0, cap, (n,), device="cuda", generator=g)
    if n >= 7:
        index[3], index[5] = -4, cap + 9                        # clamped to the ring like rows_move_kernel does
    ref_rows = ring[index.clamp(0, cap - 1)]
    outs = {}
    for mode in ("bulk", "legacy"):
        monkeypatch.setenv("MB_ROWS_LEGACY", "1" if mode == "legacy" else "0")
        dense = [t.full((n, dp), 0xEE, dtype=t.uint8, device="cuda") for _, _, dp in fields]
        _gather(ring, pitch, cap, index, n, [(off, nb, d) for (off, nb, _), d in zip(fields, dense)])
        outs[mode] = dense
        for (off, nb, dp), d in zip(fields, dense):
            assert t.equal(d[:, :nb], ref_rows[:, off:off + nb]), (mode, off)
            if dp > nb:
                assert bool((d[:, nb:] == 0xEE).all()), "wrote outside the field"
    for a, b in zip(outs["bulk"], outs["legacy"]):
        assert t.equal(a, b)


def test_bulk_gather_identity_index_and_graph_replay(monkeypatch):
    """NULL index = identity (unpacking rows received from a peer); the launch is capturable and replays on new data."""
    from machin_b200 import _lib
    monkeypatch.setenv("MB_ROWS_LEGACY", "0")
    pitch, n = 688, 600
    ring = t.randint(0, 256, (n, pitch), dtype=t.uint8, device="cuda")
    a, b = t.zeros(n, 336, dtype=t.uint8, device="cuda"), t.zeros(n, 336, dtype=t.uint8, device="cuda")
    _gather(ring, pitch, n, None, n, [(0, 336, a), (336, 336, b)])
    assert t.equal(a, ring[:, :336]) and t.equal(b, ring[:, 336:672])
    index = t.randint(0, n, (n,), device="cuda")
    arr = (_lib.MbField * 2)()
    for x, (off, d) in zip(arr, ((0, a), (336, b))):
        x.row_offset, x.nbytes, x.d_dense, x.dense_pitch = off, 336, d.data_ptr(), 336
    side = t.cuda.Stream()
    side.wait_stream(t.cuda.current_stream())
    with t.cuda.stream(side):
        _lib.check(_lib.lib.mb_rows_gather(_lib.ptr(ring), pitch, n, _lib.ptr(index), n, arr, 2, _lib.stream_ptr()))
    t.cuda.current_stream().wait_stream(side)
    graph = t.cuda.CUDAGraph()
    with t.cuda.graph(graph):
        _lib.check(_lib.lib.mb_rows_gather(_lib.ptr(ring), pitch, n, _lib.ptr(index), n, arr, 2, _lib.stream_ptr()))
    for seed in range(3):
        index.copy_(t.randint(0, n, (n,), device="cuda"))
        graph.replay()
        t.cuda.synchronize()
        assert t.equal(a, ring[index][:, :336]) and t.equal(b, ring[index][:, 336:672])
