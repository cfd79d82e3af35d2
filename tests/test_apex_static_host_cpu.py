"""CPU: the host part of DQNApex's fixed-shape update (frame/algorithms/apex.py: _static_prepare) with stand-ins for
the group and the buffer: apportioning and per-draw beta exactly like the dynamic exchange / the reference
(prioritized_buffer_d.py:229-234, prioritized_buffer.py:425-427), np.random consumed in the reference's sizes in
parity mode and in fixed blocks in device-apportion mode."""
import struct
import types

import numpy as np
import torch as t

from oracle.per import shard_sample_sizes


class _Event:
    def synchronize(self):
        pass

    def record(self):
        pass


def _learner(P, me, B, sums, device_apportion):
    from machin_b200.frame.algorithms import DQNApex
    learner = DQNApex.__new__(DQNApex)                    # no device: only the host arithmetic is under test
    learner.batch_size = B
    learner.apex_group = types.SimpleNamespace(index=me, all_gather_f64=lambda x: t.tensor(sums, dtype=t.float64).view(-1, 1))
    learner.replay_buffer = types.SimpleNamespace(curr_beta=0.4, beta_increment_per_sampling=1e-3, storage=[0] * 777,
                                                  weight_sum_tensor=lambda: None)
    cap = B
    z = lambda shape, dt: t.zeros(shape, dtype=dt)
    learner._static = {
        "P": P, "cap": cap, "capB": B + P, "device_apportion": device_apportion,
        "counts_pin": z(P, t.int64), "params_pin": z((P, 40), t.uint8), "uni_pin": z((P, cap), t.float64),
        "scale_pin": z(1, t.float64), "uni": z((P, cap), t.float64), "nstored": z(1, t.int64),
        "ring_uni": [z((P, cap), t.float64) for _ in range(4)], "ring_ns": [z(1, t.int64) for _ in range(4)],
        "ring_ev": [_Event() for _ in range(4)], "ring_i": 0}
    return learner


def test_parity_mode_prepares_what_the_reference_would_draw():
    P, me, B = 4, 2, 24
    sums = [3.0, 0.0, 11.5, 0.25]
    learner = _learner(P, me, B, sums, False)
    st, buf = learner._static, learner.replay_buffer
    np.random.seed(5)
    n = learner._static_prepare()
    ssize, total = shard_sample_sizes(sums, B)
    assert n == sum(ssize) and st["counts_pin"].tolist() == ssize
    assert st["scale_pin"].item() == (B + P) / n
    np.random.seed(5)
    beta = 0.4
    for r in range(P):
        b, tot, n_stored, step, n_draw = struct.unpack("<ddqqq", bytes(st["params_pin"][r].numpy()))
        assert (b, tot, n_stored, step, n_draw) == (beta, total, 777, 0, ssize[me])
        want = np.random.uniform(size=ssize[me])           # one draw of n_mine uniforms per requester, in member order
        assert np.array_equal(st["uni_pin"][r, :ssize[me]].numpy(), want)
        beta = float(np.min([1.0, beta + 1e-3]))
    assert buf.curr_beta == beta
    # a shard that is apportioned nothing draws nothing and keeps its beta
    learner = _learner(P, 1, B, sums, False)
    state = (np.random.seed(9), np.random.get_state())[1]
    learner._static_prepare()
    assert learner.replay_buffer.curr_beta == 0.4
    assert all(np.array_equal(a, b) if isinstance(a, np.ndarray) else a == b for a, b in zip(state, np.random.get_state()))


def test_device_apportion_mode_never_needs_the_sums():
    P, me, B = 4, 2, 24
    learner = _learner(P, me, B, None, True)               # all_gather_f64 would fail on None: it must not be called
    st, buf = learner._static, learner.replay_buffer
    for step in range(6):                                  # more steps than ring slots
        np.random.seed(step)
        assert learner._static_prepare() == B
        np.random.seed(step)
        assert np.array_equal(st["uni"].numpy(), np.random.uniform(size=(P, B)))
        assert st["nstored"].item() == 777
    assert abs(buf.curr_beta - (0.4 + 6 * P * 1e-3)) < 1e-12


def test_peer_timeout_flag_is_read_every_so_often():
    """A peer barrier that gave up leaves a flag on the device; the host-side update reads it once per PEER_CHECK_EVERY calls."""
    from machin_b200.frame.algorithms.apex import _FixedShapeExchange

    class Peer:
        checks = 0

        def check(self):
            Peer.checks += 1
            if Peer.checks == 2:
                raise RuntimeError("peer exchange: a member did not reach a barrier")

    learner = _FixedShapeExchange.__new__(_FixedShapeExchange)
    learner._check_peers_now_and_then()                       # no fixed-shape state: nothing to read
    learner._static = {"peer": None}
    learner._check_peers_now_and_then()                       # NCCL transport: nothing to read
    learner._static = {"peer": Peer()}
    learner.PEER_CHECK_EVERY = 4
    for _ in range(7):
        learner._check_peers_now_and_then()
    assert Peer.checks == 1
    import pytest
    with pytest.raises(RuntimeError, match="did not reach a barrier"):
        learner._check_peers_now_and_then()


def test_peer_timeout_flag_is_read_when_updates_turn_slow(monkeypatch):
    import time
    from machin_b200.frame.algorithms.apex import _FixedShapeExchange

    class Peer:
        checks = 0

        def check(self):
            Peer.checks += 1

    clock = [100.0]
    monkeypatch.setattr(time, "monotonic", lambda: clock[0])
    learner = _FixedShapeExchange.__new__(_FixedShapeExchange)
    learner._static = {"peer": Peer()}
    for _ in range(5):                                        # fast updates: no read
        clock[0] += 0.001
        learner._check_peers_now_and_then()
    assert Peer.checks == 0
    clock[0] += 80.0                                          # an update that sat out four 20 s barrier timeouts
    learner._check_peers_now_and_then()
    assert Peer.checks == 1
    learner._check_peers_now_and_then()
    assert Peer.checks == 1
