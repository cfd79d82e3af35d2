"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the machin learner hot path.

A NumPy / CPU-torch restatement of the reference algorithms on the path named by
BASELINE.json:north_star, each function citing the reference file:line it follows.
Pinned against (a) the reference's own known-answer tests (tests/golden/kat_*.json,
copied *vectors* from test/frame/buffers/test_prioritized_buffer.py) and (b) outputs
of the unmodified reference executed in the build container
(tests/golden/ref_*.npz, made by oracle/make_golden.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package, and only as the checker / the
timed CPU arm.  Nothing under ``machin_b200/`` imports it.
"""
