"""GPU: the scan kernels with the longest episode known to the launch -- fewer episodes per CTA keep the staged path
for long episodes, and a single episode longer than the staging tile is walked tile by tile -- against the oracle."""
import numpy as np
import pytest
import torch as t


from oracle import gae as ogae
from oracle import vtrace as ovt
from test_gpu_rows_td_scan import _vtrace_case

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("lengths,use_offsets", [([100] * 300, True), ([100] * 300, False), ([7000], True), ([7000], False),
                                                  ([2977, 3, 9000, 1], True), ([2976, 1, 2976, 50], True)])
def test_vtrace_long_episodes(lengths, use_offsets):
    from machin_b200 import ops
    case = _vtrace_case(lengths, 3, clip_c=0.9, clip_rho=1.2)
    vs, vss = ovt.vtrace_batched(case["value"], case["delta"], case["c"], case["lengths"], case["discount"])
    off = t.tensor(np.concatenate(([0], np.cumsum(lengths))), dtype=t.int64).cuda() if use_offsets else None
    out = ops.vtrace(case["value"].cuda(), case["next_value"].cuda(), case["reward"].cuda(), case["term"].cuda(),
                     case["logp"].cuda(), case["logmu"].cuda(), case["discount"], 0.9, 1.2, offsets=off,
                     n_episodes=len(lengths), max_len=max(lengths))
    np.testing.assert_allclose(out["rho"].cpu().numpy(), case["rho"].numpy(), rtol=1e-6)
    np.testing.assert_allclose(out["vs"].cpu().numpy(), vs.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["vss"].cpu().numpy(), vss.numpy(), rtol=1e-5, atol=1e-6)
    ends = np.cumsum(lengths) - 1
    assert np.all(out["vs"].cpu().numpy()[ends] == 0)
    # the launch without the length takes the old branches: same numbers bit for bit
    ref = ops.vtrace(case["value"].cuda(), case["next_value"].cuda(), case["reward"].cuda(), case["term"].cuda(),
                     case["logp"].cuda(), case["logmu"].cuda(), case["discount"], 0.9, 1.2,
                     offsets=t.tensor(np.concatenate(([0], np.cumsum(lengths))), dtype=t.int64).cuda(), n_episodes=len(lengths))
    for k in ("vs", "vss", "rho", "adv"):
        assert t.equal(out[k], ref[k]), k


@pytest.mark.parametrize("lam", [1.0, 0.0, 0.95])
@pytest.mark.parametrize("lengths", [[90] * 150, [5000], [1665, 3, 4000, 1]])
def test_gae_long_episodes(lam, lengths):
    from machin_b200 import ops
    rng = np.random.RandomState(int(lam * 100) + len(lengths))
    n = int(sum(lengths))
    reward = rng.randn(n)
    term = np.zeros(n, dtype=bool)
    term[(np.cumsum(lengths) - 1)[::2]] = True
    cv, ncv = rng.randn(n).astype(np.float32), rng.randn(n).astype(np.float32)
    off = t.tensor(np.concatenate(([0], np.cumsum(lengths))), dtype=t.int64).cuda()
    out = ops.gae(t.from_numpy(reward).cuda(), t.from_numpy(term).cuda(), t.from_numpy(cv).cuda(), t.from_numpy(ncv).cuda(),
                  0.97, lam, offsets=off, n_episodes=len(lengths), max_len=max(lengths))
    ret, adv = out["returns"].cpu().numpy(), out["gae"].cpu().numpy()
    pos = 0
    for L in lengths:
        sl = slice(pos, pos + L)
        want_ret = ogae.discounted_returns([float(x) for x in reward[sl]], 0.97)
        assert np.array_equal(ret[sl], np.array(want_ret))
        want = ogae.gae([float(x) for x in reward[sl]], [bool(x) for x in term[sl]], [float(x) for x in cv[sl]],
                        [float(x) for x in ncv[sl]], want_ret, 0.97, lam)
        assert np.array_equal(adv[sl], np.array(want))
        pos += L
