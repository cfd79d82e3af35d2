"""TEST INFRASTRUCTURE ONLY -- executable model of the small sorted tree update kernels' ALGORITHM
(machin_b200/csrc/sumtree.cu: tree_update_sorted and tree_update_sorted_v2), stepped on the CPU.

It is not the reference's algorithm (that is oracle/sumtree.py: OracleTree.update_leaf_batch, following
prioritized_buffer.py:152-186) but a thread-by-thread restatement of what the CUDA kernels do -- compaction of
equal keys, hb[] from neighbouring keys, prefetched stored siblings, the barrier walk or the polling walk with
point-to-point hand-overs, the split into 2^CB subtrees with the top combined by a leader -- so that the
merge logic, its freedom from deadlock under ANY lane scheduling and its bit-identity with the reference can be
checked where no GPU exists (tests/test_sorted_walk_model_cpu.py).  Nothing under machin_b200/ imports it.
"""
import numpy as np


def level_off(total, l):
    return total - (total >> l)


def update_sorted(tree, idx, w, cluster_bits=0, walk="barrier", rng=None, warp=32, max_rounds=10_000):
    """Apply update_leaf_batch(w, idx) for ASCENDING idx to ``tree`` (an OracleTree) the way the kernels do.
    walk: "barrier" (one CTA barrier per level) or "poll" (one non-blocking step per lane per round; lanes and
    warps are stepped in a random order drawn from ``rng`` to model arbitrary scheduling).
    Returns the number of polling rounds of the slowest warp (0 for the barrier walk)."""
    W, depth = tree.weights, tree.depth
    total, levels = 1 << depth, depth - 1
    ls = levels - cluster_bits
    idx = np.asarray(idx, dtype=np.int64)
    w = np.asarray(w)
    n = len(idx)
    assert np.all(idx[1:] >= idx[:-1]) and idx.min() >= 0 and idx.max() < tree.size
    tree.max_leaf = max(np.max(w), tree.max_leaf)
    rng = rng or np.random.RandomState(0)
    tops, worst_rounds = [], 0
    stores = []                                             # (address, value): written after every CTA's walk
    for rank in range(1 << cluster_bits):
        mine = [i for i in range(n) if (int(idx[i]) >> ls) == rank]
        winners = [i for i in mine if i == n - 1 or idx[i + 1] != idx[i]]       # last of a run of equal keys
        m = len(winners)
        key = [int(idx[i]) for i in winners]
        val = [np.float64(w[i]) for i in winners]
        leaf = list(val)
        hb = [ls if e == 0 else (key[e - 1] ^ key[e]).bit_length() - 1 for e in range(m)]
        assert all(h < ls for h in hb[1:])
        park = {}
        for e in range(m):
            lim = hb[e] if hb[e] < ls else ls - 1
            for l in range(lim + 1):
                park[(l, e)] = W[level_off(total, l) + ((key[e] >> l) ^ 1)]
        wrote = [dict() for _ in range(m)]
        if walk == "barrier":
            alive = [True] * m
            nxt = [e + 1 for e in range(m)]
            handed = {}
            for l in range(ls):
                for e in range(m):
                    if alive[e] and hb[e] == l:
                        handed[e] = val[e]
                        alive[e] = False
                new_next = list(nxt)
                for e in range(m):                           # all lanes read the state published before the barrier
                    if not alive[e]:
                        continue
                    j = nxt[e]
                    if j < m and hb[j] == l:
                        other = handed[j]
                        new_next[e] = nxt[j]
                    else:
                        other = park[(l, e)]
                    val[e] = val[e] + other
                    wrote[e][l] = val[e]
                nxt = new_next
        else:
            nx = [e + 1 for e in range(m)]
            lv = [0] * m
            lim = [(ls if e == 0 else hb[e]) for e in range(m)]
            done = [False] * m
            flag, pub_val, pub_next = [0] * m, [None] * m, [None] * m
            n_warps = (m + warp - 1) // warp
            warp_rounds = [0] * n_warps
            active = set(range(n_warps))
            guard = 0
            while active:
                guard += 1
                assert guard < max_rounds * max(1, n_warps), "polling walk does not terminate"
                wp = list(active)[int(rng.randint(len(active)))]           # any warp may run next
                lanes = [e for e in range(wp * warp, min(m, (wp + 1) * warp))]
                if all(done[e] for e in lanes):
                    active.discard(wp)
                    continue
                warp_rounds[wp] += 1
                for e in rng.permutation(lanes):                           # any lane order inside a round
                    e = int(e)
                    if done[e]:
                        continue
                    if lv[e] == lim[e]:
                        if e != 0:
                            pub_val[e], pub_next[e], flag[e] = val[e], nx[e], 1
                        done[e] = True
                    elif nx[e] < m and hb[nx[e]] == lv[e]:
                        j = nx[e]
                        if flag[j]:
                            val[e] = val[e] + pub_val[j]
                            nx[e] = pub_next[j]
                            wrote[e][lv[e]] = val[e]
                            lv[e] += 1
                    else:
                        val[e] = val[e] + park[(lv[e], e)]
                        wrote[e][lv[e]] = val[e]
                        lv[e] += 1
            worst_rounds = max([worst_rounds] + warp_rounds)
        tops.append(val[0] if m > 0 else W[level_off(total, ls) + rank])
        for e in range(m):
            stores.append((key[e], leaf[e]))
            for l, v in wrote[e].items():
                stores.append((level_off(total, l + 1) + (key[e] >> (l + 1)), v))
    for addr, v in stores:
        W[addr] = v
    v = list(tops)
    cnt = 1 << cluster_bits
    for b in range(1, cluster_bits + 1):
        cnt >>= 1
        for j in range(cnt):
            v[j] = v[2 * j] + v[2 * j + 1]
            W[level_off(total, ls + b) + j] = v[j]
    return worst_rounds
