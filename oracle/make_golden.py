"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED
reference (iffiX/machin mounted at /root/reference) in the build container.

    python -m oracle.make_golden            # from the repo root

The fixtures travel to the GPU box; the reference does not.  Every array saved here is
an input or an output of reference code -- nothing is produced by the oracle restatement
or by the CUDA path.  Seeds follow the reference test-suite convention
(test/util_fixtures.py:50-55).
"""
import os
import random
import sys

import numpy as np
import torch as t

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def seed_all(s=0):
    t.manual_seed(s)
    np.random.seed(s)
    random.seed(s)


# --------------------------------------------------------------------------- tree
def gen_tree():
    from machin.frame.buffers import WeightTree

    out = {}
    rng = np.random.RandomState(1234)
    for size in (1, 2, 3, 5, 8, 1000, 1025, 4096):
        tree = WeightTree(size)
        leaves = rng.exponential(size=size)
        leaves[rng.rand(size) < 0.1] = 0.0
        tree.update_all_leaves(leaves)
        out[f"n{size}_leaves"] = leaves
        out[f"n{size}_built"] = tree.weights.copy()
        out[f"n{size}_built_max"] = np.float64(tree.max_leaf)
        total = tree.get_weight_sum()
        q = np.concatenate((rng.uniform(-0.5, total * 1.1 + 0.5, size=257),
                            [0.0, total, total - 1e-6, -1.0, total * 2 + 1],
                            np.cumsum(leaves)[: min(size, 64)]))
        out[f"n{size}_query"] = q
        out[f"n{size}_found"] = np.asarray(tree.find_leaf_index(q)).reshape(-1)
        # batched update with duplicates (last write wins) then a second, float32 one
        nb = max(1, min(size * 2, 300))
        idx = rng.randint(0, size, size=nb).astype(np.int64)
        w = rng.exponential(size=nb)
        tree.update_leaf_batch(w, idx)
        out[f"n{size}_upd_idx"] = idx
        out[f"n{size}_upd_w"] = w
        out[f"n{size}_after_upd"] = tree.weights.copy()
        out[f"n{size}_after_upd_max"] = np.float64(tree.max_leaf)
        idx2 = rng.randint(0, size, size=nb).astype(np.int64)
        w2 = rng.exponential(size=nb).astype(np.float32)
        tree.update_leaf_batch(w2, idx2)
        out[f"n{size}_upd2_idx"] = idx2
        out[f"n{size}_upd2_w"] = w2
        out[f"n{size}_after_upd2"] = tree.weights.copy()
        # single-leaf updates
        # (size 1: the reference's update_leaf raises IndexError at :140 -- quirk, not pinned)
        for k in range(3 if size > 1 else 0):
            i = int(rng.randint(0, size))
            v = float(rng.exponential())
            tree.update_leaf(v, i)
            out[f"n{size}_single{k}"] = np.array([i, v])
        out[f"n{size}_after_single"] = tree.weights.copy()
        out[f"n{size}_found2"] = np.asarray(tree.find_leaf_index(q)).reshape(-1)
    np.savez_compressed(os.path.join(OUT, "ref_tree.npz"), **out)


# --------------------------------------------------------------------------- PER
def _episode(n, state_dim, rng, reward_int=False):
    ep = []
    for _ in range(n):
        ep.append(dict(
            state={"state": t.from_numpy(rng.randn(1, state_dim).astype(np.float32))},
            action={"action": t.tensor([[int(rng.randint(0, 6))]], dtype=t.long)},
            next_state={"state": t.from_numpy(rng.randn(1, state_dim).astype(np.float32))},
            reward=float(np.float32(rng.randn())),
            terminal=bool(rng.rand() < 0.1),
        ))
    return ep


def gen_per():
    from machin.frame.buffers import PrioritizedBuffer

    out = {}
    rng = np.random.RandomState(99)
    for tag, cap, n_store in (("small", 50, 37), ("wrap", 64, 150), ("big", 5000, 5000)):
        buf = PrioritizedBuffer(buffer_size=cap, buffer_device="cpu")
        stored = 0
        pri_all = []
        len_all = []
        while stored < n_store:
            n = int(rng.randint(1, 12))
            n = min(n, n_store - stored)
            pri = rng.exponential(size=n)
            buf.store_episode(_episode(n, 4, rng), priorities=list(pri))
            pri_all.append(pri)
            len_all.append(n)
            stored += n
        out[f"{tag}_cap"] = np.int64(cap)
        out[f"{tag}_ep_len"] = np.array(len_all, dtype=np.int64)
        out[f"{tag}_ep_pri"] = np.concatenate(pri_all)
        out[f"{tag}_tree"] = buf.wt_tree.weights.copy()
        out[f"{tag}_max_leaf"] = np.float64(buf.wt_tree.max_leaf)
        seed_all(7)
        for step in range(4):
            b = (16, 32, 5, 100)[step]
            state = np.random.get_state()
            uni = np.random.uniform(size=b)
            np.random.set_state(state)
            beta_before = float(buf.curr_beta)
            index, isw = buf.sample_index_and_weight(b)
            out[f"{tag}_s{step}_uniform"] = uni
            out[f"{tag}_s{step}_beta"] = np.float64(beta_before)
            out[f"{tag}_s{step}_nstored"] = np.int64(len(buf.storage))
            out[f"{tag}_s{step}_index"] = np.asarray(index).reshape(-1)
            out[f"{tag}_s{step}_isw"] = np.asarray(isw).reshape(-1)
            # priority write-back, float32 like abs_error from torch (dqn_per.py:132-139)
            new_p = rng.exponential(size=b).astype(np.float32)
            buf.update_priority(new_p, np.asarray(index).reshape(-1))
            out[f"{tag}_s{step}_newp"] = new_p
            out[f"{tag}_s{step}_tree_after"] = buf.wt_tree.weights.copy()
        # store without priorities -> normalize(max_leaf)   (prioritized_buffer.py:302-305)
        buf.store_episode(_episode(3, 4, rng))
        out[f"{tag}_tree_after_default_store"] = buf.wt_tree.weights.copy()
    np.savez_compressed(os.path.join(OUT, "ref_per.npz"), **out)


# --------------------------------------------------------------------------- TD (DQNPer / DDPGPer)
class _QNet(t.nn.Module):
    def __init__(self, s, a, h=16):
        super().__init__()
        self.fc1 = t.nn.Linear(s, h)
        self.fc2 = t.nn.Linear(h, h)
        self.fc3 = t.nn.Linear(h, a)

    def forward(self, state):
        x = t.relu(self.fc1(state))
        return self.fc3(t.relu(self.fc2(x)))


class _Actor(t.nn.Module):
    def __init__(self, s, a, h=16):
        super().__init__()
        self.fc1 = t.nn.Linear(s, h)
        self.fc2 = t.nn.Linear(h, a)

    def forward(self, state):
        return t.tanh(self.fc2(t.relu(self.fc1(state))))


class _Critic(t.nn.Module):
    def __init__(self, s, a, h=16):
        super().__init__()
        self.fc1 = t.nn.Linear(s + a, h)
        self.fc2 = t.nn.Linear(h, 1)

    def forward(self, state, action):
        return self.fc2(t.relu(self.fc1(t.cat([state, action], dim=1))))


def _flat_params(m):
    return np.concatenate([p.detach().numpy().reshape(-1) for p in m.parameters()])


def gen_td():
    from machin.frame.algorithms import DQNPer, DDPGPer
    from machin.model.nets.base import static_module_wrapper as smw

    out = {}
    # ---- double-DQN with PER; SGD so that the whole step is reproducible elsewhere
    seed_all(0)
    S, A, N, B = 6, 4, 200, 32
    qnet, qtgt = _QNet(S, A), _QNet(S, A)
    dqn = DQNPer(smw(qnet, "cpu", "cpu"), smw(qtgt, "cpu", "cpu"), t.optim.SGD,
                 t.nn.MSELoss(reduction="sum"), batch_size=B, learning_rate=0.01,
                 discount=0.97, update_rate=0.05, replay_size=N, mode="double")
    # the ctor copies the TARGET net into the online net (dqn.py:198-200), so save after it
    out["dqn_init_q"] = _flat_params(dqn.qnet)
    out["dqn_init_qt"] = _flat_params(dqn.qnet_target)
    rng = np.random.RandomState(5)
    st = rng.randn(N, S).astype(np.float32)
    nst = rng.randn(N, S).astype(np.float32)
    act = rng.randint(0, A, size=N).astype(np.int64)
    rew = rng.randn(N).astype(np.float32)
    term = rng.rand(N) < 0.2
    pri = rng.exponential(size=N)
    ep = [dict(state={"state": t.from_numpy(st[i:i + 1])}, action={"action": t.tensor([[act[i]]])},
               next_state={"state": t.from_numpy(nst[i:i + 1])}, reward=float(rew[i]),
               terminal=bool(term[i])) for i in range(N)]
    dqn.replay_buffer.store_episode(ep, priorities=list(pri))
    out.update(dqn_state=st, dqn_next_state=nst, dqn_action=act, dqn_reward=rew,
               dqn_terminal=term, dqn_priority=pri, dqn_tree0=dqn.replay_buffer.wt_tree.weights.copy())
    seed_all(11)
    for step in range(3):
        s0 = np.random.get_state()
        out[f"dqn_u{step}"] = np.random.uniform(size=B)
        np.random.set_state(s0)
        loss = dqn.update()
        out[f"dqn_loss{step}"] = np.float64(loss)
        out[f"dqn_tree{step + 1}"] = dqn.replay_buffer.wt_tree.weights.copy()
        out[f"dqn_q{step + 1}"] = _flat_params(dqn.qnet)
        out[f"dqn_qt{step + 1}"] = _flat_params(dqn.qnet_target)

    # ---- DDPG with PER
    seed_all(0)
    S, A, N, B = 5, 2, 150, 16
    actor, actor_t, critic, critic_t = _Actor(S, A), _Actor(S, A), _Critic(S, A), _Critic(S, A)
    ddpg = DDPGPer(smw(actor, "cpu", "cpu"), smw(actor_t, "cpu", "cpu"),
                   smw(critic, "cpu", "cpu"), smw(critic_t, "cpu", "cpu"), t.optim.SGD,
                   t.nn.MSELoss(reduction="sum"), batch_size=B, actor_learning_rate=0.005,
                   critic_learning_rate=0.01, discount=0.95, update_rate=0.1, replay_size=N)
    # ctor copies target -> online (ddpg.py:160-162)
    out["ddpg_init_actor"] = _flat_params(ddpg.actor)
    out["ddpg_init_critic"] = _flat_params(ddpg.critic)
    out["ddpg_init_actor_t"] = _flat_params(ddpg.actor_target)
    out["ddpg_init_critic_t"] = _flat_params(ddpg.critic_target)
    rng = np.random.RandomState(6)
    st = rng.randn(N, S).astype(np.float32)
    nst = rng.randn(N, S).astype(np.float32)
    act = rng.uniform(-1, 1, size=(N, A)).astype(np.float32)
    rew = rng.randn(N).astype(np.float32)
    term = rng.rand(N) < 0.2
    pri = rng.exponential(size=N)
    ep = [dict(state={"state": t.from_numpy(st[i:i + 1])}, action={"action": t.from_numpy(act[i:i + 1])},
               next_state={"state": t.from_numpy(nst[i:i + 1])}, reward=float(rew[i]),
               terminal=bool(term[i])) for i in range(N)]
    ddpg.replay_buffer.store_episode(ep, priorities=list(pri))
    out.update(ddpg_state=st, ddpg_next_state=nst, ddpg_action=act, ddpg_reward=rew,
               ddpg_terminal=term, ddpg_priority=pri)
    seed_all(12)
    for step in range(3):
        s0 = np.random.get_state()
        out[f"ddpg_u{step}"] = np.random.uniform(size=B)
        np.random.set_state(s0)
        pol, val = ddpg.update()
        out[f"ddpg_loss{step}"] = np.array([pol, val], dtype=np.float64)
        out[f"ddpg_tree{step + 1}"] = ddpg.replay_buffer.wt_tree.weights.copy()
        out[f"ddpg_actor{step + 1}"] = _flat_params(ddpg.actor)
        out[f"ddpg_critic{step + 1}"] = _flat_params(ddpg.critic)
        out[f"ddpg_actor_t{step + 1}"] = _flat_params(ddpg.actor_target)
        out[f"ddpg_critic_t{step + 1}"] = _flat_params(ddpg.critic_target)
    np.savez_compressed(os.path.join(OUT, "ref_td.npz"), **out)


# --------------------------------------------------------------------------- V-trace (IMPALA)
class _FakeFuture:
    def __init__(self, v):
        self.v = v

    def wait(self):
        return self.v


class FakeGroup:
    """In-process stand-in for RpcGroup: numeric comparison only."""

    def __init__(self):
        self.services = {}

    def is_member(self, target=None):
        return True

    def get_cur_name(self):
        return "0"

    def get_group_members(self):
        return ["0"]

    def size(self):
        return 1

    def register(self, key, fn):
        self.services[key] = fn

    def registered_async(self, key, args=(), kwargs=None):
        return _FakeFuture(self.services[key](*args, **(kwargs or {})))

    def registered_sync(self, key, args=(), kwargs=None):
        return self.services[key](*args, **(kwargs or {}))


class _NoServer:
    def push(self, model, pull_on_fail=True):
        return True

    def pull(self, model):
        return True


class _PolicyNet(t.nn.Module):
    def __init__(self, s, a, h=16):
        super().__init__()
        self.fc1 = t.nn.Linear(s, h)
        self.fc2 = t.nn.Linear(h, a)

    def forward(self, state, action=None):
        from torch.distributions import Categorical
        probs = t.softmax(self.fc2(t.relu(self.fc1(state))), dim=1)
        dist = Categorical(probs=probs)
        act = action if action is not None else dist.sample()
        ent = dist.entropy()
        logp = dist.log_prob(act.flatten())
        return act.view(-1, 1), logp.view(-1, 1), ent.view(-1, 1)


class _ValueNet(t.nn.Module):
    def __init__(self, s, h=16):
        super().__init__()
        self.fc1 = t.nn.Linear(s, h)
        self.fc2 = t.nn.Linear(h, 1)

    def forward(self, state):
        return self.fc2(t.relu(self.fc1(state)))


def gen_vtrace():
    from machin.frame.algorithms import IMPALA
    from machin.model.nets.base import static_module_wrapper as smw

    out = {}
    for tag, lengths, clip_c, clip_rho in (("fixed", [20] * 12, 1.0, 1.0),
                                           ("ragged", [1, 7, 2, 33, 5, 1, 12, 40, 3], 0.9, 1.3)):
        seed_all(0)
        S, A = 8, 3
        actor, critic = _PolicyNet(S, A), _ValueNet(S)
        out[f"{tag}_actor0"] = _flat_params(actor)
        out[f"{tag}_critic0"] = _flat_params(critic)
        captured = {}

        def criterion(value, vs, _c=captured):
            _c["value"] = value.detach().clone()
            _c["vs"] = vs.detach().clone()
            return ((value - vs) ** 2).sum()

        imp = IMPALA(smw(actor, "cpu", "cpu"), smw(critic, "cpu", "cpu"), t.optim.SGD, criterion,
                     FakeGroup(), (_NoServer(),), batch_size=len(lengths), learning_rate=0.01,
                     isw_clip_c=clip_c, isw_clip_rho=clip_rho, discount=0.98,
                     replay_size=4 * sum(lengths))
        rng = np.random.RandomState(21)
        total = sum(lengths)
        st = rng.randn(total, S).astype(np.float32)
        nst = rng.randn(total, S).astype(np.float32)
        act = rng.randint(0, A, size=total).astype(np.int64)
        rew = rng.randn(total).astype(np.float32)
        logmu = (-rng.exponential(size=total)).astype(np.float32)
        term = np.zeros(total, dtype=bool)
        pos = 0
        episodes = []
        for n in lengths:
            term[pos + n - 1] = True
            episodes.append([dict(state={"state": t.from_numpy(st[i:i + 1])},
                                  action={"action": t.tensor([[act[i]]])},
                                  next_state={"state": t.from_numpy(nst[i:i + 1])},
                                  reward=float(rew[i]), action_log_prob=float(logmu[i]),
                                  terminal=bool(term[i])) for i in range(pos, pos + n)])
            pos += n
        for ep in episodes:
            imp.store_episode(ep)
        # IMPALABuffer samples episodes with random.choices (impala.py:58): record the picks
        seed_all(3)
        s0 = random.getstate()
        picks = random.choices(range(len(lengths)), k=len(lengths))
        random.setstate(s0)
        pol, val = imp.update()
        out.update({f"{tag}_lengths": np.array(lengths), f"{tag}_state": st, f"{tag}_next_state": nst,
                    f"{tag}_action": act, f"{tag}_reward": rew, f"{tag}_logmu": logmu,
                    f"{tag}_terminal": term, f"{tag}_picks": np.array(picks),
                    f"{tag}_clip": np.array([clip_c, clip_rho, 0.98]),
                    f"{tag}_value": captured["value"].numpy().reshape(-1),
                    f"{tag}_vs": captured["vs"].numpy().reshape(-1),
                    f"{tag}_losses": np.array([pol, val], dtype=np.float64),
                    f"{tag}_actor1": _flat_params(imp.actor), f"{tag}_critic1": _flat_params(imp.critic)})
    np.savez_compressed(os.path.join(OUT, "ref_vtrace.npz"), **out)


# --------------------------------------------------------------------------- GAE (A2C / A3C)
def gen_gae():
    from machin.frame.algorithms import A2C
    from machin.model.nets.base import static_module_wrapper as smw

    out = {}
    for lam in (1.0, 0.0, 0.95):
        seed_all(0)
        S, A = 8, 3
        actor, critic = _PolicyNet(S, A), _ValueNet(S)
        out[f"critic0"] = _flat_params(critic)
        a2c = A2C(smw(actor, "cpu", "cpu"), smw(critic, "cpu", "cpu"), t.optim.SGD,
                  t.nn.MSELoss(reduction="sum"), batch_size=8, discount=0.97, gae_lambda=lam)
        rng = np.random.RandomState(31)
        for e, n in enumerate((1, 2, 9, 30)):
            st = rng.randn(n, S).astype(np.float32)
            nst = rng.randn(n, S).astype(np.float32)
            rew = rng.randn(n).astype(np.float32)
            term = np.zeros(n, dtype=bool)
            term[-1] = e % 2 == 0
            ep = [dict(state={"state": t.from_numpy(st[i:i + 1])}, action={"action": t.tensor([[0]])},
                       next_state={"state": t.from_numpy(nst[i:i + 1])}, reward=float(rew[i]),
                       terminal=bool(term[i])) for i in range(n)]
            a2c.store_episode(ep)
            key = f"lam{lam}_ep{e}"
            out[f"{key}_state"] = st
            out[f"{key}_next_state"] = nst
            out[f"{key}_reward"] = rew
            out[f"{key}_terminal"] = term
            out[f"{key}_value"] = np.array([tr["value"] for tr in ep], dtype=np.float64)
            out[f"{key}_gae"] = np.array([tr["gae"] for tr in ep], dtype=np.float64)
            with t.no_grad():
                out[f"{key}_critic_v"] = critic(t.from_numpy(st)).numpy().reshape(-1)
                out[f"{key}_critic_nv"] = critic(t.from_numpy(nst)).numpy().reshape(-1)
    np.savez_compressed(os.path.join(OUT, "ref_gae.npz"), **out)


# --------------------------------------------------------------------------- uniform Buffer bookkeeping
def gen_buffer():
    """Slot assignment, whole-episode eviction and the order in which the reference's Buffer offers
    slots to random.sample / random.choices (buffer.py:46-109, :224-252; storage.py:85-114)."""
    from machin.frame.buffers import Buffer

    out = {}
    for tag, cap, lens in (("a", 10, [3, 4, 2, 5, 1, 6, 2, 2, 7, 1]), ("b", 64, [9, 30, 17, 25, 3, 11, 40, 8, 5, 22, 1, 13])):
        buf = Buffer(buffer_size=cap)
        rng = np.random.RandomState(4)
        keys, picks_u, picks_r, sizes = [], [], [], []
        for n in lens:
            buf.store_episode(_episode(n, 2, rng))
            ks = list(buf.transition_episode_number.keys())
            keys.append(np.array(ks + [-1] * (cap - len(ks)), dtype=np.int64))
            sizes.append(buf.size())
            random.seed(5)
            bs, batch = buf.sample_method_random_unique(4)
            random.seed(5)
            picks_u.append(np.array(random.sample(list(buf.transition_episode_number.keys()), k=bs) + [-1] * (4 - bs)))
            random.seed(6)
            picks_r.append(np.array(random.choices(list(buf.transition_episode_number.keys()), k=min(len(ks), 4)) + [-1] * (4 - min(len(ks), 4))))
        out[f"{tag}_cap"] = np.int64(cap)
        out[f"{tag}_lens"] = np.array(lens, dtype=np.int64)
        out[f"{tag}_keys"] = np.stack(keys)
        out[f"{tag}_size"] = np.array(sizes, dtype=np.int64)
        out[f"{tag}_pick_unique"] = np.stack(picks_u)
        out[f"{tag}_pick_random"] = np.stack(picks_r)
    np.savez_compressed(os.path.join(OUT, "ref_buffer.npz"), **out)


# --------------------------------------------------------------------------- sharded PER: default priorities
def gen_per_d():
    """DistributedPrioritizedBuffer.store_episode without priorities (prioritized_buffer_d.py:123-131): one
    update_leaf per transition with normalize(get_leaf_max()) re-read every time -- the maximum compounds
    inside an episode -- plus the version ("taint") counters (:131)."""
    from machin.frame.buffers import DistributedPrioritizedBuffer

    out = {}
    rng = np.random.RandomState(8)
    buf = DistributedPrioritizedBuffer("buffer", FakeGroup(), 40)
    trees, versions, lens = [], [], []
    for step, n in enumerate((4, 7, 3, 12, 9, 20)):
        buf.store_episode(_episode(n, 2, rng))
        lens.append(n)
        trees.append(buf.wt_tree.weights.copy())
        versions.append(buf.buffer_version_table.astype(np.int64).copy())
        if step == 1:   # a float32 write-back in between changes the dtype of max_leaf (np.float32)
            idx = np.array([0, 3, 5], dtype=np.int64)
            pri = np.array([2.5, 0.1, 7.25], dtype=np.float32)
            buf._update_priority_service(pri, idx, buf.buffer_version_table[idx])
            out["wb_idx"], out["wb_pri"] = idx, pri
            out["tree_after_wb"] = buf.wt_tree.weights.copy()
    out["lens"] = np.array(lens)
    out["trees"] = np.stack(trees)
    out["versions"] = np.stack(versions)
    np.savez_compressed(os.path.join(OUT, "ref_per_d.npz"), **out)


# --------------------------------------------------------------------------- RNN (sequence) buffers
def _indexed_episode(n, state_dim, rng, counter):
    ep = _episode(n, state_dim, rng)
    for i, tr in enumerate(ep):
        tr["index"] = counter + i
    return ep


def gen_rnn():
    """RNNBuffer window selection (rnn_buffers.py:78-167) and RNNPrioritizedBuffer store / sample
    (:343-413), driven through the unmodified reference."""
    from machin.frame.buffers import PrioritizedBuffer
    from machin.frame.buffers.rnn_buffers import RNNBuffer, RNNPrioritizedBuffer

    out = {}
    # ---- uniform: which windows the three sample methods return, after every store
    for tag, cap, L, lens in (("a", 30, 3, [5, 2, 7, 3, 9, 4, 1, 6, 8, 3]), ("b", 200, 5, [12, 4, 30, 5, 17, 9, 41, 6, 25, 3, 50, 14])):
        buf = RNNBuffer(sample_length=L, buffer_size=cap)
        rng = np.random.RandomState(11)
        counter = 0
        uniq, rand, allw, nall = [], [], [], []
        for n in lens:
            buf.store_episode(_indexed_episode(n, 2, rng, counter))
            counter += n
            random.seed(5)
            bs, batch = buf.sample_method_random_unique(4)
            uniq.append(np.array([tr["index"] for tr in batch] + [-1] * (4 * L - bs * L), dtype=np.int64))
            random.seed(6)
            bs, batch = buf.sample_method_random(4)
            rand.append(np.array([tr["index"] for tr in batch] + [-1] * (4 * L - bs * L), dtype=np.int64))
            bs, batch = buf.sample_method_all(0)
            nall.append(bs)
            allw.append(np.array([tr["index"] for tr in batch], dtype=np.int64))
        out[f"u{tag}_cap"], out[f"u{tag}_L"] = np.int64(cap), np.int64(L)
        out[f"u{tag}_lens"] = np.array(lens, dtype=np.int64)
        out[f"u{tag}_unique"], out[f"u{tag}_random"] = np.stack(uniq), np.stack(rand)
        out[f"u{tag}_all_count"] = np.array(nall, dtype=np.int64)
        for k, a in enumerate(allw):
            out[f"u{tag}_all{k}"] = a
        # a concatenated batch: window axis after the batch axis (:169-182)
        random.seed(9)
        bs, (state, reward, idx) = buf.sample_batch(3, sample_method="random_unique", sample_attrs=["state", "reward", "index"],
                                                    additional_concat_custom_attrs=["index"])
        out[f"u{tag}_batch_state"] = state["state"].numpy()
        out[f"u{tag}_batch_reward"] = reward.numpy()
        out[f"u{tag}_batch_index"] = idx.numpy()
    # ---- prioritized: leaf weights after each store (given / default priorities), draws, write-back
    for tag, cap, L, lens in (("a", 64, 4, [10, 3, 12, 8, 2, 9, 20]), ("b", 40, 3, [10, 10, 12, 8, 9, 11, 5, 15])):
        # RNNPrioritizedBuffer(...) itself raises TypeError in the reference (RNNBuffer.__init__ forwards
        # storage=None to PrioritizedBuffer.__init__, which passes storage again, rnn_buffers.py:71 /
        # prioritized_buffer.py:269): initialise the instance through the base initialiser and set the two
        # attributes RNNBuffer.__init__ would have set; every method under test is the unmodified one
        buf = RNNPrioritizedBuffer.__new__(RNNPrioritizedBuffer)
        PrioritizedBuffer.__init__(buf, buffer_size=cap)
        buf.sample_length, buf.sample_dimension = L, 1
        rng = np.random.RandomState(12)
        counter = 0
        trees, pri_in, states = [], [], []
        for k, n in enumerate(lens):
            ep = _indexed_episode(n, 3, rng, counter)
            states.append(np.concatenate([tr["state"]["state"].numpy() for tr in ep]))
            counter += n
            if k % 3 == 2:
                buf.store_episode(ep)                       # default priority
                pri_in.append(np.full(n, -1.0))
            else:
                pri = rng.exponential(size=n)
                pri_in.append(pri.copy())
                buf.store_episode(ep, priorities=pri)       # ndarray: the reference slices / zeroes it
            trees.append(buf.wt_tree.weights.copy())
            if k == 3:                                      # float32 write-back: max_leaf becomes np.float32
                seed_all(21)
                n_s, _, index, isw = buf.sample_batch(6, sample_attrs=["index"], additional_concat_custom_attrs=["index"])
                new_p = rng.exponential(size=n_s).astype(np.float32) * 5
                buf.update_priority(new_p, index)
                out[f"p{tag}_mid_index"], out[f"p{tag}_mid_newp"] = np.asarray(index), new_p
                out[f"p{tag}_mid_tree"] = buf.wt_tree.weights.copy()
        out[f"p{tag}_cap"], out[f"p{tag}_L"] = np.int64(cap), np.int64(L)
        out[f"p{tag}_lens"] = np.array(lens, dtype=np.int64)
        out[f"p{tag}_pri"] = np.concatenate(pri_in)
        out[f"p{tag}_states"] = np.concatenate(states)
        out[f"p{tag}_trees"] = np.stack(trees)
        seed_all(7)
        for step, b in enumerate((8, 3, 16)):
            st = np.random.get_state()
            out[f"p{tag}_s{step}_uniform"] = np.random.uniform(size=b)
            np.random.set_state(st)
            out[f"p{tag}_s{step}_beta"] = np.float64(buf.curr_beta)
            n_s, (state, reward, idx), index, isw = buf.sample_batch(
                b, sample_attrs=["state", "reward", "index"], additional_concat_custom_attrs=["index"])
            out[f"p{tag}_s{step}_index"], out[f"p{tag}_s{step}_isw"] = np.asarray(index), np.asarray(isw)
            out[f"p{tag}_s{step}_state"] = state["state"].numpy()
            out[f"p{tag}_s{step}_reward"] = reward.numpy()
            out[f"p{tag}_s{step}_seq"] = idx.numpy()
            new_p = rng.exponential(size=n_s).astype(np.float32)
            buf.update_priority(new_p, index)
            out[f"p{tag}_s{step}_newp"] = new_p
            out[f"p{tag}_s{step}_tree_after"] = buf.wt_tree.weights.copy()
    np.savez_compressed(os.path.join(OUT, "ref_rnn.npz"), **out)


# --------------------------------------------------------------------------- RAINBOW (distributional target)
class _DistQNet(t.nn.Module):
    """The shape contract of rainbow.py:44-48: [batch, action_num, atom_num], last dimension soft-maxed."""

    def __init__(self, s, a, atoms, h=16):
        super().__init__()
        self.fc1 = t.nn.Linear(s, h)
        self.fc2 = t.nn.Linear(h, a * atoms)
        self.a, self.atoms = a, atoms

    def forward(self, state):
        return t.softmax(self.fc2(t.relu(self.fc1(state))).view(-1, self.a, self.atoms), dim=-1)


def gen_rainbow():
    from machin.frame.algorithms import RAINBOW
    from machin.model.nets.base import static_module_wrapper as smw

    out = {}
    seed_all(0)
    S, A, ATOMS, B, NSTEP = 5, 3, 11, 32, 3
    VMIN, VMAX, DISC = -2.0, 3.0, 0.9          # delta_z = 0.5: rewards on the grid land on atoms exactly
    qnet, qtgt = _DistQNet(S, A, ATOMS), _DistQNet(S, A, ATOMS)
    rb = RAINBOW(smw(qnet, "cpu", "cpu"), smw(qtgt, "cpu", "cpu"), t.optim.SGD, VMIN, VMAX, batch_size=B,
                 learning_rate=0.05, discount=DISC, update_rate=0.1, reward_future_steps=NSTEP, replay_size=400)
    out["init_q"], out["init_qt"] = _flat_params(rb.qnet), _flat_params(rb.qnet_target)
    out["config"] = np.array([S, A, ATOMS, B, NSTEP, VMIN, VMAX, DISC, 0.05, 0.1])
    rng = np.random.RandomState(21)
    # episodes of several lengths (shorter than, equal to, longer than the n-step horizon); rewards: half of them
    # on the 0.5 grid (so that terminal targets hit atoms exactly: both neighbour weights are 0 there,
    # rainbow.py:257-262), some far outside [v_min, v_max] (clamped)
    lens = [1, 2, 3, 4, 1, 2, 7, 1, 3, 12, 2, 1, 30, 2, 1, 3, 41, 1, 2, 2]
    ep_state, ep_next, ep_act, ep_rew, ep_term, ep_value = [], [], [], [], [], []
    for ei, n in enumerate(lens):
        st, nst = rng.randn(n, S).astype(np.float32), rng.randn(n, S).astype(np.float32)
        act = rng.randint(0, A, size=n).astype(np.int64)
        grid = rng.choice([-1.0, -0.5, 0.0, 0.5, 1.0, 6.0, -7.0], size=n)
        rew = np.where(rng.rand(n) < 0.5, grid, rng.randn(n)).astype(np.float64)
        if n <= 2:
            rew[-1] = (-1.0, -0.5, 0.0, 0.5, 1.0)[ei % 5]      # terminal target exactly on an atom
        term = np.zeros(n, dtype=bool)
        term[-1] = ei % 5 != 3                 # some episodes end without a terminal flag
        ep = [dict(state={"state": t.from_numpy(st[i:i + 1])}, action={"action": t.tensor([[act[i]]])},
                   next_state={"state": t.from_numpy(nst[i:i + 1])}, reward=float(rew[i]), terminal=bool(term[i]))
              for i in range(n)]
        rb.store_episode(ep)
        ep_state.append(st), ep_next.append(nst), ep_act.append(act), ep_rew.append(rew), ep_term.append(term)
        ep_value.append(np.array([tr["value"] for tr in ep], dtype=np.float64))     # rainbow.py:186-196
    out.update(lens=np.array(lens), state=np.concatenate(ep_state), next_state=np.concatenate(ep_next),
               action=np.concatenate(ep_act), reward=np.concatenate(ep_rew), terminal=np.concatenate(ep_term),
               value=np.concatenate(ep_value), tree0=rb.replay_buffer.wt_tree.weights.copy())
    # record what one update saw and produced, without touching its code: wrap the buffer calls around it and
    # Tensor.index_add_ (its second call leaves the projected target distribution in `self`)
    buf = rb.replay_buffer
    real_sample, real_update, real_index_add = buf.sample_batch, buf.update_priority, t.Tensor.index_add_
    seen = {}

    def sample_batch(*a, **k):
        r = real_sample(*a, **k)
        seen["batch"], seen["index"], seen["isw"] = r[1], r[2], r[3]
        return r

    def update_priority(priorities, indexes):
        seen["abs_error"] = np.array(priorities)
        return real_update(priorities, indexes)

    def index_add_(self, *a, **k):
        r = real_index_add(self, *a, **k)
        seen["target_dist"] = self
        return r

    buf.sample_batch, buf.update_priority = sample_batch, update_priority
    seed_all(13)
    for step in range(3):
        s0 = np.random.get_state()
        out[f"u{step}"] = np.random.uniform(size=B)
        np.random.set_state(s0)
        t.Tensor.index_add_ = index_add_
        try:
            loss = rb.update()
        finally:
            t.Tensor.index_add_ = real_index_add
        state, action, value, next_state, terminal, _ = seen["batch"]
        out[f"index{step}"], out[f"isw{step}"] = np.array(seen["index"]), np.array(seen["isw"])
        out[f"b_state{step}"], out[f"b_next{step}"] = state["state"].numpy(), next_state["state"].numpy()
        out[f"b_action{step}"], out[f"b_value{step}"] = action["action"].numpy(), value.numpy()
        out[f"b_terminal{step}"] = terminal.numpy()
        out[f"target_dist{step}"] = seen["target_dist"].view(B, ATOMS).numpy().copy()
        out[f"abs_error{step}"] = seen["abs_error"]
        out[f"loss{step}"] = np.float64(loss)
        out[f"tree{step + 1}"] = buf.wt_tree.weights.copy()
        out[f"q{step + 1}"], out[f"qt{step + 1}"] = _flat_params(rb.qnet), _flat_params(rb.qnet_target)
    # acting (rainbow.py:110-133): expectation over the support, argmax
    probe = rng.randn(9, S).astype(np.float32)
    out["act_state"] = probe
    out["act_online"] = rb.act_discrete({"state": t.from_numpy(probe)}).numpy()
    out["act_target"] = rb.act_discrete({"state": t.from_numpy(probe)}, use_target=True).numpy()
    np.savez_compressed(os.path.join(OUT, "ref_rainbow.npz"), **out)


# --------------------------------------------------------------------------- default configs (generate_config)
def gen_configs():
    """What <Framework>.generate_config produces in the unmodified reference: from an empty config, and from one whose
    frame_config already holds values (the defaults win, dqn.py:530-535)."""
    import json
    from machin.frame.algorithms import A2C, A3C, DDPGApex, DDPGPer, DQN, DQNApex, DQNPer, IMPALA, RAINBOW
    out = {}
    for cls in (DQN, DQNPer, DDPGPer, DQNApex, DDPGApex, IMPALA, A2C, A3C, RAINBOW):
        out[cls.__name__] = {"empty": cls.generate_config({}),
                             "prefilled": cls.generate_config({"frame_config": {"batch_size": 7, "my_own_key": 1}, "other": 2})}
    with open(os.path.join(OUT, "ref_configs.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)            # (Infinity for gradient_max: json's own extension)


# --------------------------------------------------------------------------- DQN with uniform replay (BASELINE config C1)
def gen_dqn():
    """C1's shapes (CartPole: state 4, 2 actions, ring 10 000, batch 32) with synthetic transitions: five updates per
    mode; uniform sampling consumes Python's random stream (buffer.py:234)."""
    from machin.frame.algorithms import DQN
    from machin.model.nets.base import static_module_wrapper as smw

    out = {}
    S, A, N, B = 4, 2, 300, 32
    rng = np.random.RandomState(31)
    st, nst = rng.randn(N, S).astype(np.float32), rng.randn(N, S).astype(np.float32)
    act = rng.randint(0, A, size=N).astype(np.int64)
    rew = rng.choice([0.0, 1.0], size=N)
    term = rng.rand(N) < 0.05
    out.update(state=st, next_state=nst, action=act, reward=rew, terminal=term)
    for mode in ("vanilla", "fixed_target", "double"):
        seed_all(0)
        qnet, qtgt = _QNet(S, A), _QNet(S, A)
        dqn = DQN(smw(qnet, "cpu", "cpu"), smw(qtgt, "cpu", "cpu"), t.optim.SGD, t.nn.MSELoss(), batch_size=B,
                  learning_rate=0.2, discount=0.97, update_rate=0.1, replay_size=10000, mode=mode)
        out[f"{mode}_init_q"], out[f"{mode}_init_qt"] = _flat_params(dqn.qnet), _flat_params(dqn.qnet_target)
        ep = [dict(state={"state": t.from_numpy(st[i:i + 1])}, action={"action": t.tensor([[act[i]]])},
                   next_state={"state": t.from_numpy(nst[i:i + 1])}, reward=float(rew[i]), terminal=bool(term[i]))
              for i in range(N)]
        for lo in range(0, N, 50):                      # six episodes of 50 steps
            dqn.store_episode(ep[lo:lo + 50])
        seed_all(17)
        for step in range(5):
            out[f"{mode}_loss{step}"] = np.float64(dqn.update())
            out[f"{mode}_q{step + 1}"], out[f"{mode}_qt{step + 1}"] = _flat_params(dqn.qnet), _flat_params(dqn.qnet_target)
    np.savez_compressed(os.path.join(OUT, "ref_dqn.npz"), **out)


# --------------------------------------------------------------------------- public surface of the mirrored classes
def gen_surface():
    """Public attributes (methods, properties, class attributes) of every reference class this repository mirrors,
    and the parameter names of their constructors / public methods -- the drop-in boundary of SURVEY 8b as data."""
    import inspect
    import json
    from machin.frame import algorithms as A
    from machin.frame import buffers as B
    from machin.frame.buffers import rnn_buffers as R
    from machin.frame.buffers.prioritized_buffer import WeightTree
    from machin.parallel.server import PushPullGradServer, PushPullModelServer
    classes = {"WeightTree": WeightTree, "Buffer": B.Buffer, "PrioritizedBuffer": B.PrioritizedBuffer,
               "DistributedBuffer": B.DistributedBuffer, "DistributedPrioritizedBuffer": B.DistributedPrioritizedBuffer,
               "RNNBuffer": R.RNNBuffer, "RNNPrioritizedBuffer": R.RNNPrioritizedBuffer,
               "DQN": A.DQN, "DQNPer": A.DQNPer, "DDPGPer": A.DDPGPer, "DQNApex": A.DQNApex, "DDPGApex": A.DDPGApex,
               "IMPALA": A.IMPALA, "A2C": A.A2C, "A3C": A.A3C, "RAINBOW": A.RAINBOW,
               "PushPullModelServer": PushPullModelServer, "PushPullGradServer": PushPullGradServer}
    out = {}
    for name, cls in classes.items():
        members = {}
        for attr in dir(cls):
            if attr.startswith("_") and attr != "__init__":
                continue
            raw = inspect.getattr_static(cls, attr)
            obj = getattr(cls, attr)
            if isinstance(raw, property):
                members[attr] = "property"
            elif callable(obj):
                try:
                    params = [p for p in inspect.signature(obj).parameters if p not in ("self", "cls")]
                except (TypeError, ValueError):
                    params = None
                members[attr] = params
            else:
                members[attr] = "attribute"
        out[name] = members
    with open(os.path.join(OUT, "ref_surface.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


def gen_auto_surface():
    """Public functions / classes / methods of the reference's automation layer (machin/auto) with their parameter names.
    Read from the SOURCE with ``ast``: the package imports pytorch-lightning and gym, which this image does not have."""
    import ast
    import json
    root = os.path.join(ref_shim.REFERENCE_ROOT, "machin", "auto")
    out = {}
    for module in ("config", "dataset", "launcher", os.path.join("envs", "openai_gym")):
        tree = ast.parse(open(os.path.join(root, module + ".py")).read())

        def params(fn):
            args = fn.args
            names = [a.arg for a in args.posonlyargs + args.args + args.kwonlyargs]
            return [n for n in names if n not in ("self", "cls")]

        members = {}
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and not node.name.startswith("_"):
                members[node.name] = params(node)
            elif isinstance(node, ast.ClassDef) and not node.name.startswith("_"):
                members[node.name] = {m.name: params(m) for m in node.body if isinstance(m, ast.FunctionDef)
                                      and (not m.name.startswith("_") or m.name in ("__init__", "__next__", "__len__", "__iter__"))}
        out[module.replace(os.sep, ".")] = members
    with open(os.path.join(OUT, "ref_auto_surface.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


def gen_auto_trace():
    """What the UNMODIFIED launchers of machin/auto/launcher.py do with a batch: the order of ``store_episode`` / ``update``
    calls on the framework, the log entries that reach ``self.log`` and the handler tuples, for the local launcher and for
    the distributed one on the first batch of epoch 0 / a later batch / a later epoch.  pytorch-lightning is not in the
    image: ``LightningModule`` is replaced by a recording ``nn.Module`` stub, and only launcher.py / dataset.py / config.py
    are imported (the package's own __init__ pulls in gym and the lightning plugins)."""
    import importlib.abc
    import importlib.machinery
    import json
    import types
    from unittest.mock import MagicMock
    import torch.nn as nn

    class MockFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        roots = ("pytorch_lightning", "gym", "matplotlib", "PIL", "moviepy")

        def find_spec(self, name, path, target=None):
            if name.split(".")[0] in self.roots and name != "pytorch_lightning":
                return importlib.machinery.ModuleSpec(name, self, is_package=True)
            return None

        def create_module(self, spec):
            m = MagicMock()
            m.__name__, m.__path__, m.__spec__, m.__loader__ = spec.name, [], spec, self
            return m

        def exec_module(self, module):
            pass

    class LightningModule(nn.Module):
        def __init__(self):
            super().__init__()
            self.logged, self.current_epoch, self.device = [], 0, "cpu"

        def log(self, key, value, **kw):
            self.logged.append([key, value, dict(sorted(kw.items()))])

    for name in [k for k in sys.modules if k.split(".")[0] in MockFinder.roots]:
        del sys.modules[name]
    stub = types.ModuleType("pytorch_lightning")
    stub.__path__, stub.LightningModule, stub.Trainer, stub.trainer = [], LightningModule, MagicMock(), MagicMock()
    sys.modules["pytorch_lightning"] = stub
    sys.meta_path.insert(0, MockFinder())
    package = types.ModuleType("machin.auto")
    package.__path__ = [os.path.join(ref_shim.REFERENCE_ROOT, "machin", "auto")]
    sys.modules["machin.auto"] = package
    sys.modules["machin.auto.envs"] = types.ModuleType("machin.auto.envs")
    from machin.auto import dataset, launcher

    class Frame:
        def __init__(self):
            self.events = []

        def store_episode(self, episode):
            self.events.append(["store", len(episode)])

        def update(self):
            self.events.append(["update"])

    class Group:
        def __init__(self, frame):
            self.frame = frame

        def barrier(self):
            self.frame.events.append(["barrier"])

    def handler(module, key, value):
        module.frame.events.append(["handler", key, value])

    def batch():
        first, second = dataset.DatasetResult(), dataset.DatasetResult()
        for i in range(3):
            first.add_observation({"i": i})
        first.add_log({"a": 1.5})
        first.add_log({"v": (7, handler)})
        for i in range(2):
            second.add_observation({"i": i})
        second.add_log({"total_reward": 9.0})
        return [first, second]

    def fresh(cls):
        module = cls.__new__(cls)
        LightningModule.__init__(module)
        module.frame = Frame()
        return module

    out = {}
    local = fresh(launcher.Launcher)
    local.training_step(batch(), 0)
    out["launcher_training_step"] = {"events": local.frame.events, "logged": local.logged}
    local = fresh(launcher.Launcher)
    local.test_step(batch(), 0)
    out["launcher_test_step"] = {"events": local.frame.events, "logged": local.logged}
    for name, epoch, batch_idx in (("first_batch", 0, 0), ("later_batch", 0, 1), ("later_epoch", 1, 0)):
        d = fresh(launcher.DistributedLauncher)
        d.group, d._train_esm, d._test_esm, d.current_epoch = Group(d.frame), "total_reward", None, epoch
        d.training_step(list(reversed(batch())), batch_idx)       # (the monitored key sits in the FIRST sample)
        out["distributed_training_step_" + name] = {"events": d.frame.events, "logged": d.logged}
    d = fresh(launcher.DistributedLauncher)
    d.group, d._train_esm, d._test_esm = Group(d.frame), "total_reward", None
    d.test_step(batch(), 0)
    out["distributed_test_step"] = {"events": d.frame.events, "logged": d.logged}
    with open(os.path.join(OUT, "ref_auto_trace.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


class AutoQNet(t.nn.Module):
    """Q-net of the dataset fixture (tests/test_auto_cpu.py defines the same class: same layers in the same order)."""

    def __init__(self, state_dim, action_num):
        super().__init__()
        self.fc1, self.fc2 = t.nn.Linear(state_dim, 16), t.nn.Linear(16, action_num)

    def forward(self, state):
        return self.fc2(t.relu(self.fc1(state)))


class AutoSoftActor(t.nn.Module):
    def __init__(self, state_dim, action_dim):
        super().__init__()
        self.fc = t.nn.Linear(state_dim, action_dim)

    def forward(self, state):
        return t.softmax(self.fc(state), dim=1)


class AutoCritic(t.nn.Module):
    def __init__(self, state_dim, action_dim=0):
        super().__init__()
        self.fc = t.nn.Linear(state_dim + action_dim, 1)

    def forward(self, state, action=None):
        return self.fc(state if action is None else t.cat([state, action], dim=1))


class AutoPolicy(t.nn.Module):
    def __init__(self, state_dim, action_num):
        super().__init__()
        self.fc = t.nn.Linear(state_dim, action_num)

    def forward(self, state, action=None):
        law = t.distributions.Categorical(logits=self.fc(state))
        act = action if action is not None else law.sample().view(-1, 1)
        return act, law.log_prob(act.flatten()).view(-1, 1), law.entropy().view(-1, 1)


def gen_auto_dataset():
    """Two episodes collected by the UNMODIFIED ``RLGymDiscActDataset`` (machin/auto/envs/openai_gym.py:32-140) with the
    reference's DQN acting (epsilon 0.5, so that both the greedy and the random branch occur) on the cart-pole of
    machin_b200/auto/envs/classic_control.py behind a stub ``gym`` (spaces + a render method).  The fixture holds every
    transition; machin_b200's dataset must collect the same ones from the same seeds."""
    import importlib.abc
    import importlib.machinery
    import types
    from unittest.mock import MagicMock
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from machin_b200.auto.envs import classic_control

    class Space:
        def __init__(self, *a, **k):
            pass

    gym = types.ModuleType("gym")
    gym.__path__ = []
    spaces = types.ModuleType("gym.spaces")
    for name in ("Box", "Discrete", "MultiBinary", "MultiDiscrete"):
        setattr(spaces, name, type(name, (Space,), {}))
    rendering = types.ModuleType("gym.envs.classic_control.rendering")
    rendering.Viewer = type("Viewer", (), {"__init__": lambda self, *a, **k: None})
    gym.spaces = spaces
    for name in [k for k in sys.modules if k.split(".")[0] in ("gym", "pytorch_lightning", "matplotlib", "PIL", "moviepy")]:
        del sys.modules[name]
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.envs": types.ModuleType("gym.envs"),
                        "gym.envs.classic_control": types.ModuleType("gym.envs.classic_control"),
                        "gym.envs.classic_control.rendering": rendering})
    sys.modules["gym.envs.classic_control"].rendering = rendering

    class MockFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        roots = ("pytorch_lightning", "matplotlib", "PIL", "moviepy", "gym")

        def find_spec(self, name, path, target=None):
            if name.split(".")[0] in self.roots and name not in sys.modules and name != "pytorch_lightning":
                return importlib.machinery.ModuleSpec(name, self, is_package=True)
            return None

        def create_module(self, spec):
            m = MagicMock()
            m.__name__, m.__path__, m.__spec__, m.__loader__ = spec.name, [], spec, self
            return m

        def exec_module(self, module):
            pass

    class LightningModule(t.nn.Module):
        pass

    stub = types.ModuleType("pytorch_lightning")
    stub.__path__, stub.LightningModule, stub.Trainer, stub.trainer = [], LightningModule, MagicMock(), MagicMock()
    sys.modules["pytorch_lightning"] = stub
    sys.meta_path.insert(0, MockFinder())
    for name in [k for k in sys.modules if k.startswith("machin.auto")]:
        del sys.modules[name]
    from machin.auto.config import generate_algorithm_config, init_algorithm_from_config
    from machin.auto.envs.openai_gym import RLGymDiscActDataset

    class Env(classic_control.CartPole):
        def __init__(self, seed):
            super().__init__(seed=seed)
            self.action_space, self.observation_space = spaces.Discrete(), spaces.Box()

        def render(self, mode="rgb_array"):
            return np.zeros((2, 2, 3), dtype=np.uint8)

    globals()["QNet"] = AutoQNet                      # the reference resolves model names in the globals of the call stack
    config = generate_algorithm_config("DQN", {})
    config["frame_config"]["models"] = ["QNet", "QNet"]
    config["frame_config"]["model_kwargs"] = [{"state_dim": 4, "action_num": 2}] * 2
    t.manual_seed(11)
    dqn = init_algorithm_from_config(config)
    dqn.epsilon = 0.5
    dataset = RLGymDiscActDataset(dqn, Env(seed=5), render_every_episode=10 ** 9)
    out = {}
    t.manual_seed(12)
    for e in range(2):
        result = next(dataset)
        obs = result.observations
        out[f"ep{e}_state"] = np.concatenate([o["state"]["state"].numpy() for o in obs])
        out[f"ep{e}_next_state"] = np.concatenate([o["next_state"]["state"].numpy() for o in obs])
        out[f"ep{e}_action"] = np.concatenate([o["action"]["action"].numpy() for o in obs])
        out[f"ep{e}_reward"] = np.array([o["reward"] for o in obs], dtype=np.float64)
        out[f"ep{e}_terminal"] = np.array([o["terminal"] for o in obs], dtype=np.bool_)
        out[f"ep{e}_total_reward"] = np.float64([log["total_reward"] for log in result.logs if "total_reward" in log][-1])
    out["epsilon_after"] = np.float64(dqn.epsilon)
    # the other two branches of the discrete dataset: the DDPG family (the probability row is what gets stored,
    # openai_gym.py:121-122) and the policy-gradient family (``act(...)[0]``), one episode each
    globals().update(SoftActor=AutoSoftActor, Critic=AutoCritic, Policy=AutoPolicy)
    for tag, algo, models, kwargs in (
            ("ddpg", "DDPGPer", ["SoftActor", "SoftActor", "Critic", "Critic"], [{"state_dim": 4, "action_dim": 2}] * 4),
            ("a2c", "A2C", ["Policy", "Critic"], [{"state_dim": 4, "action_num": 2}, {"state_dim": 4}])):
        config = generate_algorithm_config(algo, {})
        config["frame_config"]["models"], config["frame_config"]["model_kwargs"] = models, kwargs
        t.manual_seed(21)
        frame = init_algorithm_from_config(config)
        dataset = RLGymDiscActDataset(frame, Env(seed=6), render_every_episode=10 ** 9)
        t.manual_seed(22)
        obs = next(dataset).observations
        out[f"{tag}_state"] = np.concatenate([o["state"]["state"].numpy() for o in obs])
        out[f"{tag}_action"] = np.concatenate([o["action"]["action"].numpy() for o in obs])
        out[f"{tag}_terminal"] = np.array([o["terminal"] for o in obs], dtype=np.bool_)
    np.savez_compressed(os.path.join(OUT, "ref_auto_dataset.npz"), **out)


if __name__ == "__main__":
    ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["tree", "per", "td", "vtrace", "gae", "buffer", "per_d", "rnn", "rainbow", "configs", "dqn", "surface"]
    for name in which:
        globals()["gen_" + name]()
        print("wrote", name)
