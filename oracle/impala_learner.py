"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- CPU restatement of the reference IMPALA learner step.

One ``step()`` is what ``IMPALA.update()`` does on the reference's CPU path
(machin/frame/algorithms/impala.py:242-397) for a learner whose buffer already holds the episodes:

  DistributedBuffer.sample_batch(sample_method="episode")   buffer_d.py:131-198 -> buffer.py:213-260, 380-413
    random.choices over the stored episodes, per-transition Python objects,
    post_process_batch: one torch.cat per attribute
  re-evaluate pi(a|s), V(s), V(s')                           impala.py:315-332
  V-trace: the per-step scalar recurrence                    impala.py:340-362  (oracle/vtrace.py: vtrace_loop)
  policy / value losses, backward, clip, both optimizer steps   impala.py:364-389

Nothing under machin_b200/ imports this module; bench.py uses it as the CPU arm of workloads C3 / C5 only.
"""
import random

import numpy as np
import torch as t

from . import vtrace as ovt


class PolicyMLP(t.nn.Module):
    def __init__(self, s, a, h=256):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = t.nn.Linear(s, h), t.nn.Linear(h, h), t.nn.Linear(h, a)

    def forward(self, state, action=None):
        logp_all = t.log_softmax(self.fc3(t.relu(self.fc2(t.relu(self.fc1(state))))), dim=1)
        act = action if action is not None else t.multinomial(logp_all.exp(), 1)
        logp = logp_all.gather(1, act.view(-1, 1).long())
        ent = -(logp_all.exp() * logp_all).sum(1, keepdim=True)
        return act.view(-1, 1), logp, ent


class ValueMLP(t.nn.Module):
    def __init__(self, s, h=256):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = t.nn.Linear(s, h), t.nn.Linear(h, h), t.nn.Linear(h, 1)

    def forward(self, state):
        return self.fc3(t.relu(self.fc2(t.relu(self.fc1(state)))))


class CpuImpalaLearner:
    def __init__(self, n_episodes, episode_len, state_dim, n_actions, batch_episodes, hidden=256, discount=0.99,
                 clip_c=1.0, clip_rho=1.0, lr=1e-3, seed=0):
        rng = np.random.RandomState(seed)
        self.batch, self.T, self.discount, self.clip_c, self.clip_rho = batch_episodes, episode_len, discount, clip_c, clip_rho
        # one list of per-transition dicts per episode, like TransitionStorageBasic + Buffer.episode_transition_handles
        self.episodes = []
        for _ in range(n_episodes):
            ep = []
            for k in range(episode_len):
                ep.append(dict(state=t.from_numpy(rng.randn(1, state_dim).astype(np.float32)),
                               action=t.tensor([[int(rng.randint(0, n_actions))]]),
                               next_state=t.from_numpy(rng.randn(1, state_dim).astype(np.float32)),
                               reward=float(np.float32(rng.randn())), terminal=k == episode_len - 1,
                               action_log_prob=float(-rng.rand()), episode_length=episode_len if k == 0 else 0))
            self.episodes.append(ep)
        t.manual_seed(seed)
        self.actor, self.critic = PolicyMLP(state_dim, n_actions, hidden), ValueMLP(state_dim, hidden)
        self.actor_optim = t.optim.Adam(self.actor.parameters(), lr=lr)
        self.critic_optim = t.optim.Adam(self.critic.parameters(), lr=lr)

    def sample(self):
        """buffer.py:213-260 (sample_method_episode) + :380-413 (post_process_batch with concatenation)."""
        picked = random.choices(self.episodes, k=self.batch)
        batch = [tr for ep in picked for tr in ep]
        state = t.cat([tr["state"] for tr in batch], dim=0)
        action = t.cat([tr["action"] for tr in batch], dim=0)
        next_state = t.cat([tr["next_state"] for tr in batch], dim=0)
        reward = t.tensor([tr["reward"] for tr in batch], dtype=t.float32).view(-1, 1)
        terminal = t.tensor([tr["terminal"] for tr in batch], dtype=t.bool).view(-1, 1)
        logmu = t.tensor([tr["action_log_prob"] for tr in batch], dtype=t.float32).view(-1, 1)
        lengths = [tr["episode_length"] for tr in batch if tr["episode_length"] != 0]
        return state, action, reward, next_state, terminal, logmu, lengths

    def step(self, vtrace_episode_limit=None):
        """One update.  ``vtrace_episode_limit``: run the scalar V-trace loop on the first k episodes only and the
        vectorised restatement (same float32 arithmetic) on the rest -- for a time-bounded baseline sample; the
        caller scales the loop time back up.  Returns (policy value, value loss, seconds spent in the loop)."""
        import time
        self.actor.train()
        self.critic.train()
        state, action, reward, next_state, terminal, logmu, lengths = self.sample()
        n = sum(lengths)
        _, logp, _ent = self.actor(state, action)
        logp = logp.view(n, 1)
        value = self.critic(state).view(n, 1)
        next_value = self.critic(next_state).view(n, 1)
        c, rho, delta = ovt.vtrace_prepare(logp, logmu, value, next_value, reward, terminal, self.discount,
                                           self.clip_c, self.clip_rho)
        with t.no_grad():
            t0 = time.perf_counter()
            k = len(lengths) if vtrace_episode_limit is None else min(vtrace_episode_limit, len(lengths))
            cut = sum(lengths[:k])
            vs_a, vss_a = ovt.vtrace_loop(value[:cut], delta[:cut], c[:cut], lengths[:k], self.discount)
            loop_s = time.perf_counter() - t0
            if k < len(lengths):
                vs_b, vss_b = ovt.vtrace_batched(value[cut:], delta[cut:], c[cut:], lengths[k:], self.discount)
                vs, vss = t.cat([vs_a, vs_b]), t.cat([vss_a, vss_b])
            else:
                vs, vss = vs_a, vss_a
        pol, val = ovt.impala_losses(rho, logp, reward, vss, value, vs, self.discount)
        self.actor.zero_grad()
        pol.backward()
        t.nn.utils.clip_grad_norm_(self.actor.parameters(), np.inf)
        self.actor_optim.step()
        self.critic.zero_grad()
        val.backward()
        t.nn.utils.clip_grad_norm_(self.critic.parameters(), np.inf)
        self.critic_optim.step()
        self.actor.eval()
        self.critic.eval()
        return -pol.item(), val.item(), loop_s
