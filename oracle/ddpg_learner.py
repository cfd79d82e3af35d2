"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- CPU restatement of the reference APEX-DDPG learner step.

One ``step()`` is what ``DDPGPer.update()`` does on the reference's CPU path
(machin/frame/algorithms/ddpg_per.py:92-185) over a local prioritized buffer:

  PrioritizedBuffer.sample_batch (NumPy tree, per-transition Python objects, one cat per attribute)
                                                        prioritized_buffer.py:330-428, buffer.py:380-413
  y = r + gamma * (1 - terminal) * Q'(s', mu'(s'))       ddpg_per.py:117-128, ddpg.py:506-510
  IS-weighted critic loss, critic step                   :130-146
  actor loss -Q(s, mu(s)), |Q(s, mu(s)) - y| -> update_priority, actor step   :148-171
  soft update of both targets                            :173-181

The sum tree has the full configured shard size; the Python-object storage is capped at ``storage_rows`` and sampled
slots are folded onto it (the per-transition assembly cost does not depend on the storage size).  Nothing under
machin_b200/ imports this module; bench.py uses it as the CPU arm of workload C4 only.
"""
import numpy as np
import torch as t

from . import per as oper
from . import td as otd
from .sumtree import OracleTree


class ActorMLP(t.nn.Module):
    def __init__(self, s, a, h=256):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = t.nn.Linear(s, h), t.nn.Linear(h, h), t.nn.Linear(h, a)

    def forward(self, state):
        return t.tanh(self.fc3(t.relu(self.fc2(t.relu(self.fc1(state))))))


class CriticMLP(t.nn.Module):
    def __init__(self, s, a, h=256):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = t.nn.Linear(s + a, h), t.nn.Linear(h, h), t.nn.Linear(h, 1)

    def forward(self, state, action):
        return self.fc3(t.relu(self.fc2(t.relu(self.fc1(t.cat([state, action], dim=1))))))


class CpuDdpgPerLearner:
    def __init__(self, tree_size, storage_rows, state_dim, action_dim, batch_size, hidden=256, discount=0.99,
                 update_rate=0.005, seed=0):
        rng = np.random.RandomState(seed)
        self.batch_size, self.discount, self.update_rate = batch_size, discount, update_rate
        self.n_stored, self.rows = tree_size, storage_rows
        self.tree = OracleTree(tree_size)
        self.tree.update_all_leaves(oper.normalize_priority(rng.exponential(size=tree_size)))
        self.storage = [dict(state=t.from_numpy(rng.randn(1, state_dim).astype(np.float32)),
                             action=t.from_numpy((rng.rand(1, action_dim) * 2 - 1).astype(np.float32)),
                             next_state=t.from_numpy(rng.randn(1, state_dim).astype(np.float32)),
                             reward=float(np.float32(rng.randn())), terminal=bool(rng.rand() < 0.01))
                        for _ in range(storage_rows)]
        t.manual_seed(seed)
        self.actor, self.actor_target = ActorMLP(state_dim, action_dim, hidden), ActorMLP(state_dim, action_dim, hidden)
        self.critic, self.critic_target = CriticMLP(state_dim, action_dim, hidden), CriticMLP(state_dim, action_dim, hidden)
        self.actor_target.load_state_dict(self.actor.state_dict())
        self.critic_target.load_state_dict(self.critic.state_dict())
        self.actor_optim = t.optim.Adam(self.actor.parameters(), lr=5e-4)
        self.critic_optim = t.optim.Adam(self.critic.parameters(), lr=1e-3)
        self.beta = 0.4

    def step(self):
        B = self.batch_size
        index, isw = oper.sample_index_and_weight(self.tree, B, self.n_stored, self.beta)
        self.beta = min(1.0, self.beta + 0.001)
        batch = [self.storage[i % self.rows] for i in index]
        state = t.cat([b["state"] for b in batch], dim=0)
        action = t.cat([b["action"] for b in batch], dim=0)
        next_state = t.cat([b["next_state"] for b in batch], dim=0)
        reward = t.tensor([b["reward"] for b in batch]).view(B, -1)
        terminal = t.tensor([b["terminal"] for b in batch]).view(B, -1)
        with t.no_grad():
            next_value = self.critic_target(next_state, self.actor_target(next_state)).view(B, -1)
            y = otd.ddpg_target(next_value, reward, terminal, self.discount)
        cur_value = self.critic(state, action)
        value_loss = otd.per_weighted_loss((cur_value - y) ** 2, t.from_numpy(isw))
        self.critic.zero_grad()
        value_loss.backward()
        t.nn.utils.clip_grad_norm_(self.critic.parameters(), np.inf)
        self.critic_optim.step()
        act_value = self.critic(state, self.actor(state))
        policy_loss = -act_value.mean()
        abs_error = otd.abs_error(act_value.detach(), y)
        self.tree.update_leaf_batch(oper.normalize_priority(abs_error.numpy()), index)
        self.actor.zero_grad()
        policy_loss.backward()
        t.nn.utils.clip_grad_norm_(self.actor.parameters(), np.inf)
        self.actor_optim.step()
        otd.soft_update_(list(self.actor_target.parameters()), list(self.actor.parameters()), self.update_rate)
        otd.soft_update_(list(self.critic_target.parameters()), list(self.critic.parameters()), self.update_rate)
        return -policy_loss.item(), value_loss.item()
