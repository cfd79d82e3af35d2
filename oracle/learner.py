"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- CPU restatement of the reference APEX-DQN learner step.

One ``step()`` is what ``DQNPer.update()`` does on the reference's CPU path
(machin/frame/algorithms/dqn_per.py:89-161):

  PrioritizedBuffer.sample_batch            prioritized_buffer.py:330-395
    sample_index_and_weight (NumPy tree)    :397-428
    [storage[i] for i in index]             :391         (list of per-transition Python objects)
    post_process_batch: per attribute cat   buffer.py:261-353,380-413
  double-DQN target, IS-weighted loss       dqn_per.py:105-130
  abs error -> update_priority              :132-139 -> prioritized_buffer.py:319-328
  backward, clip, optimizer step, soft update   dqn_per.py:144-157

The sum tree has the full configured size; the Python-object storage is capped at
``storage_rows`` (10 M ``Transition`` objects are impractical in the reference as well --
BASELINE.md section 3) and sampled slots are folded onto it with ``index % storage_rows``; the
per-transition assembly cost does not depend on the storage size.
"""
import numpy as np
import torch as t

from . import per as oper
from . import td as otd
from .sumtree import OracleTree


class MLP(t.nn.Module):
    def __init__(self, s, a, h=256):
        super().__init__()
        self.fc1 = t.nn.Linear(s, h)
        self.fc2 = t.nn.Linear(h, h)
        self.fc3 = t.nn.Linear(h, a)

    def forward(self, state):
        return self.fc3(t.relu(self.fc2(t.relu(self.fc1(state)))))


class CpuDqnPerLearner:
    def __init__(self, tree_size, storage_rows, state_dim, n_actions, batch_size, hidden=256, discount=0.99,
                 update_rate=0.005, lr=1e-3, seed=0):
        rng = np.random.RandomState(seed)
        self.batch_size, self.discount, self.update_rate = batch_size, discount, update_rate
        self.n_stored = tree_size
        self.tree = OracleTree(tree_size)
        self.tree.update_all_leaves(oper.normalize_priority(rng.exponential(size=tree_size)))
        self.rows = storage_rows
        # one dict of [1, d] tensors per transition, like TransitionStorageBasic.data (storage.py:85-114)
        self.storage = [dict(state=t.from_numpy(rng.randn(1, state_dim).astype(np.float32)),
                             action=t.tensor([[int(rng.randint(0, n_actions))]]),
                             next_state=t.from_numpy(rng.randn(1, state_dim).astype(np.float32)),
                             reward=float(np.float32(rng.randn())), terminal=bool(rng.rand() < 0.01))
                        for _ in range(storage_rows)]
        t.manual_seed(seed)
        self.qnet, self.qnet_target = MLP(state_dim, n_actions, hidden), MLP(state_dim, n_actions, hidden)
        self.qnet_target.load_state_dict(self.qnet.state_dict())
        self.optim = t.optim.Adam(self.qnet.parameters(), lr=lr)
        self.beta = 0.4

    def sample_only(self):
        """PrioritizedBuffer.sample_batch: index draw + batch assembly."""
        B = self.batch_size
        index, isw = oper.sample_index_and_weight(self.tree, B, self.n_stored, self.beta)
        self.beta = min(1.0, self.beta + 0.001)
        batch = [self.storage[i % self.rows] for i in index]
        state = t.cat([b["state"] for b in batch], dim=0)
        action = t.cat([b["action"] for b in batch], dim=0)
        next_state = t.cat([b["next_state"] for b in batch], dim=0)
        reward = t.tensor([b["reward"] for b in batch]).view(B, -1)
        terminal = t.tensor([b["terminal"] for b in batch]).view(B, -1)
        return state, action, reward, next_state, terminal, index, isw

    def step(self):
        state, action, reward, next_state, terminal, index, isw = self.sample_only()
        with t.no_grad():
            qn = self.qnet(next_state)
            qtn = self.qnet_target(next_state)
        q = self.qnet(state)
        av, y, abs_err = otd.dqn_per_target(q, qn, qtn, action, reward, terminal, self.discount)
        loss = otd.per_weighted_loss((av - y) ** 2, t.from_numpy(isw))
        self.tree.update_leaf_batch(oper.normalize_priority(abs_err.detach().numpy()), index)
        self.qnet.zero_grad()
        loss.backward()
        t.nn.utils.clip_grad_norm_(self.qnet.parameters(), np.inf)
        self.optim.step()
        otd.soft_update_(list(self.qnet_target.parameters()), list(self.qnet.parameters()), self.update_rate)
        return loss.item()
