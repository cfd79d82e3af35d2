"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- the UNMODIFIED reference learner (iffiX/machin's own
``DQNPer`` over its own ``PrioritizedBuffer``, machin/frame/algorithms/dqn_per.py:89-161) set up on BASELINE config C2
for bench.py's reference arm.  The code that is timed is the reference's: this module only builds the networks and fills
the buffer.

The prioritized buffer is created at the full configured size (the reference's NumPy sum tree gets its 2^25 - 1 nodes,
so every draw / write-back walks the full depth); the number of stored transitions is capped at ``storage_rows`` --
10 M ``Transition`` objects of four tensors each are impractical in the reference itself and the
per-transition assembly cost of a batch does not depend on how many are stored.
"""
import numpy as np
import torch as t

from . import ref_shim


class QNet(t.nn.Module):
    def __init__(self, s, a, h=256):
        super().__init__()
        self.fc1, self.fc2, self.fc3 = t.nn.Linear(s, h), t.nn.Linear(h, h), t.nn.Linear(h, a)

    def forward(self, state):
        return self.fc3(t.relu(self.fc2(t.relu(self.fc1(state)))))


class ReferenceDqnPerLearner:
    def __init__(self, tree_size, storage_rows, state_dim, n_actions, batch_size, hidden=256, seed=0):
        ref_shim.load()
        from machin.frame.algorithms import DQNPer
        rng = np.random.RandomState(seed)
        t.manual_seed(seed)
        self.dqn = DQNPer(QNet(state_dim, n_actions, hidden), QNet(state_dim, n_actions, hidden), t.optim.Adam,
                          t.nn.MSELoss(reduction="sum"), batch_size=batch_size, learning_rate=1e-3, discount=0.99,
                          update_rate=0.005, replay_size=tree_size, replay_device="cpu")
        done = 0
        while done < storage_rows:
            n = min(1000, storage_rows - done)
            episode = [dict(state={"state": t.from_numpy(rng.randn(1, state_dim).astype(np.float32))},
                            action={"action": t.tensor([[int(rng.randint(0, n_actions))]])},
                            next_state={"state": t.from_numpy(rng.randn(1, state_dim).astype(np.float32))},
                            reward=float(np.float32(rng.randn())), terminal=bool(rng.rand() < 0.01)) for _ in range(n)]
            self.dqn.replay_buffer.store_episode(episode, priorities=list(rng.exponential(size=n)))
            done += n
        self.stored = done

    def step(self):
        return self.dqn.update()
