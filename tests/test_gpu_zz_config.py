"""GPU: frameworks built from configs (generate_config / init_from_config) train on the device like hand-built ones.
The tables and resolvers are pinned on the CPU (tests/test_config_cpu.py); this checks the constructed objects against
real buffers: one store + update per framework, outside torch.distributed (every group is a group of one)."""
import numpy as np
import pytest
import torch as t


from test_config_cpu import Actor, Critic, QNet  # noqa: F401  (resolved by name from this module's globals)

pytestmark = [pytest.mark.gpu]
S, A = 4, 3


class DistQ(t.nn.Module):
    def __init__(self, atoms=11):
        super().__init__()
        self.fc, self.atoms = t.nn.Linear(S, A * atoms), atoms

    def forward(self, state):
        return t.softmax(self.fc(state).view(-1, A, self.atoms), dim=-1)


class Policy(t.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = t.nn.Linear(S, A)

    def forward(self, state, action=None):
        dist = t.distributions.Categorical(logits=self.fc(state))
        act = action if action is not None else dist.sample().view(-1, 1)
        return act, dist.log_prob(act.flatten()).view(-1, 1), dist.entropy().view(-1, 1)


def _episode(n, discrete=True, extra=None):
    rng = np.random.RandomState(n)
    return [dict(state={"state": t.from_numpy(rng.randn(1, S).astype(np.float32))},
                 action={"action": t.tensor([[int(rng.randint(A))]]) if discrete else t.from_numpy(rng.uniform(-1, 1, (1, 2)).astype(np.float32))},
                 next_state={"state": t.from_numpy(rng.randn(1, S).astype(np.float32))}, reward=float(rng.randn()),
                 terminal=i == n - 1, **(extra or {})) for i in range(n)]


def test_value_based_frameworks_from_config():
    from machin_b200.frame.algorithms import DQNApex, DQNPer, RAINBOW
    for cls, models in ((DQNPer, ["QNet", "QNet"]), (DQNApex, ["QNet", "QNet"]), (RAINBOW, [DistQ, DistQ])):
        cfg = cls.generate_config({})
        cfg["frame_config"].update(models=models, batch_size=16, replay_size=1000,
                                   model_kwargs=[{"state_dim": S, "action_num": A}] * 2 if models[0] == "QNet" else [{}, {}])
        algo = cls.init_from_config(cfg)
        assert next(algo.qnet.parameters()).is_cuda
        for n in (5, 40):
            algo.store_episode(_episode(n))
        loss = algo.update()
        assert np.isfinite(loss)
        assert algo.act_discrete({"state": t.zeros(2, S, device="cuda")}).shape == (2, 1)


def test_actor_critic_frameworks_from_config():
    from machin_b200.frame.algorithms import A2C, A3C, DDPGApex, DDPGPer, IMPALA
    for cls in (DDPGPer, DDPGApex):
        cfg = cls.generate_config({})
        cfg["frame_config"].update(model_kwargs=[{"state_dim": S, "action_dim": 2}] * 4, batch_size=16, replay_size=1000)
        algo = cls.init_from_config(cfg)
        algo.store_episode(_episode(40, discrete=False))
        assert all(np.isfinite(x) for x in algo.update())
    for cls in (A2C, A3C):
        cfg = cls.generate_config({})
        cfg["frame_config"].update(models=[Policy, "Critic"], model_kwargs=[{}, {"state_dim": S}], batch_size=8, replay_size=1000)
        algo = cls.init_from_config(cfg)
        algo.store_episode(_episode(30))
        assert all(np.isfinite(x) for x in algo.update())
    cfg = IMPALA.generate_config({})
    cfg["frame_config"].update(models=[Policy, "Critic"], model_kwargs=[{}, {"state_dim": S}], batch_size=2, replay_size=500)
    imp = IMPALA.init_from_config(cfg)
    for n in (6, 9):
        imp.store_episode(_episode(n, extra={"action_log_prob": -1.0}))
    assert all(np.isfinite(x) for x in imp.update())
