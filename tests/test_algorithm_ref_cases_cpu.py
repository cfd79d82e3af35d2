"""CPU: the case list of the reference's own algorithm tests (test/frame/algorithms/test_dqn_per.py, test_ddpg_per.py,
test_rainbow.py, test_a2c.py: criterion check, act, store, update with every flag combination, lr scheduler, save / load,
config init), run through the product's classes as DRY RUNS -- buffers over host-memory doubles, the library as the
signature-checking recorder -- so the cases exercise the Python surface exactly as a user of the reference would
(int32 actions, integer rewards, zeros everywhere).  Numbers are checked elsewhere (fixtures of the unmodified
reference in tests/test_algorithms_over_doubles_cpu.py and on the device)."""
import itertools

import numpy as np
import pytest
import torch as t

from test_config_cpu import Actor, Critic, QNet, doubles  # noqa: F401  (doubles: fixture)

S, A = 4, 2
# (the reference's lr-scheduler case steps the scheduler before any optimizer step, test_dqn.py:test_lr_scheduler)
pytestmark = pytest.mark.filterwarnings("ignore:Detected call of `lr_scheduler.step")


def _steps(n, action):
    old_state = state = t.zeros([1, S], dtype=t.float32)
    return [{"state": {"state": old_state}, "action": {"action": action}, "next_state": {"state": state}, "reward": 0,
             "terminal": False} for _ in range(n)]


class DistQ(t.nn.Module):
    def __init__(self, atoms=10):
        super().__init__()
        self.fc, self.atoms = t.nn.Linear(S, A * atoms), atoms

    def forward(self, state):
        return t.softmax(self.fc(state).view(-1, A, self.atoms), dim=-1)


def _value_based(kind, dev="cpu", **kw):
    from machin_b200.frame.algorithms import DQN, DQNPer, RAINBOW
    if kind == "RAINBOW":
        return RAINBOW(DistQ().to(dev), DistQ().to(dev), t.optim.Adam, 0, 40, reward_future_steps=20, replay_size=100, **kw)
    cls = {"DQN": DQN, "DQNPer": DQNPer}[kind]
    return cls(QNet(S, A).to(dev), QNet(S, A).to(dev), t.optim.Adam, t.nn.MSELoss(reduction="sum"), replay_size=100, **kw)


@pytest.mark.parametrize("kind", ["DQNPer", "DQN", "RAINBOW"])
def test_value_based_reference_cases(kind, doubles, tmp_path):  # noqa: F811
    value_based_cases(kind, tmp_path, "cpu")


def value_based_cases(kind, tmp_path, dev):
    algo = _value_based(kind, dev)
    # act: [1, 1] actions, epsilon decays only with noise (test_dqn.py:test_act)
    state = {"state": t.zeros([1, S])}
    assert algo.act_discrete(state).shape == (1, 1) and algo.act_discrete(state, True).shape == (1, 1)
    eps = algo.epsilon
    assert algo.act_discrete_with_noise(state).shape == (1, 1) and algo.epsilon < eps
    assert algo.act_discrete_with_noise(state, True, decay_epsilon=False).shape == (1, 1)
    # store + update with every flag combination; int32 actions and integer rewards like the reference's cases
    algo.store_episode(_steps(3, t.zeros([1, 1], dtype=t.int)))
    for update_value, update_target in itertools.product((True, False), repeat=2):
        loss = algo.update(update_value=update_value, update_target=update_target, concatenate_samples=True)
        assert np.isfinite(loss)
    algo.store_episode(_steps(3, t.zeros([1, 1], dtype=t.int)))
    assert np.isfinite(algo.update())
    # save / load (test_dqn.py:test_save_load): restorable = the target net, loading copies it into the online net
    algo.save(str(tmp_path), network_map={"qnet_target": "qt"}, version=1000)
    with t.no_grad():
        for p in algo.qnet_target.parameters():
            p.add_(1.0)
    algo.load(str(tmp_path), network_map={"qnet_target": "qt"}, version=1000)
    assert all(t.equal(a, b) for a, b in zip(algo.qnet.parameters(), algo.qnet_target.parameters()))
    # lr scheduler (test_dqn.py:test_lr_scheduler)
    sched = _value_based(kind, dev, lr_scheduler=t.optim.lr_scheduler.LambdaLR, lr_scheduler_args=((lambda step: 0.5 ** step,),))
    before = sched.qnet_optim.param_groups[0]["lr"]
    sched.update_lr_scheduler()
    assert sched.qnet_optim.param_groups[0]["lr"] == before * 0.5 and len(sched.lr_schedulers) == 1


def test_criterion_and_mode_checks(doubles):  # noqa: F811
    from machin_b200.frame.algorithms import DDPGPer, DQN, DQNPer
    with pytest.raises(RuntimeError, match="Criterion does not have the 'reduction' property"):
        DQNPer(QNet(S, A), QNet(S, A), t.optim.Adam, lambda a, b: a - b, replay_size=10)
    with pytest.raises(RuntimeError, match="Criterion does not have the 'reduction' property"):
        DDPGPer(Actor(S, 2), Actor(S, 2), Critic(S, 2), Critic(S, 2), t.optim.Adam, lambda a, b: a - b, replay_size=10)
    with pytest.raises(ValueError, match="update_rate or update_steps"):
        DQNPer(QNet(S, A), QNet(S, A), t.optim.Adam, t.nn.MSELoss(), update_rate=0.1, update_steps=5, replay_size=10)
    with pytest.raises(ValueError, match="Unknown DQN mode"):
        DQN(QNet(S, A), QNet(S, A), t.optim.Adam, t.nn.MSELoss(), mode="invalid_mode", replay_size=10)
    keep = DQN(QNet(S, A), QNet(S, A), t.optim.Adam, lambda a, b: ((a - b) ** 2).mean(), replay_size=10)     # any callable
    keep.store_episode(_steps(3, t.zeros([1, 1], dtype=t.int)))
    assert np.isfinite(keep.update())
    hard = DQNPer(QNet(S, A), QNet(S, A), t.optim.Adam, t.nn.MSELoss(), update_rate=None, update_steps=2, replay_size=10)
    hard.store_episode(_steps(3, t.zeros([1, 1], dtype=t.int)))
    for _ in range(2):
        hard.update()
    assert hard._update_counter == 2 and all(t.equal(a, b) for a, b in zip(hard.qnet.parameters(), hard.qnet_target.parameters()))


def test_ddpg_per_reference_cases(doubles, tmp_path):  # noqa: F811
    ddpg_cases(tmp_path, "cpu")


def ddpg_cases(tmp_path, dev):
    from machin_b200.frame.algorithms import DDPGPer
    nets = lambda: (Actor(S, 2).to(dev), Actor(S, 2).to(dev), Critic(S, 2).to(dev), Critic(S, 2).to(dev))
    ddpg = DDPGPer(*nets(), t.optim.Adam, t.nn.MSELoss(reduction="sum"), replay_size=100)
    state = {"state": t.zeros([1, S])}
    assert ddpg.act(state).shape == (1, 2) and ddpg.act(state, use_target=True).shape == (1, 2)
    assert ddpg.act_with_noise(state, noise_param=(0.0, 1.0), mode="uniform").shape == (1, 2)
    assert ddpg.act_with_noise(state, noise_param=(0.0, 1.0), mode="normal").shape == (1, 2)
    with pytest.raises(ValueError, match="Unknown noise type"):
        ddpg.act_with_noise(state, mode="some_unknown_noise")
    assert ddpg.act_discrete(state)[0].shape == (1, 1)
    ddpg.store_episode(_steps(3, t.zeros([1, 2])))
    for flags in itertools.product((True, False), repeat=3):
        pol, val = ddpg.update(update_value=flags[0], update_policy=flags[1], update_target=flags[2], concatenate_samples=True)
        assert np.isfinite(pol) and np.isfinite(val)
    ddpg.save(str(tmp_path), version=3)
    ddpg.load(str(tmp_path), version=3)
    assert all(t.equal(a, b) for a, b in zip(ddpg.actor.parameters(), ddpg.actor_target.parameters()))
    sched = DDPGPer(*nets(), t.optim.Adam, t.nn.MSELoss(reduction="sum"), replay_size=10,
                    lr_scheduler=t.optim.lr_scheduler.LambdaLR, lr_scheduler_args=((lambda s: 0.1,), (lambda s: 0.1,)))
    sched.update_lr_scheduler()
    assert len(sched.lr_schedulers) == 2 and len(sched.optimizers) == 2


def test_a2c_reference_cases(doubles, monkeypatch):  # noqa: F811
    import machin_b200.frame.algorithms.a2c as a2c_mod
    monkeypatch.setattr(a2c_mod.ops, "gae", lambda reward, terminal, cv, ncv, discount, lam, offsets=None, n_episodes=None,
                        max_len=0: {"returns": t.zeros_like(reward, dtype=t.float64), "gae": t.zeros_like(reward, dtype=t.float64)})

    class Policy(t.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = t.nn.Linear(S, A)

        def forward(self, state, action=None):
            dist = t.distributions.Categorical(logits=self.fc(state))
            act = action if action is not None else dist.sample().view(-1, 1)
            return act, dist.log_prob(act.flatten()).view(-1, 1), dist.entropy().view(-1, 1)

    for entropy_weight in (None, 1e-2):
        a2c = a2c_mod.A2C(Policy(), Critic(S), t.optim.Adam, t.nn.MSELoss(reduction="sum"), batch_size=2, replay_size=100,
                          entropy_weight=entropy_weight, actor_update_times=2, critic_update_times=2)
        act, logp, ent = a2c.act({"state": t.zeros([1, S])})
        assert act.shape == (1, 1) and logp.shape == (1, 1)
        for flags in itertools.product((True, False), repeat=2):
            a2c.store_episode(_steps(3, t.zeros([1, 1], dtype=t.int)))
            pol, val = a2c.update(update_value=flags[0], update_policy=flags[1], concatenate_samples=True)
            assert np.isfinite(pol) and np.isfinite(val) and a2c.replay_buffer.size() == 0


class _Policy(t.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = t.nn.Linear(S, A)

    def forward(self, state, action=None):
        dist = t.distributions.Categorical(logits=self.fc(state))
        act = action if action is not None else dist.sample().view(-1, 1)
        return act, dist.log_prob(act.flatten()).view(-1, 1), dist.entropy().view(-1, 1)


def test_impala_and_a3c_reference_cases_on_a_group_of_one(doubles):  # noqa: F811
    """test_impala.py / test_a3c.py case lists (act, store, update with every flag combination, set_sync / manual_sync),
    with the group of one that exists outside torch.distributed: model and gradient servers are real, collectives are
    identities."""
    from machin_b200.frame.algorithms import A3C, IMPALA
    from machin_b200.frame.helpers import grad_server_helper, model_server_helper
    from machin_b200.parallel import LocalGroup
    dev = t.device("cpu", 0)
    group = LocalGroup("impala", device="cpu")
    group.device = dev
    imp = IMPALA(_Policy(), Critic(S), t.optim.Adam, t.nn.MSELoss(reduction="sum"), group,
                 model_server_helper(1, device=dev), batch_size=2, replay_size=100)
    act, logp, *_ = imp.act({"state": t.zeros([1, S])})
    assert act.shape == (1, 1) and imp.is_distributed()
    with pytest.raises(ValueError, match="Episode must be non-empty"):
        imp.store_episode([])
    for n in (3, 5):
        ep = _steps(n, t.zeros([1, 1], dtype=t.int))
        for tr in ep:
            tr["action_log_prob"] = 0.1
        imp.store_episode(ep)
        assert ep[0]["episode_length"] == n and all(tr["episode_length"] == 0 for tr in ep[1:])
    for flags in itertools.product((True, False), repeat=2):
        pol, val = imp.update(update_value=flags[0], update_policy=flags[1])
        assert np.isfinite(pol) and np.isfinite(val)
    imp.set_sync(False)
    imp.manual_sync()
    assert imp.actor.pp_version >= 1                     # pushed after every update (impala.py:393-396)

    servers = grad_server_helper([_Policy, lambda: Critic(S)], learning_rate=[1e-3, 1e-3], device=dev)
    a3c = A3C(_Policy(), Critic(S), t.nn.MSELoss(reduction="sum"), servers, batch_size=2, replay_size=100,
              actor_update_times=1, critic_update_times=1)
    assert a3c.is_distributed() and a3c.optimizers == [] and a3c.lr_schedulers == []
    before = [p.detach().clone() for p in servers[1].model.parameters()]
    for flags in itertools.product((True, False), repeat=2):
        a3c.store_episode(_steps(3, t.zeros([1, 1], dtype=t.int)))
        pol, val = a3c.update(update_value=flags[0], update_policy=flags[1])
        assert np.isfinite(pol) and np.isfinite(val)
    a3c.set_sync(False)
    a3c.manual_sync()
    # (advantages are zero in a dry run, so only the critic sees gradients) the server's copy stepped and was pulled
    assert any(not t.equal(a, b) for a, b in zip(before, servers[1].model.parameters()))
    assert all(t.equal(a, b) for a, b in zip(a3c.critic.parameters(), servers[1].model.parameters()))
    assert all(t.equal(a, b) for a, b in zip(a3c.actor.parameters(), servers[0].model.parameters()))


def test_apex_reference_cases_on_a_group_of_one(doubles):  # noqa: F811
    """test_apex.py case list for DQNApex / DDPGApex (act with and without syncing, store, update -> weights pushed,
    manual_sync, the sampler-side serve_update) with real model servers on a group of one."""
    from machin_b200.frame.algorithms import DDPGApex, DQNApex
    from machin_b200.frame.helpers import model_server_helper
    from machin_b200.parallel import LocalGroup
    dev = t.device("cpu", 0)

    def group(name):
        g = LocalGroup(name, device="cpu")
        g.device = dev
        return g

    apex = DQNApex(QNet(S, A), QNet(S, A), t.optim.Adam, t.nn.MSELoss(reduction="sum"), group("dqn_apex"),
                   model_server_helper(1, device=dev), batch_size=2, replay_size=100)
    state = {"state": t.zeros([1, S])}
    apex.store_episode(_steps(3, t.zeros([1, 1], dtype=t.int)))
    for flags in itertools.product((True, False), repeat=2):
        assert np.isfinite(apex.update(update_value=flags[0], update_target=flags[1]))
    version = apex.qnet.pp_version
    assert version >= 4                                          # published after every update (apex.py:141-150)
    assert apex.act_discrete(state).shape == (1, 1) and apex.act_discrete(state, True).shape == (1, 1)
    assert apex.act_discrete_with_noise(state).shape == (1, 1)
    apex.set_sync(False)
    assert apex.act_discrete(state).shape == (1, 1) and not apex.is_syncing
    apex.manual_sync()
    apex.serve_update()                                          # what a member that only holds a shard calls instead
    assert DQNApex.is_distributed()

    ddpg = DDPGApex(Actor(S, 2), Actor(S, 2), Critic(S, 2), Critic(S, 2), t.optim.Adam, t.nn.MSELoss(reduction="sum"),
                    group("ddpg_apex"), model_server_helper(1, device=dev), batch_size=2, replay_size=100)
    ddpg.store_episode(_steps(3, t.zeros([1, 2])))
    for flags in itertools.product((True, False), repeat=3):
        pol, val = ddpg.update(update_value=flags[0], update_policy=flags[1], update_target=flags[2])
        assert np.isfinite(pol) and np.isfinite(val)
    assert ddpg.actor.pp_version >= 8                            # the ACTOR is what gets published (apex.py:372-381)
    assert ddpg.act(state).shape == (1, 2) and ddpg.act(state, use_target=True).shape == (1, 2)
    ddpg.set_sync(False)
    ddpg.manual_sync()
    ddpg.serve_update()
    assert ddpg.act_with_noise(state, noise_param=(0.0, 0.1)).shape == (1, 2)
