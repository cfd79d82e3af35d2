"""machin_b200.auto on the CPU: the cases of the reference's test/auto/test_launcher.py, test/auto/test_dataset.py and
test/auto/env/test_openai_gym.py that do not need pytorch-lightning or gym, over the host-memory buffer doubles of
tests/test_config_cpu.py (the real buffers need a GPU: tests/test_gpu_zz_auto.py runs the same loop on the device)."""
import json
import os
import sys

import numpy as np
import pytest
import torch as t

from test_config_cpu import Actor, Critic, QNet, doubles  # noqa: F401  (fixture + classes resolved by name)

from machin_b200.auto import config as auto_config
from machin_b200.auto.dataset import DatasetResult, RLDataset, determine_precision
from machin_b200.auto.envs import classic_control, openai_gym
from machin_b200.auto.launcher import DistributedLauncher, EarlyStopping, Launcher, ModelCheckpoint, Trainer

S, A = 4, 2
seen = {}


def note(_launcher, key, value):
    seen[key] = value


class OneTransition(RLDataset):
    early_stopping_monitor = "total_reward"

    def __init__(self, rewards=None):
        super().__init__()
        self.rewards, self.n = rewards, 0

    def __next__(self):
        result = DatasetResult()
        for i in range(3):
            result.add_observation({"state": {"state": t.ones(1, S)}, "action": {"action": t.ones(1, 1, dtype=t.long)},
                                    "next_state": {"state": t.ones(1, S)}, "terminal": i == 2, "reward": 1.0})
        result.add_log({"test_log": (1, note)})
        result.add_log({"total_reward": self.rewards[min(self.n, len(self.rewards) - 1)] if self.rewards else 3.0})
        self.n += 1
        return result


def dqn_config(frame="DQNPer"):
    from machin_b200.frame import algorithms
    cfg = getattr(algorithms, frame).generate_config({})
    cfg["frame_config"].update(models=["QNet", "QNet"], model_kwargs=[{"state_dim": S, "action_num": A}] * 2, batch_size=4,
                               replay_size=64)
    cfg.update(train_env_config={}, test_env_config={})
    return cfg


# ------------------------------------------------------------------------------------------------ dataset.py
def test_determine_precision():
    assert determine_precision([QNet(10, 2)]) == t.float32
    mixed = t.nn.Sequential(t.nn.Linear(4, 4), t.nn.Linear(4, 2).double())
    with pytest.raises(RuntimeError, match="Multiple data types of parameters"):
        determine_precision([mixed])
    with pytest.raises(RuntimeError, match="Multiple data types of parameters"):
        determine_precision([QNet(10, 2), QNet(10, 2).double()])


def test_dataset_result():
    result = DatasetResult()
    result.add_observation({})
    result.add_log({})
    assert len(result.observations) == 1 and len(result) == 1 and len(result.logs) == 1
    assert DatasetResult().observations is not DatasetResult().observations
    with pytest.raises(StopIteration):
        next(iter(RLDataset()))


# ------------------------------------------------------------------------------------------------ config.py
def test_available_and_generate():
    algos, environments = auto_config.get_available_algorithms(), auto_config.get_available_environments()
    assert {"DQNPer", "DQNApex", "DDPGApex", "IMPALA", "A3C"} <= set(algos) and "TorchFramework" not in algos
    assert environments == ["classic_control", "openai_gym"]
    cfg = auto_config.generate_env_config("openai_gym", {})
    assert cfg["env"] == "openai_gym" and cfg["train_env_config"]["env_name"] == "CartPole-v1"
    cfg = auto_config.generate_algorithm_config("DQNApex", cfg)
    assert cfg["frame"] == "DQNApex" and cfg["gpus"] == [0, 0, 0] and cfg["num_nodes"] == 1 and cfg["batch_num"]["sampler"] == 10
    assert auto_config.generate_algorithm_config("DQNPer")["gpus"] == [0]
    cfg = auto_config.generate_training_config(config=cfg)
    assert (cfg["root_dir"], cfg["episode_per_epoch"], cfg["max_episodes"], cfg["early_stopping_patience"]) == ("trial", 10, 10000, 3)
    auto_config.assert_training_config_complete(cfg)
    auto_config.assert_env_config_complete(cfg)
    auto_config.assert_algorithm_config_complete(cfg)
    assert auto_config.is_algorithm_distributed(cfg) and not auto_config.is_algorithm_distributed(dqn_config())
    kept = auto_config.generate_env_config("openai_gym", {"env": "mine"})
    assert kept["env"] == "mine" and "train_env_config" in kept          # fill_default never overwrites
    with pytest.raises(ValueError, match="Invalid algorithm: Nope"):
        auto_config.generate_algorithm_config("Nope")
    with pytest.raises(ValueError, match="Invalid environment: nope"):
        auto_config.generate_env_config("nope")
    with pytest.raises(ValueError, match="Invalid algorithm"):
        auto_config.init_algorithm_from_config({"frame": "TorchFramework", "frame_config": {}})
    with pytest.raises(AssertionError, match='Missing key "frame"'):
        auto_config.is_algorithm_distributed({})
    with pytest.raises(AssertionError, match='Missing key "max_episodes"'):
        auto_config.launch({"root_dir": "x", "episode_per_epoch": 1})


def test_cli(tmp_path, capsys):
    from machin_b200.auto.__main__ import main
    main(["list", "--algo"])
    assert "DQNApex" in capsys.readouterr().out
    main(["list", "--env"])
    assert "openai_gym" in capsys.readouterr().out
    main(["list"])
    assert "You can list --algo or --env." in capsys.readouterr().out
    out = str(tmp_path / "c.json")
    main(["generate", "--algo", "Nope", "--env", "openai_gym", "--output", out])
    assert "not a valid algorithm" in capsys.readouterr().out and not os.path.exists(out)
    main(["generate", "--algo", "IMPALA", "--env", "openai_gym", "--output", out, "--print"])
    assert "Config saved to" in capsys.readouterr().out
    cfg = json.load(open(out))
    assert cfg["frame"] == "IMPALA" and cfg["frame_config"]["gradient_max"] == float("inf") and cfg["max_episodes"] == 10000


# ------------------------------------------------------------------------------------------------ launcher.py
def test_launcher_steps(doubles):
    launcher = Launcher(dqn_config(), lambda frame, env_config: OneTransition(), model_device="cpu")
    launcher.on_train_start()
    launcher.on_test_start()
    batch = next(iter(launcher.train_dataloader()))
    assert isinstance(batch[0], DatasetResult) and len(batch) == 1 and "state" in batch[0].observations[0]
    assert isinstance(next(iter(launcher.test_dataloader()))[0], DatasetResult)
    optimizers, schedulers = launcher.configure_optimizers()
    assert len(optimizers) == 1 and isinstance(optimizers[0], t.optim.Optimizer) and len(schedulers) == 0
    updates = []
    real_update = launcher.frame.update
    launcher.frame.update = lambda *a, **k: updates.append(real_update(*a, **k))
    seen.clear()
    launcher.training_step(batch, 0)
    assert seen == {"test_log": 1} and launcher.logged["total_reward"] == 3.0
    assert len(updates) == 3 and all(np.isfinite(x) for x in updates)      # one update per stored transition
    assert launcher.frame.replay_buffer.size() == 3
    seen.clear()
    launcher.test_step(batch, 0)
    assert seen == {"test_log": 1} and len(updates) == 3
    with pytest.raises(AssertionError):
        DistributedLauncher(dqn_config(), lambda f, c: OneTransition())
    with pytest.raises(AssertionError):
        Launcher(dqn_config("DQNApex"), lambda f, c: OneTransition(), model_device="cpu")


def test_trainer_fit_limits_and_callbacks(doubles, tmp_path):
    # the reference's test_real: limit_train_batches=1, max_steps=1
    launcher = Launcher(dqn_config(), lambda frame, env_config: OneTransition(), model_device="cpu")
    trainer = Trainer(limit_train_batches=1, max_steps=1).fit(launcher)
    assert (trainer.global_step, trainer.current_epoch) == (1, 1) and launcher.frame.replay_buffer.size() == 3
    with pytest.raises(ValueError, match="max_steps or max_epochs"):
        Trainer()
    with pytest.raises(ValueError, match="limit_train_batches with max_epochs"):
        Trainer(max_epochs=2)
    # early stopping on the monitored key: best at epoch 1, three epochs without improvement end the run at epoch 4
    rewards = [1.0, 5.0, 4.0, 5.0, 3.0, 9.0, 9.0]
    launcher = Launcher(dqn_config(), lambda frame, env_config: OneTransition(rewards), model_device="cpu")
    steps = []

    class Steps:
        def on_step_end(self, trainer, module, metrics):
            steps.append(metrics)

    model_dir = str(tmp_path / "model")
    trainer = Trainer(limit_train_batches=1, max_steps=100, callbacks=[
        ModelCheckpoint(model_dir, "total_reward"), EarlyStopping("total_reward", patience=3), Steps()]).fit(launcher)
    assert trainer.should_stop and trainer.current_epoch == 5 and trainer.global_step == 5
    assert [m["total_reward"] for m in steps] == rewards[:5]
    assert sorted(os.listdir(model_dir)) == ["qnet_target_1.pt"]           # only the best epoch's restorable models remain
    launcher.frame.load(model_dir)
    # max_epochs, several batches per epoch
    launcher = Launcher(dqn_config(), lambda frame, env_config: OneTransition(), model_device="cpu")
    trainer = Trainer(limit_train_batches=3, max_epochs=2).fit(launcher)
    assert (trainer.global_step, trainer.current_epoch) == (6, 2)
    assert Trainer(max_steps=1, limit_test_batches=2).test(launcher)["total_reward"] == 3.0


def test_distributed_launcher_world_of_one(doubles):
    cfg = dqn_config("DQNApex")
    cfg["batch_num"] = {"learner": 2, "sampler": 5}
    launcher = DistributedLauncher(cfg, lambda frame, env_config: OneTransition(), model_device="cpu")
    assert launcher.frame is None and launcher.configure_optimizers() is None
    loader = launcher.train_dataloader()                                   # builds the frame (collective in a real job)
    assert launcher.frame.role == "learner" and launcher._batch_num == 2
    batch = next(iter(loader))
    assert len(batch) == 2
    updates = []
    real_update = launcher.frame.update
    launcher.frame.update = lambda *a, **k: updates.append(real_update(*a, **k))
    seen.clear()
    launcher.training_step(batch, 0)
    assert updates == [] and launcher.frame.replay_buffer.size() == 6      # the first batch only fills the buffer
    assert seen == {"test_log": 1} and launcher.logged["total_reward"] == 3.0
    launcher.training_step(batch, 1)
    assert len(updates) == 6                                               # one per stored transition of the batch
    launcher.current_epoch = 1
    launcher.training_step(batch, 0)
    assert len(updates) == 12
    trainer = Trainer(limit_train_batches=2, max_steps=3).fit(
        DistributedLauncher(cfg, lambda frame, env_config: OneTransition(), model_device="cpu"))
    assert trainer.global_step == 3 and trainer.callback_metrics["total_reward"] == 3.0


# ------------------------------------------------------------------------------------------------ envs
def test_classic_control_tasks():
    env = classic_control.make("CartPole-v1", seed=3)
    first = env.reset()
    assert first.shape == (4,) and first.dtype == np.float32 and np.all(np.abs(first) <= 0.05)
    steps, done = 0, False
    while not done:                                    # always pushing right topples the pole within a few dozen steps
        obs, reward, done, info = env.step(1)
        steps += 1
        assert reward == 1.0 and info == {}
    assert 5 < steps < 60 and (abs(obs[2]) > env.ANGLE_LIMIT or abs(obs[0]) > env.TRACK_LIMIT)
    again = classic_control.make("CartPole-v1", seed=3)
    assert np.array_equal(again.reset(), first)        # seeded
    # one Euler step against the equations of motion evaluated by hand at a known state
    env.state, env.steps = np.array([0.0, 0.0, 0.1, 0.0]), 0
    obs, *_ = env.step(1)
    sin, cos = np.sin(0.1), np.cos(0.1)
    tmp = 10.0 / 1.1
    alpha = (9.8 * sin - cos * tmp) / (0.5 * (4.0 / 3.0 - 0.1 * cos * cos / 1.1))
    acc = tmp - 0.05 * alpha * cos / 1.1
    np.testing.assert_allclose(obs, [0.0, 0.02 * acc, 0.1, 0.02 * alpha], rtol=1e-6)
    pend = classic_control.make("Pendulum-v0", seed=1)
    obs = pend.reset()
    assert obs.shape == (3,) and abs(obs[0] ** 2 + obs[1] ** 2 - 1.0) < 1e-6
    total, done, n = 0.0, False, 0
    while not done:
        obs, reward, done, _ = pend.step(np.array([5.0]))                  # clipped to the torque limit
        total, n = total + reward, n + 1
        assert reward <= 0.0 and abs(obs[2]) <= 8.0
    assert n == 200 and total < 0
    pend.th, pend.w, pend.steps = 0.0, 0.0, 0
    _, reward, *_ = pend.step([0.0])
    assert reward == 0.0 and pend.w == 0.0                                 # upright, at rest, no torque: no cost, no motion
    with pytest.raises(ValueError, match="Unknown classic control task"):
        classic_control.make("Nope-v0")


def test_gym_datasets_act_and_collect(doubles):
    from machin_b200.frame.algorithms import DDPGPer, DQNPer
    dqn = Launcher(dqn_config(), lambda f, c: OneTransition(), model_device="cpu").frame
    dataset = openai_gym.gym_env_dataset_creator(dqn, {"env_name": "CartPole-v1", "render_every_episode": 100, "act_kwargs": {}})
    assert isinstance(dataset, openai_gym.RLGymDiscActDataset) and dataset.early_stopping_monitor == "total_reward"
    episode = next(dataset)
    assert len(episode) >= 5 and episode.logs[-1] == {"total_reward": float(len(episode))}
    first, last = episode.observations[0], episode.observations[-1]
    assert first["state"]["state"].shape == (1, 4) and first["state"]["state"].dtype == t.float32
    assert first["action"]["action"].shape == (1, 1) and first["action"]["action"].dtype == t.int64
    assert last["terminal"] is True and not any(o["terminal"] for o in episode.observations[:-1])
    assert t.equal(episode.observations[1]["state"]["state"], first["next_state"]["state"])
    dqn.store_episode(episode.observations)
    assert dataset.counter == 1
    # continuous actions: the DDPG family on the pendulum
    cfg = DDPGPer.generate_config({})
    cfg["frame_config"].update(model_kwargs=[{"state_dim": 3, "action_dim": 1}] * 4, batch_size=4, replay_size=512)
    cfg.update(train_env_config={}, test_env_config={})
    ddpg = Launcher(cfg, lambda f, c: OneTransition(), model_device="cpu").frame
    dataset = openai_gym.dataset_for(ddpg, classic_control.make("Pendulum-v0"), {"act_kwargs": {"noise_param": (-0.1, 0.1)}})
    assert isinstance(dataset, openai_gym.RLGymContActDataset)
    episode = next(dataset)
    assert len(episode) == 200 and episode.observations[0]["action"]["action"].shape == (1, 1)
    assert episode.observations[0]["action"]["action"].dtype == t.float32 and episode.logs[-1]["total_reward"] < 0
    ddpg.store_episode(episode.observations)
    # a discrete dataset refuses a continuous environment and vice versa; an unknown name without gym is an ImportError
    with pytest.raises(AssertionError):
        openai_gym.RLGymDiscActDataset(dqn, classic_control.make("Pendulum-v0"))
    with pytest.raises(AssertionError):
        openai_gym.RLGymContActDataset(ddpg, classic_control.make("CartPole-v1"))
    if "gym" not in sys.modules:
        try:
            import gym  # noqa: F401
        except ImportError:
            with pytest.raises(ImportError, match="gym is not installed"):
                openai_gym.gym_env_dataset_creator(dqn, {"env_name": "Breakout-v4"})


def test_launch_from_config(doubles, tmp_path):
    cfg = auto_config.generate_env_config("classic_control", {})
    cfg = auto_config.generate_algorithm_config("DQNPer", cfg)
    cfg = auto_config.generate_training_config(root_dir=str(tmp_path / "trial"), episode_per_epoch=2, max_episodes=3, config=cfg)
    cfg["frame_config"].update(models=["QNet", "QNet"], model_kwargs=[{"state_dim": 4, "action_num": 2}] * 2, batch_size=8,
                               replay_size=2048)
    cfg["model_device"] = "cpu"
    marks = []

    class Mark:
        def on_epoch_end(self, trainer, module):
            marks.append((trainer.current_epoch, trainer.global_step))

    trainer = auto_config.launch(cfg, [Mark()])
    assert trainer.global_step == 3 and marks == [(0, 2), (1, 3)]
    lines = [json.loads(x) for x in open(tmp_path / "trial" / "log" / "train_log" / "metrics.jsonl")]
    assert len(lines) == 3 and all(x["total_reward"] >= 5 for x in lines) and [x["step"] for x in lines] == [1, 2, 3]
    assert any(name.startswith("qnet_target_") for name in os.listdir(tmp_path / "trial" / "model"))


# ------------------------------------------------------------------------------------------------ two ranks (gloo)
class _CountingFrame:
    """Stands for a sharded framework: ``update`` is a collective (every rank must enter it equally often)."""

    def __init__(self, role):
        self.role, self.stored, self.updates = role, 0, 0

    def store_episode(self, episode):
        self.stored += len(episode)

    def update(self):
        import torch.distributed as dist
        cell = t.ones(1)
        dist.all_reduce(cell)                          # hangs (and the test times out) if the ranks disagree on the count
        self.updates += 1
        return float(cell.item())


def _two_rank_launcher(rank, world):
    from machin_b200.auto.launcher import DistributedLauncher, Trainer
    from test_auto_cpu import OneTransition, _CountingFrame, dqn_config
    cfg = dqn_config("DQNApex")
    rewards = [[2.0, 4.0, 6.0, 8.0], [10.0] * 8][rank]
    launcher = DistributedLauncher(cfg, lambda frame, env_config: OneTransition(rewards), model_device="cpu")
    launcher.frame = _CountingFrame("learner" if rank == 0 else "sampler")     # (init_frame keeps an existing frame)
    launcher._batch_num = 1 if rank == 0 else 2                                # batch_num by role
    trainer = Trainer(limit_train_batches=2, max_steps=3).fit(launcher)
    # rank 0 stores one 3-transition episode per step, rank 1 two; the learner's count (3) is run by both, not at step 0
    return (launcher.frame.stored, launcher.frame.updates, trainer.global_step, trainer.callback_metrics["total_reward"],
            launcher.logged["total_reward"])


def test_distributed_launcher_two_ranks():
    from test_distributed_cpu import run_multi
    out = run_multi(_two_rank_launcher, 2)
    assert out[0][:3] == (9, 6, 3) and out[1][:3] == (18, 6, 3)
    # the monitored key is the mean over the ranks of each rank's first episode of the last step:
    # rank 0 drew rewards[2] = 6, rank 1 drew its 5th episode = 10
    assert out[0][3] == out[1][3] == 8.0 and out[0][4] == out[1][4] == 8.0


# ------------------------------------------------------------------------------------------------ policy-gradient families
class Policy(t.nn.Module):
    """Two-action categorical policy with the (action, log-probability, entropy) output convention."""

    def __init__(self):
        super().__init__()
        self.fc = t.nn.Linear(4, 2)

    def forward(self, state, action=None):
        law = t.distributions.Categorical(logits=self.fc(state))
        act = action if action is not None else law.sample().view(-1, 1)
        return act, law.log_prob(act.flatten()).view(-1, 1), law.entropy().view(-1, 1)


class ShortCartPole(classic_control.CartPole):
    MAX_STEPS = 12                                     # bounded episodes, so that the per-transition updates stay few


def short_cart_pole(frame, env_config):
    return openai_gym.dataset_for(frame, ShortCartPole(seed=1), env_config)


def policy_config(algo, **frame):
    cfg = auto_config.generate_algorithm_config(algo, auto_config.generate_env_config("classic_control", {}))
    cfg["frame_config"].update(models=[Policy, "Critic"], model_kwargs=[{}, {"state_dim": 4}], **frame)
    return cfg


def test_policy_gradient_frameworks_through_the_launchers(doubles):
    cfg = policy_config("A2C", batch_size=8, replay_size=256, actor_update_times=1, critic_update_times=1)
    launcher = Launcher(cfg, short_cart_pole, model_device="cpu")
    episode = next(iter(launcher.train_dataloader()))[0]
    assert 1 <= len(episode) <= 12 and "action_log_prob" not in episode.observations[0]
    assert episode.observations[0]["action"]["action"].dtype == t.int64
    Trainer(limit_train_batches=1, max_steps=1).fit(launcher)
    assert np.isfinite(launcher.logged["total_reward"])
    cfg = policy_config("IMPALA", batch_size=2, replay_size=64)
    cfg["batch_num"] = {"learner": 2, "sampler": 2}
    launcher = DistributedLauncher(cfg, short_cart_pole, model_device="cpu")
    batch = next(iter(launcher.train_dataloader()))
    assert len(batch) == 2 and all(isinstance(o["action_log_prob"], float) and o["action_log_prob"] <= 0.0
                                   for o in batch[0].observations)
    trainer = Trainer(limit_train_batches=2, max_steps=2).fit(launcher)
    assert trainer.global_step == 2 and np.isfinite(launcher.logged["total_reward"])
    cfg = policy_config("A3C", batch_size=8, replay_size=256, actor_update_times=1, critic_update_times=1)
    launcher = DistributedLauncher(cfg, short_cart_pole, model_device="cpu")
    Trainer(limit_train_batches=1, max_steps=2).fit(launcher)
    assert np.isfinite(launcher.logged["total_reward"]) and launcher._batch_num == 1


# ------------------------------------------------------------------------------------------------ the surface as data
def test_auto_surface_matches_reference():
    """tests/golden/ref_auto_surface.json: public functions / classes / methods of machin/auto with their parameter names
    (oracle/make_golden.py::gen_auto_surface, read from the reference's source).  All exist here with the same parameter
    names, apart from the image / video log handlers of dataset.py:31-73 (media utilities, out of scope)."""
    import importlib
    import inspect
    from conftest import GOLDEN
    ref = json.load(open(os.path.join(GOLDEN, "ref_auto_surface.json")))
    absent = {("dataset", "get_loggers_as_list"), ("dataset", "log_image"), ("dataset", "log_video")}
    problems = []

    def check(where, obj, params):
        have = inspect.signature(obj).parameters
        lacking = [p for p in params if p not in have]
        if lacking:
            problems.append(f"{where} lacks parameters {lacking}")

    for module_name, members in ref.items():
        module = importlib.import_module("machin_b200.auto." + module_name)
        for name, spec in members.items():
            if (module_name, name) in absent:
                continue
            if not hasattr(module, name):
                problems.append(f"{module_name}.{name} missing")
            elif isinstance(spec, dict):
                for method, params in spec.items():
                    if not hasattr(getattr(module, name), method):
                        problems.append(f"{module_name}.{name}.{method} missing")
                    else:
                        check(f"{module_name}.{name}.{method}", getattr(getattr(module, name), method), params)
            else:
                check(f"{module_name}.{name}", getattr(module, name), spec)
    assert not problems, problems


# ------------------------------------------------------------------------------------------------ the loop against the reference's
def test_launcher_steps_match_reference_trace():
    """tests/golden/ref_auto_trace.json: what the UNMODIFIED launchers (machin/auto/launcher.py, run in the build container
    over a recording LightningModule stub: oracle/make_golden.py::gen_auto_trace) do with the same batches.  The local
    launcher must make the same calls in the same order.  The distributed one must make the same calls per step -- stores,
    one update per stored transition except on the first batch of epoch 0, logs of the first sample only -- in its own
    order (all stores, then the agreed number of updates: DESIGN.md 6) and without a barrier in a world of one."""
    from collections import Counter
    from conftest import GOLDEN
    from machin_b200.auto.launcher import _EpisodeModule
    ref = json.load(open(os.path.join(GOLDEN, "ref_auto_trace.json")))

    class Frame:
        role = "learner"

        def __init__(self):
            self.events = []

        def store_episode(self, episode):
            self.events.append(["store", len(episode)])

        def update(self):
            self.events.append(["update"])

    def handler(module, key, value):
        module.frame.events.append(["handler", key, value])

    def batch():
        first, second = DatasetResult(), DatasetResult()
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
        _EpisodeModule.__init__(module, {}, None)
        module.frame, module._graph_pending, module.calls = Frame(), False, []
        plain = module.log
        module.log = lambda key, value, prog_bar=False, sync_dist=False: (
            module.calls.append([key, value, {"prog_bar": prog_bar, "sync_dist": sync_dist}]), plain(key, value, prog_bar, sync_dist))
        return module

    def same_logs(mine, theirs):
        return [[k, v, {"sync_dist": False, **kw}] for k, v, kw in theirs] == mine

    local = fresh(Launcher)
    local.training_step(batch(), 0)
    assert local.frame.events == ref["launcher_training_step"]["events"]
    assert same_logs(local.calls, ref["launcher_training_step"]["logged"])
    local = fresh(Launcher)
    local.test_step(batch(), 0)
    assert local.frame.events == ref["launcher_test_step"]["events"] and same_logs(local.calls, ref["launcher_test_step"]["logged"])
    for name, epoch, batch_idx in (("first_batch", 0, 0), ("later_batch", 0, 1), ("later_epoch", 1, 0)):
        theirs = ref["distributed_training_step_" + name]
        d = fresh(DistributedLauncher)
        d._train_esm, d._test_esm, d.current_epoch = "total_reward", None, epoch
        d.training_step(list(reversed(batch())), batch_idx)
        count = lambda events: Counter(tuple(e) for e in events if e != ["barrier"])
        assert count(d.frame.events) == count(theirs["events"]), name
        stores = [e for e in d.frame.events if e[0] == "store"]
        assert d.frame.events[:len(stores)] == stores == [e for e in theirs["events"] if e[0] == "store"]
        assert same_logs(d.calls, theirs["logged"]), name
    d = fresh(DistributedLauncher)
    d._train_esm, d._test_esm = "total_reward", None
    d.test_step(batch(), 0)
    assert d.frame.events == ref["distributed_test_step"]["events"] and same_logs(d.calls, ref["distributed_test_step"]["logged"])


class AutoQNet(t.nn.Module):
    """The Q-net of tests/golden/ref_auto_dataset.npz (oracle/make_golden.py::AutoQNet: same layers in the same order)."""

    def __init__(self, state_dim, action_num):
        super().__init__()
        self.fc1, self.fc2 = t.nn.Linear(state_dim, 16), t.nn.Linear(16, action_num)

    def forward(self, state):
        return self.fc2(t.relu(self.fc1(state)))


def test_discrete_dataset_collects_the_reference_episodes(doubles):
    """tests/golden/ref_auto_dataset.npz: two episodes collected by the UNMODIFIED RLGymDiscActDataset with the reference's
    DQN acting (oracle/make_golden.py::gen_auto_dataset).  Same seeds, same net, same task: the dataset here must hand the
    launcher the very same transitions -- states bit for bit, the epsilon-greedy actions (same random stream), rewards,
    terminal flags, total reward -- and leave epsilon where the reference leaves it."""
    from conftest import GOLDEN
    from machin_b200.frame.algorithms import DQN
    ref = np.load(os.path.join(GOLDEN, "ref_auto_dataset.npz"))
    cfg = DQN.generate_config({})
    cfg["frame_config"].update(models=["AutoQNet", "AutoQNet"], model_kwargs=[{"state_dim": 4, "action_num": 2}] * 2)
    t.manual_seed(11)
    dqn = auto_config.init_algorithm_from_config(cfg, model_device="cpu")
    dqn.epsilon = 0.5
    dataset = openai_gym.RLGymDiscActDataset(dqn, classic_control.CartPole(seed=5), render_every_episode=10 ** 9)
    t.manual_seed(12)
    for e in range(2):
        result = next(dataset)
        obs = result.observations
        assert len(obs) == len(ref[f"ep{e}_reward"])
        assert np.array_equal(np.concatenate([o["state"]["state"].numpy() for o in obs]), ref[f"ep{e}_state"])
        assert np.array_equal(np.concatenate([o["next_state"]["state"].numpy() for o in obs]), ref[f"ep{e}_next_state"])
        assert np.array_equal(np.concatenate([o["action"]["action"].numpy() for o in obs]), ref[f"ep{e}_action"])
        assert [o["reward"] for o in obs] == list(ref[f"ep{e}_reward"]) and all(isinstance(o["reward"], float) for o in obs)
        assert [o["terminal"] for o in obs] == list(ref[f"ep{e}_terminal"]) and all(isinstance(o["terminal"], bool) for o in obs)
        assert result.logs[-1] == {"total_reward": float(ref[f"ep{e}_total_reward"])}
    assert dqn.epsilon == float(ref["epsilon_after"])


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


@pytest.mark.parametrize("tag", ["ddpg", "a2c"])
def test_discrete_dataset_other_branches_match_the_reference(doubles, tag):
    """The DDPG family (stores the softened probability row it sampled from, openai_gym.py:121-122) and the
    policy-gradient family (``act(...)[0]``) of the discrete dataset: one episode each against the unmodified reference's
    (tests/golden/ref_auto_dataset.npz) -- same states bit for bit, same sampled actions / stored rows."""
    from conftest import GOLDEN
    from machin_b200.frame.algorithms import A2C, DDPGPer
    ref = np.load(os.path.join(GOLDEN, "ref_auto_dataset.npz"))
    cls, models, kwargs = {
        "ddpg": (DDPGPer, ["AutoSoftActor", "AutoSoftActor", "AutoCritic", "AutoCritic"], [{"state_dim": 4, "action_dim": 2}] * 4),
        "a2c": (A2C, ["AutoPolicy", "AutoCritic"], [{"state_dim": 4, "action_num": 2}, {"state_dim": 4}])}[tag]
    cfg = cls.generate_config({})
    cfg["frame_config"].update(models=models, model_kwargs=kwargs)
    t.manual_seed(21)
    frame = auto_config.init_algorithm_from_config(cfg, model_device="cpu")
    dataset = openai_gym.RLGymDiscActDataset(frame, classic_control.CartPole(seed=6), render_every_episode=10 ** 9)
    t.manual_seed(22)
    obs = next(dataset).observations
    assert np.array_equal(np.concatenate([o["state"]["state"].numpy() for o in obs]), ref[f"{tag}_state"])
    action = np.concatenate([o["action"]["action"].numpy() for o in obs])
    assert action.dtype == ref[f"{tag}_action"].dtype and np.array_equal(action, ref[f"{tag}_action"])
    assert [o["terminal"] for o in obs] == list(ref[f"{tag}_terminal"])
