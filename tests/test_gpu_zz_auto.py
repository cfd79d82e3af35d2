"""GPU: machin_b200.auto drives the device-resident learners -- the launcher loop of tests/test_auto_cpu.py over real
buffers and kernels (reference cases: test/auto/test_launcher.py::test_real, test/auto/env/test_openai_gym.py)."""
import json
import os

import numpy as np
import pytest
import torch as t

from test_config_cpu import Actor, Critic, QNet  # noqa: F401  (resolved by name from this module's globals)
from test_auto_cpu import Policy, short_cart_pole  # noqa: F401

from machin_b200.auto import config as auto_config
from machin_b200.auto.envs import classic_control
from machin_b200.auto.launcher import DistributedLauncher, Launcher, Trainer

pytestmark = [pytest.mark.gpu]


def _config(algo, env_name, root, **frame):
    cfg = auto_config.generate_env_config("classic_control", {})
    for part in ("train_env_config", "test_env_config"):
        cfg[part]["env_name"] = env_name
    cfg = auto_config.generate_algorithm_config(algo, cfg)
    cfg = auto_config.generate_training_config(root_dir=str(root), episode_per_epoch=2, max_episodes=4, config=cfg)
    cfg["frame_config"].update(frame)
    return cfg


@pytest.mark.parametrize("graph", [False, True])
def test_launch_dqn_per_on_cart_pole(tmp_path, graph):
    cfg = _config("DQNPer", "CartPole-v1", tmp_path / "trial", models=["QNet", "QNet"],
                  model_kwargs=[{"state_dim": 4, "action_num": 2}] * 2, batch_size=16, replay_size=4096)
    cfg["cuda_graph"] = graph
    trainer = auto_config.launch(cfg)
    assert trainer.global_step == 4 and trainer.current_epoch == 2
    lines = [json.loads(x) for x in open(tmp_path / "trial" / "log" / "train_log" / "metrics.jsonl")]
    assert [x["step"] for x in lines] == [1, 2, 3, 4] and all(x["total_reward"] >= 5 for x in lines)
    assert any(name.startswith("qnet_target_") for name in os.listdir(tmp_path / "trial" / "model"))


def test_launcher_updates_on_device(tmp_path):
    cfg = _config("DQNPer", "CartPole-v1", tmp_path, models=["QNet", "QNet"],
                  model_kwargs=[{"state_dim": 4, "action_num": 2}] * 2, batch_size=16, replay_size=4096)
    launcher = Launcher(cfg, classic_control._dataset)
    assert next(launcher.frame.qnet.parameters()).is_cuda
    before = [p.detach().clone() for p in launcher.frame.qnet.parameters()]
    losses, real_update = [], launcher.frame.update
    launcher.frame.update = lambda *a, **k: losses.append(real_update(*a, **k))
    Trainer(limit_train_batches=1, max_steps=2).fit(launcher)
    stored = launcher.frame.replay_buffer.size()
    assert stored >= 10 and len(losses) == stored and all(np.isfinite(x) for x in losses)
    assert any(not t.equal(a, b) for a, b in zip(before, launcher.frame.qnet.parameters()))
    launcher.frame.replay_buffer.wt_tree.check()
    assert Trainer(max_steps=1).test(launcher)["total_reward"] >= 5


def test_launch_ddpg_per_on_pendulum(tmp_path):
    cfg = _config("DDPGPer", "Pendulum-v0", tmp_path, model_kwargs=[{"state_dim": 3, "action_dim": 1}] * 4, batch_size=16,
                  replay_size=4096)
    cfg["train_env_config"]["act_kwargs"] = {"noise_param": (-0.1, 0.1)}
    cfg["max_episodes"] = 2
    launcher = Launcher(cfg, classic_control._dataset)
    Trainer(limit_train_batches=1, max_steps=2).fit(launcher)
    assert launcher.frame.replay_buffer.size() == 400 and launcher.logged["total_reward"] < 0
    # the DDPG family on a discrete task stores the probability row as the action
    cfg = _config("DDPGPer", "CartPole-v1", tmp_path, models=["SoftActor", "SoftActor", "Critic", "Critic"],
                  model_kwargs=[{}, {}, {"state_dim": 4, "action_dim": 2}, {"state_dim": 4, "action_dim": 2}], batch_size=8,
                  replay_size=1024)
    launcher = Launcher(cfg, classic_control._dataset)
    episode = next(iter(launcher.train_dataloader()))[0]
    row = episode.observations[0]["action"]["action"]
    assert row.shape == (1, 2) and abs(float(row.sum()) - 1.0) < 1e-5
    launcher.training_step([episode], 0)


class SoftActor(t.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = t.nn.Linear(4, 2)

    def forward(self, state):
        return t.softmax(self.fc(state), dim=1)


def test_distributed_launchers_world_of_one(tmp_path):
    cfg = _config("DQNApex", "CartPole-v1", tmp_path, models=["QNet", "QNet"],
                  model_kwargs=[{"state_dim": 4, "action_num": 2}] * 2, batch_size=16, replay_size=4096)
    cfg["batch_num"] = {"learner": 2, "sampler": 2}
    launcher = DistributedLauncher(cfg, classic_control._dataset)
    trainer = Trainer(limit_train_batches=2, max_steps=3).fit(launcher)
    assert trainer.global_step == 3 and launcher.frame.role == "learner" and launcher.frame.replay_buffer.size() >= 30
    launcher.frame.replay_buffer.wt_tree.check()
    # IMPALA: transitions carry the behaviour policy's log-probability
    cfg = _config("IMPALA", "CartPole-v1", tmp_path, models=[Policy, "Critic"], model_kwargs=[{}, {"state_dim": 4}],
                  batch_size=2, replay_size=64)
    cfg["batch_num"] = {"learner": 2, "sampler": 2}
    launcher = DistributedLauncher(cfg, short_cart_pole)
    trainer = Trainer(limit_train_batches=2, max_steps=2).fit(launcher)
    assert trainer.global_step == 2 and np.isfinite(launcher.logged["total_reward"])
    # A3C on its gradient server of one
    cfg = _config("A3C", "CartPole-v1", tmp_path, models=[Policy, "Critic"], model_kwargs=[{}, {"state_dim": 4}],
                  batch_size=8, replay_size=1024, actor_update_times=1, critic_update_times=1)
    launcher = DistributedLauncher(cfg, short_cart_pole)
    Trainer(limit_train_batches=1, max_steps=2).fit(launcher)
    assert np.isfinite(launcher.logged["total_reward"])
