"""Two classic control tasks with the gym calling convention (``reset() -> obs``, ``step(a) -> obs, reward, done, info``,
``action_space`` / ``observation_space``), so that the launcher has something to run where ``gym`` is not installed.

Written from the textbook equations, not from gym's sources: the cart-pole of Barto, Sutton & Anderson (1983) with the
usual constants (pole half-length 0.5 m, masses 1.0 / 0.1 kg, 10 N pushes, 20 ms Euler steps, failure beyond 12 degrees
or 2.4 m, episodes capped at 500 steps) and the torque-limited pendulum swing-up (m = l = 1, g = 10, |u| <= 2,
|omega| <= 8, 50 ms steps, 200 steps, cost angle^2 + 0.1 omega^2 + 0.001 u^2).  This module is also an environment
module of its own (``env: "classic_control"``)."""
import math
from copy import deepcopy

import numpy as np


class Discrete:
    def __init__(self, n: int):
        self.n, self.shape, self.dtype = int(n), (), np.int64


class Box:
    def __init__(self, low, high, shape):
        self.low = np.broadcast_to(np.asarray(low, dtype=np.float32), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=np.float32), shape).copy()
        self.shape, self.dtype = tuple(shape), np.float32


class CartPole:
    GRAVITY, CART, POLE, HALF_LENGTH, PUSH, DT = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    ANGLE_LIMIT, TRACK_LIMIT, MAX_STEPS = 12.0 * math.pi / 180.0, 2.4, 500

    def __init__(self, seed: int = 0):
        self.action_space = Discrete(2)
        bound = np.array([2 * self.TRACK_LIMIT, np.inf, 2 * self.ANGLE_LIMIT, np.inf], dtype=np.float32)
        self.observation_space = Box(-bound, bound, (4,))
        self.rng = np.random.RandomState(seed)
        self.state, self.steps = None, 0

    def seed(self, seed):
        self.rng = np.random.RandomState(seed)

    def reset(self):
        self.state, self.steps = self.rng.uniform(-0.05, 0.05, size=4), 0
        return self.state.astype(np.float32)

    def step(self, action):
        x, v, th, w = self.state
        force = self.PUSH if int(action) == 1 else -self.PUSH
        total, pole_ml = self.CART + self.POLE, self.POLE * self.HALF_LENGTH
        sin, cos = math.sin(th), math.cos(th)
        tmp = (force + pole_ml * w * w * sin) / total
        alpha = (self.GRAVITY * sin - cos * tmp) / (self.HALF_LENGTH * (4.0 / 3.0 - self.POLE * cos * cos / total))
        acc = tmp - pole_ml * alpha * cos / total
        self.state = np.array([x + self.DT * v, v + self.DT * acc, th + self.DT * w, w + self.DT * alpha])
        self.steps += 1
        fell = abs(self.state[0]) > self.TRACK_LIMIT or abs(self.state[2]) > self.ANGLE_LIMIT
        return self.state.astype(np.float32), 1.0, bool(fell or self.steps >= self.MAX_STEPS), {}


class Pendulum:
    MAX_SPEED, MAX_TORQUE, DT, G, M, L, MAX_STEPS = 8.0, 2.0, 0.05, 10.0, 1.0, 1.0, 200

    def __init__(self, seed: int = 0):
        self.action_space = Box(-self.MAX_TORQUE, self.MAX_TORQUE, (1,))
        self.observation_space = Box([-1.0, -1.0, -self.MAX_SPEED], [1.0, 1.0, self.MAX_SPEED], (3,))
        self.rng = np.random.RandomState(seed)
        self.th, self.w, self.steps = 0.0, 0.0, 0

    def seed(self, seed):
        self.rng = np.random.RandomState(seed)

    def _obs(self):
        return np.array([math.cos(self.th), math.sin(self.th), self.w], dtype=np.float32)

    def reset(self):
        self.th, self.w, self.steps = self.rng.uniform(-math.pi, math.pi), self.rng.uniform(-1.0, 1.0), 0
        return self._obs()

    def step(self, action):
        u = float(np.clip(np.asarray(action, dtype=np.float64).reshape(-1)[0], -self.MAX_TORQUE, self.MAX_TORQUE))
        wrapped = (self.th + math.pi) % (2 * math.pi) - math.pi          # angle from upright in [-pi, pi)
        cost = wrapped ** 2 + 0.1 * self.w ** 2 + 0.001 * u ** 2
        self.w += (3 * self.G / (2 * self.L) * math.sin(self.th) + 3.0 / (self.M * self.L ** 2) * u) * self.DT
        self.w = float(np.clip(self.w, -self.MAX_SPEED, self.MAX_SPEED))
        self.th += self.w * self.DT
        self.steps += 1
        return self._obs(), -cost, self.steps >= self.MAX_STEPS, {}


REGISTRY = {"CartPole-v1": CartPole, "CartPole-v0": CartPole, "Pendulum-v0": Pendulum, "Pendulum-v1": Pendulum}


def make(name: str, seed: int = 0):
    if name not in REGISTRY:
        raise ValueError(f"Unknown classic control task: {name}, valid ones are: {sorted(REGISTRY)}")
    return REGISTRY[name](seed=seed)


def generate_env_config(config=None):
    from ..config import fill_default
    config = deepcopy(config) if config else {}
    part = {"env_name": "CartPole-v1", "render_every_episode": 100, "act_kwargs": {}}
    return fill_default({"env": "classic_control", "train_env_config": dict(part), "test_env_config": dict(part)}, config)


def launch(config, pl_callbacks=None):
    from .openai_gym import launch_with
    return launch_with(config, pl_callbacks, lambda frame, env_config: _dataset(frame, env_config))


def _dataset(frame, env_config):
    from .openai_gym import dataset_for
    return dataset_for(frame, make(env_config["env_name"], seed=env_config.get("seed", 0)), env_config)
