"""Config generation and launch (surface of machin/auto/config.py:12-142).

A launch config is one flat mapping holding three parts: the training part (``root_dir``, ``episode_per_epoch``,
``max_episodes``, ``early_stopping_patience``), the algorithm part (``frame``, ``frame_config`` and, for the sharded
frameworks, ``batch_num`` -- frame/algorithms/config.py) and the environment part (``env``, ``train_env_config``,
``test_env_config`` -- written by the environment module of that name in ``machin_b200.auto.envs``).

Different from the reference: ``generate_env_config`` tests ``callable(...)`` (the reference's
``isinstance(x, callable)`` at config.py:83-86 raises TypeError for every input), its error message lists environments
rather than algorithms, and the process layout keys are ``gpus`` (devices of this node, one rank each) and ``num_nodes``
for a ``torchrun`` launch instead of pytorch-lightning's ``gpus`` / ``num_processes``."""
import inspect
from copy import deepcopy
from typing import Any, Dict, List

from ..frame import algorithms
from ..frame.algorithms.base import TorchFramework
from . import envs


def fill_default(default: Dict[str, Any], config: Dict[str, Any]):
    """Add to ``config`` the keys of ``default`` it does not have; returns ``config``."""
    for key, value in default.items():
        config.setdefault(key, value)
    return config


def _framework(name):
    cls = getattr(algorithms, name, None) if isinstance(name, str) else None
    if inspect.isclass(cls) and issubclass(cls, TorchFramework) and cls is not TorchFramework:
        return cls
    return None


def _environment(name):
    module = getattr(envs, name, None) if isinstance(name, str) else None
    if inspect.ismodule(module) and callable(getattr(module, "launch", None)) and \
            callable(getattr(module, "generate_env_config", None)):
        return module
    return None


def get_available_algorithms() -> List[str]:
    return [name for name in dir(algorithms) if _framework(name) is not None]


def get_available_environments() -> List[str]:
    return [name for name in dir(envs) if _environment(name) is not None]


def generate_training_config(root_dir: str = "trial", episode_per_epoch: int = 10, max_episodes: int = 10000,
                             config: Dict[str, Any] = None):
    config = deepcopy(config) if config else {}
    config.update(root_dir=root_dir, episode_per_epoch=episode_per_epoch, max_episodes=max_episodes,
                  early_stopping_patience=3)
    return config


def generate_algorithm_config(algorithm: str, config: Dict[str, Any] = None):
    config = deepcopy(config) if config else {}
    cls = _framework(algorithm)
    if cls is None:
        raise ValueError(f"Invalid algorithm: {algorithm}, valid ones are: {get_available_algorithms()}")
    config = cls.generate_config(config)
    # one rank per listed device; the sharded frameworks default to three processes (config.py:65-71), here ranks of one
    # torchrun job that may share a device
    config["gpus"] = [0, 0, 0] if cls.is_distributed() else [0]
    if cls.is_distributed():
        config["num_nodes"] = 1
    return config


def generate_env_config(environment: str, config: Dict[str, Any] = None):
    config = deepcopy(config) if config else {}
    module = _environment(environment)
    if module is None:
        raise ValueError(f"Invalid environment: {environment}, valid ones are: {get_available_environments()}")
    return module.generate_env_config(config)


def assert_training_config_complete(config):
    for key in ("root_dir", "episode_per_epoch", "max_episodes", "early_stopping_patience"):
        assert key in config, f'Missing key "{key}"'


def assert_algorithm_config_complete(config):
    for key in ("frame", "frame_config"):
        assert key in config, f'Missing key "{key}" in config.'


def assert_env_config_complete(config):
    for key in ("env", "train_env_config", "test_env_config"):
        assert key in config, f'Missing key "{key}" in config.'


def _framework_of(config):
    assert_algorithm_config_complete(config)
    cls = _framework(config["frame"])
    if cls is None:
        raise ValueError(f"Invalid algorithm: {config['frame']}, valid ones are: {get_available_algorithms()}")
    return cls


def init_algorithm_from_config(config: Dict[str, Any], model_device="cuda"):
    """Build the framework ``config["frame"]`` names (collective for the sharded ones).  The default device is the GPU:
    replay memory and the update live there (the reference defaults to "cpu", config.py:96-98)."""
    return _framework_of(config).init_from_config(config, model_device=model_device)


def is_algorithm_distributed(config: Dict[str, Any]) -> bool:
    return _framework_of(config).is_distributed()


def launch(config: Dict[str, Any], pl_callbacks: List[Any] = None):
    """Train ``config["frame"]`` on ``config["env"]``; returns what the environment module's ``launch`` returns (the
    fitted trainer).  ``pl_callbacks`` keeps the reference's parameter name; the objects follow ``launcher.Trainer``'s
    callback protocol (``on_step_end`` / ``on_epoch_end``), not pytorch-lightning's."""
    assert_training_config_complete(config)
    assert_env_config_complete(config)
    assert_algorithm_config_complete(config)
    module = _environment(config["env"])
    if module is None:
        raise ValueError(f"Invalid environment: {config['env']}, valid ones are: {get_available_environments()}")
    return module.launch(config, pl_callbacks)
