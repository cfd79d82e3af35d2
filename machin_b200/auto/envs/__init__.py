"""Environment modules of the launcher: each offers ``generate_env_config(config)`` and ``launch(config, pl_callbacks)``
(machin/auto/envs/__init__.py, machin/auto/config.py:34-40)."""
from . import classic_control, openai_gym

__all__ = ["openai_gym", "classic_control"]
