"""Automation layer above the learner (surface of machin/auto/__init__.py:1-17): config generation, episode datasets and
the launcher loop ``store_episode -> update`` (machin/auto/launcher.py:61-66 and :152-171) that feeds the device-resident
hot path.  The reference drives this loop through a pytorch-lightning ``Trainer`` and boots its RPC ``World`` from a patched
DDP plugin (pl_plugin.py:34-69); neither exists here -- ranks are the SPMD processes of one ``torchrun`` job, and
:class:`launcher.Trainer` is the plain loop with the knobs ``envs.openai_gym.launch`` sets."""
from . import envs  # first: envs.openai_gym imports .config, which looks up this (then partially built) package
from . import config, dataset, launcher  # noqa: E402

__all__ = ["envs", "config", "dataset", "launcher"]
