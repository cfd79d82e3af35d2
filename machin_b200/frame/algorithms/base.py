"""Minimal TorchFramework base (surface of machin/frame/algorithms/base.py:11-184): the model registry the
algorithms declare through ``_is_top`` / ``_is_restorable``, a replaceable backward function, and save / load of the
restorable models as ``<name>_<version>.pt`` files.  Visualisation is a no-op here."""
import os
import re
from typing import Callable, Dict

import torch as t


def _abstract(what):
    def getter(self):
        raise NotImplementedError(f"{type(self).__name__} must define {what}")
    return property(getter)


class TorchFramework:
    _is_top = []          # attribute names of the models a user passes in
    _is_restorable = []   # attribute names of the models save() / load() cover

    optimizers = _abstract("optimizers")
    lr_schedulers = _abstract("lr_schedulers")
    top_models = property(lambda self: [getattr(self, name) for name in self._is_top])
    restorable_models = property(lambda self: [getattr(self, name) for name in self._is_restorable])
    backward_function = property(lambda self: self._backward)

    _config_frame = None  # key of config.DEFAULTS, and how init_from_config builds: "local", "apex", "impala", "a3c"
    _config_kind = None

    def __init__(self):
        self._backward, self._visualized = t.autograd.backward, set()

    @classmethod
    def generate_config(cls, config):
        """A copy of ``config`` with this framework's name and default ``frame_config`` (see config.py)."""
        from .config import generate_config
        return generate_config(cls._config_frame, config)

    @classmethod
    def init_from_config(cls, config, model_device="cuda"):
        """Build the framework from a config: models / optimizer / criterion / lr_scheduler given as classes or names.
        Collective for the distributed frameworks (every rank calls it)."""
        from . import config as builders
        if cls._config_kind == "local":
            return builders.init_local(cls, config, model_device)
        if cls._config_kind == "a3c":
            return builders.init_a3c(cls, config, model_device)
        return builders.init_sharded(cls, config, model_device, cls._config_kind)

    @classmethod
    def get_top_model_names(cls):
        return cls._is_top

    @classmethod
    def get_restorable_model_names(cls):
        return cls._is_restorable

    @classmethod
    def is_distributed(cls):
        return False

    def enable_multiprocessing(self):
        """base.py:86-92: put every top model into shared memory (a no-op for CUDA parameters)."""
        for model in self.top_models:
            model.share_memory()

    def set_backward_function(self, backward_func: Callable):
        if not callable(backward_func):
            raise AssertionError("Backward function must be callable.")
        self._backward = backward_func

    @staticmethod
    def _file_name(name, version):
        return f"{name}_{version}.pt"

    def save(self, model_dir: str, network_map: Dict[str, str] = None, version: int = 0):
        os.makedirs(model_dir, exist_ok=True)
        for attr in self._is_restorable:
            name = (network_map or {}).get(attr, attr)
            t.save(getattr(self, attr), os.path.join(model_dir, self._file_name(name, version)))

    def load(self, model_dir: str, network_map: Dict[str, str] = None, version: int = -1):
        """base.py:94-129 over utils/prepare.py:52-117 (prep_load_model): the requested ``version`` if EVERY restorable
        model has it, otherwise (and for ``version`` < 0) the newest version that all of them share -- never a mixed
        set of checkpoints.  ``RuntimeError`` if the directory is missing or no common version exists."""
        if not os.path.isdir(model_dir):
            raise RuntimeError("Model directory doesn't exist!")
        names = {attr: (network_map or {}).get(attr, attr) for attr in self._is_restorable}
        have = {name: set() for name in names.values()}
        for entry in os.listdir(model_dir):
            m = re.fullmatch(r"([a-zA-Z0-9_-]+)_([0-9]+)\.pt", entry)
            if m is not None and m.group(1) in have:
                have[m.group(1)].add(int(m.group(2)))
        if version is None or version < 0 or not all(version in v for v in have.values()):
            common = set.intersection(*have.values()) if have else set()
            if not common:
                raise RuntimeError("Cannot find a valid version for all models!")
            version = max(common)
        for attr, name in names.items():
            loaded = t.load(os.path.join(model_dir, self._file_name(name, version)), map_location="cpu", weights_only=False)
            getattr(self, attr).load_state_dict(loaded.state_dict() if hasattr(loaded, "state_dict") else loaded)

    def visualize_model(self, *_, **__):
        pass
