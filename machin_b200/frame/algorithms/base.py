"""Minimal TorchFramework base (machin/frame/algorithms/base.py:11-184): model registry + save/load."""
import os
from typing import Callable, Dict

import torch as t


class TorchFramework:
    _is_top = []
    _is_restorable = []

    def __init__(self):
        self._backward = t.autograd.backward
        self._visualized = set()

    @property
    def optimizers(self):
        raise NotImplementedError

    @property
    def lr_schedulers(self):
        raise NotImplementedError

    @property
    def top_models(self):
        return [getattr(self, m) for m in self._is_top]

    @property
    def restorable_models(self):
        return [getattr(self, m) for m in self._is_restorable]

    @property
    def backward_function(self):
        return self._backward

    @classmethod
    def get_top_model_names(cls):
        return cls._is_top

    @classmethod
    def get_restorable_model_names(cls):
        return cls._is_restorable

    @classmethod
    def is_distributed(cls):
        return False

    def set_backward_function(self, backward_func: Callable):
        assert callable(backward_func), "Backward function must be callable."
        self._backward = backward_func

    def save(self, model_dir: str, network_map: Dict[str, str] = None, version: int = 0):
        network_map = network_map or {}
        os.makedirs(model_dir, exist_ok=True)
        for r in self._is_restorable:
            t.save(getattr(self, r), os.path.join(model_dir, f"{network_map.get(r, r)}_{version}.pt"))

    def load(self, model_dir: str, network_map: Dict[str, str] = None, version: int = -1):
        network_map = network_map or {}
        for r in self._is_restorable:
            name = network_map.get(r, r)
            if version < 0:
                found = [int(f[len(name) + 1:-3]) for f in os.listdir(model_dir)
                         if f.startswith(name + "_") and f.endswith(".pt")]
                if not found:
                    raise RuntimeError(f"no checkpoint for {name} in {model_dir}")
                v = max(found)
            else:
                v = version
            loaded = t.load(os.path.join(model_dir, f"{name}_{v}.pt"), map_location="cpu", weights_only=False)
            getattr(self, r).load_state_dict(loaded.state_dict() if hasattr(loaded, "state_dict") else loaded)

    def visualize_model(self, *_, **__):
        pass
