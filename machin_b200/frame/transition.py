"""Transition record accepted by the device buffers.

Host-side mirror of machin/frame/transition.py:224-286 (class Transition): three major
attributes (dicts of ``[1, ...]`` tensors), two sub attributes (``reward``: scalar or
tensor, ``terminal``: bool or bool tensor) and alphabetically ordered custom attributes.
The device buffers only read a transition once -- to pack it into a replay row -- so this
class is a validating view, not a storage format.
"""
from typing import Any, Dict, Iterable

import numpy as np
import torch as t

MAJOR = ("state", "action", "next_state")
SUB = ("reward", "terminal")


class TransitionBase:
    def __init__(self, major_attr, sub_attr, custom_attr, major_data, sub_data, custom_data):
        self._major_attr = list(major_attr)
        self._sub_attr = list(sub_attr)
        self._custom_attr = list(custom_attr)
        self._keys = self._major_attr + self._sub_attr + self._custom_attr
        self._data = {}
        for k, v in zip(self._keys, list(major_data) + list(sub_data) + list(custom_data)):
            self._data[k] = v
        for k in self._major_attr:
            self._data[k] = {kk: vv.detach() for kk, vv in self._data[k].items()}
        for k in self._sub_attr:
            if t.is_tensor(self._data[k]):
                self._data[k] = self._data[k].detach()
        self._check_validity()

    major_attr = property(lambda self: self._major_attr)
    sub_attr = property(lambda self: self._sub_attr)
    custom_attr = property(lambda self: self._custom_attr)

    def __len__(self):
        return len(self._keys)

    def __getitem__(self, item):
        return self._data[item]

    def __getattr__(self, item):
        data = self.__dict__.get("_data")
        if data is not None and item in data:
            return data[item]
        raise AttributeError(item)

    def __setitem__(self, key, value):
        if key not in self._keys:
            raise RuntimeError("You cannot dynamically set new attributes in a Transition object!")
        self._data[key] = value
        self._check_validity()

    def keys(self):
        return self._keys

    def items(self):
        for k in self._keys:
            yield k, self._data[k]

    def has_keys(self, keys: Iterable[str]):
        return all(k in self._keys for k in keys)

    def to(self, device):
        for k in self._major_attr:
            self._data[k] = {kk: vv.to(device) for kk, vv in self._data[k].items()}
        for k in self._sub_attr:
            if t.is_tensor(self._data[k]):
                self._data[k] = self._data[k].to(device)
        return self

    def _check_validity(self):
        batch = None
        for ma in self._major_attr:
            for k, v in self._data[ma].items():
                if not t.is_tensor(v) or v.dim() < 1:
                    raise ValueError(f'Key "{k}" of transition major attribute "{ma}" is an invalid tensor')
                if batch is None:
                    batch = v.shape[0]
                elif batch != v.shape[0]:
                    raise ValueError(
                        f'Key "{k}" of transition major attribute "{ma}" has invalid batch size {v.shape[0]}.')
        for sa in self._sub_attr:
            v = self._data[sa]
            if np.isscalar(v):
                if batch != 1:
                    raise ValueError(f'Transition sub attribute "{sa}" is a scalar, but batch size is {batch}.')
            elif t.is_tensor(v):
                if v.dim() < 1:
                    raise ValueError(f'Transition sub attribute "{sa}" is an invalid tensor.')
                if v.shape[0] != batch:
                    raise ValueError(f'Transition sub attribute "{sa}" has invalid batch size {v.shape[0]}.')
            else:
                raise ValueError(
                    f'Transition sub attribute "{sa}" has invalid value {v}, requires scalar or tensor.')
        self._batch_size = batch


class Transition(TransitionBase):
    def __init__(self, state: Dict[str, t.Tensor], action: Dict[str, t.Tensor],
                 next_state: Dict[str, t.Tensor], reward: Any, terminal: Any, **kwargs):
        custom = sorted(kwargs.keys())
        assert isinstance(terminal, bool) or (t.is_tensor(terminal) and terminal.dtype == t.bool)
        super().__init__(MAJOR, SUB, custom, (state, action, next_state), (reward, terminal),
                         [kwargs[k] for k in custom])
        if self._batch_size != 1:
            raise ValueError("Batch size of the default transition implementation must be 1.")
