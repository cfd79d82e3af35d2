"""Uniform replay buffer over the device ring, with the reference's ``Buffer`` surface.

Mirror of machin/frame/buffers/buffer.py:11-432.  Episode bookkeeping (which episode owns
which slot, whole-episode eviction when one of its slots is overwritten, :93-109) stays on the
host; transitions live in HBM (:class:`DeviceRing`) and batch assembly -- the reference's list
gather + ``torch.cat`` (:261-353, :380-413), ~97 % of its sample time -- is one gather kernel.

Sampling order: the reference draws from ``list(transition_episode_number.keys())``, a dict in
insertion order, which is "live episodes ascending, slots in stored order"; the same list is
rebuilt here from ``episode_transition_handles`` so Python's ``random`` stream is consumed
identically.
"""
import random
from typing import Any, Callable, Dict, List, Tuple, Union

import numpy as np
import torch as t

from ... import _lib
from ..transition import Transition, TransitionBase
from .device_ring import DeviceRing, Field

Scalar = Union[int, float, bool]


def _scalar_dtype(v):
    return t.tensor([v]).dtype


class EpisodeBook:
    """Host-side episode bookkeeping of ``Buffer`` (buffer.py:41-44, :93-109) plus the ring cursor of
    ``TransitionStorageBasic`` (storage.py:97-114), without any storage behind it.

    ``handles``: episode -> slots in stored order; a slot overwritten by a newer episode evicts the WHOLE older
    episode from the sampleable set.  ``sampleable()`` lists live slots exactly in the order of the
    reference's ``transition_episode_number`` dict (episodes ascending, slots in stored order)."""

    def __init__(self, capacity: int):
        self.capacity = int(capacity)
        self.count = 0                      # len(storage)
        self.cursor = 0                     # storage.index
        self.episode_counter = 0
        self.handles = {}                   # type: Dict[int, np.ndarray]
        self.slot_episode = np.full(self.capacity, -1, dtype=np.int64)
        # (first slot, length) per episode number: an episode's slots are consecutive modulo the capacity, so whole-
        # episode sampling can build its slot list with array arithmetic instead of one Python object per episode
        self._ep_first = np.zeros(1024, dtype=np.int64)
        self._ep_len = np.zeros(1024, dtype=np.int64)

    def assign(self, n: int):
        """Reserve ``n`` consecutive ring slots -> (first_slot, positions)."""
        if n > self.capacity:
            raise ValueError("Failed to store episode.")
        first = self.count if self.count < self.capacity else self.cursor
        positions = (first + np.arange(n, dtype=np.int64)) % self.capacity
        self.count = min(self.capacity, self.count + n)
        self.cursor = int((positions[-1] + 1) % self.capacity)
        return int(first), positions

    def record(self, positions: np.ndarray) -> int:
        episode_number = self.episode_counter
        self.episode_counter += 1
        old = np.unique(self.slot_episode[positions])
        for ep in old[old >= 0]:            # evict whole older episodes (buffer.py:95-105)
            slots = np.asarray(self.handles.pop(int(ep)))
            self.slot_episode[slots[self.slot_episode[slots] == ep]] = -1
        self.slot_episode[positions] = episode_number
        self.handles[episode_number] = positions
        if episode_number >= self._ep_first.shape[0]:
            grow = max(2 * self._ep_first.shape[0], episode_number + 1)
            self._ep_first = np.concatenate([self._ep_first, np.zeros(grow - self._ep_first.shape[0], dtype=np.int64)])
            self._ep_len = np.concatenate([self._ep_len, np.zeros(grow - self._ep_len.shape[0], dtype=np.int64)])
        self._ep_first[episode_number] = positions[0]
        self._ep_len[episode_number] = len(positions)
        return episode_number

    def live_episode_range(self):
        """(lo, hi) if the live episodes are exactly the numbers lo .. hi-1 (the ring evicts oldest first, so this is
        the normal case), else None."""
        n = len(self.handles)
        if n == 0:
            return None
        lo = next(iter(self.handles))
        return (lo, lo + n) if lo + n == self.episode_counter else None

    def episode_slots(self, episodes: np.ndarray):
        """Slots of the given episode numbers, concatenated in order -> (slots i64[sum len], lengths i64[k])."""
        episodes = np.asarray(episodes, dtype=np.int64)
        lens = self._ep_len[episodes]
        total = int(lens.sum())
        if total == 0:
            return np.empty(0, dtype=np.int64), lens
        begin = np.cumsum(lens) - lens
        within = np.arange(total, dtype=np.int64) - np.repeat(begin, lens)
        return (np.repeat(self._ep_first[episodes], lens) + within) % self.capacity, lens

    def sampleable(self) -> List[int]:
        return [int(s) for slots in self.handles.values() for s in slots]

    def clear(self):
        self.count = 0
        self.handles = {}
        self.slot_episode.fill(-1)


class Buffer:
    def __init__(self, buffer_size: int = 1000000, buffer_device: Union[str, t.device] = "cuda",
                 storage=None, **__):
        if storage is not None:
            raise ValueError("machin_b200 buffers own a device ring; custom storage is not supported")
        dev = t.device(buffer_device)
        if dev.type != "cuda":
            raise ValueError(
                "machin_b200 buffers are device-resident: buffer_device must be a CUDA device "
                f"(got {buffer_device!r}); there is no CPU fallback on this path")
        self.buffer_device = t.device("cuda", dev.index if dev.index is not None else 0)
        _lib.require_device(self.buffer_device.index)
        self.storage = DeviceRing(int(buffer_size), self.buffer_device)
        self.book = EpisodeBook(int(buffer_size))
        self._major_attr: List[str] = []
        self._sub_attr: List[str] = []
        self._custom_attr: List[str] = []
        self._host_custom: Dict[str, np.ndarray] = {}

    # ------------------------------------------------------------------ bookkeeping (host side: EpisodeBook)
    @property
    def episode_transition_handles(self) -> Dict[int, Any]:
        return self.book.handles

    @property
    def episode_counter(self) -> int:
        return self.book.episode_counter

    @property
    def transition_episode_number(self) -> Dict[int, int]:
        """slot -> owning episode, in the reference's dict order (buffer.py:42)."""
        return {int(s): ep for ep, slots in self.book.handles.items() for s in slots}

    def _sampleable_handles(self) -> List[int]:
        return self.book.sampleable()

    def _record_episode(self, positions: np.ndarray) -> int:
        return self.book.record(positions)

    def _assign_slots(self, n: int):
        first, positions = self.book.assign(n)
        self.storage.count, self.storage.index = self.book.count, self.book.cursor
        return first, positions

    # ------------------------------------------------------------------ schema
    def _make_schema(self, tr: TransitionBase):
        fields = []
        for attr in tr.major_attr:
            for key, v in tr[attr].items():
                fields.append(Field(f"{attr}/{key}", attr, key, v.dtype, v.shape[1:], "major"))
        for attr in tr.sub_attr:
            v = tr[attr]
            if t.is_tensor(v):
                fields.append(Field(attr, attr, None, v.dtype, v.shape[1:], "sub"))
            else:
                fields.append(Field(attr, attr, None, _scalar_dtype(v), (1,), "sub", scalar=True))
        for attr in tr.custom_attr:
            v = tr[attr]
            if isinstance(v, (bool, int, float, np.number)) and not t.is_tensor(v):
                fields.append(Field(attr, attr, None, _scalar_dtype(v), (1,), "custom", scalar=True))
        self._major_attr, self._sub_attr, self._custom_attr = list(tr.major_attr), list(tr.sub_attr), list(tr.custom_attr)
        self.storage.set_schema(fields)

    # ------------------------------------------------------------------ store
    def store_episode(self, episode: List[Union[TransitionBase, Dict]],
                      required_attrs=("state", "action", "next_state", "reward", "terminal")):
        if len(episode) == 0:
            raise ValueError("Episode must be non-empty.")
        for idx, transition in enumerate(episode):
            if isinstance(transition, dict):
                transition = Transition(**transition)
            elif not isinstance(transition, TransitionBase) and not hasattr(transition, "major_attr"):
                raise ValueError("Transition object must be a dict or an instance of the Transition class.")
            if not transition.has_keys(required_attrs):
                missing_keys = set(required_attrs) - set(transition.keys())
                raise ValueError(f"Transition object missing attributes: {missing_keys}, object is {transition}.")
            episode[idx] = transition
        if self.storage.data is None:
            self._make_schema(episode[0])
        n = len(episode)
        first, positions = self._assign_slots(n)
        dense = {}
        for f in self.storage.fields.values():
            if f.kind == "major":
                col = t.cat([tr[f.attr][f.key] for tr in episode], dim=0)
            else:
                vals = [tr[f.attr] for tr in episode]
                col = t.cat(vals, dim=0) if t.is_tensor(vals[0]) else t.tensor(vals, dtype=f.dtype).view(n, 1)
            if tuple(col.shape[1:]) != f.shape or col.dtype != f.dtype:
                raise ValueError(f"Transition field {f.name} changed shape/dtype: {col.dtype}{tuple(col.shape[1:])}"
                                 f" vs stored {f.dtype}{f.shape}")
            dense[f.name] = col.to(self.buffer_device).contiguous()
        self.storage.write(first, n, dense)
        for attr in self._custom_attr:
            if attr not in self._host_custom:
                self._host_custom[attr] = np.empty(self.storage.max_size, dtype=object)
            col = self._host_custom[attr]
            for p, tr in zip(positions, episode):
                col[p] = tr[attr] if attr in tr.keys() else None
        self._record_episode(positions)
        return positions

    def store_episode_tensors(self, major: Dict[str, Dict[str, t.Tensor]], sub: Dict[str, t.Tensor],
                              custom: Dict[str, t.Tensor] = None):
        """Tensor-native ``store_episode``: one episode given as batched ``[n, ...]`` tensors
        (no per-transition Python objects).  Returns the slot positions."""
        custom = custom or {}
        n = int(next(iter(next(iter(major.values())).values())).shape[0])
        if n == 0:
            raise ValueError("Episode must be non-empty.")
        if self.storage.data is None:
            fields = [Field(f"{a}/{k}", a, k, v.dtype, v.shape[1:], "major") for a, d in major.items() for k, v in d.items()]
            fields += [Field(a, a, None, v.dtype, v.shape[1:] if v.dim() > 1 else (1,), "sub") for a, v in sub.items()]
            fields += [Field(a, a, None, v.dtype, v.shape[1:] if v.dim() > 1 else (1,), "custom") for a, v in custom.items()]
            self._major_attr, self._sub_attr, self._custom_attr = list(major), list(sub), sorted(custom)
            self.storage.set_schema(fields)
        first, positions = self._assign_slots(n)
        dense = {}
        for f in self.storage.fields.values():
            src = major[f.attr][f.key] if f.kind == "major" else (sub[f.attr] if f.kind == "sub" else custom[f.attr])
            dense[f.name] = src.to(device=self.buffer_device, dtype=f.dtype).reshape((n,) + f.shape).contiguous()
        self.storage.write(first, n, dense)
        self._record_episode(positions)
        return positions

    def size(self):
        return len(self.storage)

    def clear(self):
        # The reference only empties the storage (buffer.py:117-121) and leaves stale handles in its
        # dicts (sampling them raises IndexError); here the bookkeeping is dropped as well.
        self.storage.clear()
        self.book.clear()

    # ------------------------------------------------------------------ sample
    def sample_batch(self, batch_size: int, concatenate: bool = True, device: Union[str, t.device] = None,
                     sample_method: Union[Callable, str] = "random_unique", sample_attrs: List[str] = None,
                     additional_concat_custom_attrs: List[str] = None, *_, **__) -> Tuple[int, Union[None, tuple]]:
        if isinstance(sample_method, str):
            if not hasattr(self, "sample_method_" + sample_method):
                raise RuntimeError(f"Cannot find specified sample method: {sample_method}")
            batch_size, handles = getattr(self, "sample_method_" + sample_method)(batch_size)
        else:
            batch_size, handles = sample_method(self, batch_size)
        return batch_size, self.post_process_batch(handles, device, concatenate, sample_attrs,
                                                   additional_concat_custom_attrs)

    def sample_method_random_unique(self, batch_size: int):
        keys = self._sampleable_handles()
        batch_size = min(len(keys), batch_size)
        return batch_size, random.sample(keys, k=batch_size)

    def sample_method_random(self, batch_size: int):
        keys = self._sampleable_handles()
        batch_size = min(len(keys), batch_size)
        return batch_size, random.choices(keys, k=batch_size)

    def sample_method_all(self, _):
        keys = self._sampleable_handles()
        return len(keys), keys

    def post_process_batch(self, handles, device, concatenate, sample_attrs, additional_concat_custom_attrs,
                           index_tensor: t.Tensor = None, packed_rows: t.Tensor = None):
        """``handles``: slot list / array (host).  ``index_tensor``: the same slots already on
        the device (PER path) -- avoids the host round trip.  ``packed_rows``: rows that were
        already gathered elsewhere (sharded buffers) -- only unpacked here."""
        if packed_rows is not None:
            n = int(packed_rows.shape[0])
        else:
            n = len(handles) if handles is not None else int(index_tensor.shape[0])
        if n == 0:
            return None
        if index_tensor is None and packed_rows is None:
            index_tensor = t.as_tensor(np.asarray(handles, dtype=np.int64)).to(self.buffer_device)
        out_device = self.buffer_device if device is None else t.device(device)
        if sample_attrs is None:
            sample_attrs = self._major_attr + self._sub_attr + self._custom_attr
        if additional_concat_custom_attrs is None:
            additional_concat_custom_attrs = []
        fields = self.storage.fields
        want = []
        for attr in sample_attrs:
            if attr in self._major_attr:
                want += [f.name for f in fields.values() if f.attr == attr and f.kind == "major"]
            elif attr in self._sub_attr:
                want.append(attr)
            elif attr in fields and concatenate and attr in additional_concat_custom_attrs:
                want.append(attr)
            elif attr == "*":
                want += [a for a in self._custom_attr if a in fields and concatenate and a in additional_concat_custom_attrs]
        if packed_rows is not None:
            dense = self.storage.unpack(packed_rows, names=list(dict.fromkeys(want)))
        else:
            dense = self.storage.gather(index_tensor, names=list(dict.fromkeys(want)))
        host_idx = None

        def finish(attr, key, tensor, scalar=False):
            if not concatenate and scalar:
                # buffer.py:402-413 returns the sampled Python values as they were stored (reward 10.0, terminal True)
                return self.post_process_attribute(attr, key, tensor.reshape(-1).tolist())
            tensor = tensor.to(out_device)
            if not concatenate:
                tensor = list(tensor.split(1, dim=0))
            return self.post_process_attribute(attr, key, tensor)

        def custom_value(attr):
            nonlocal host_idx
            if concatenate and attr in additional_concat_custom_attrs:
                if attr not in dense:
                    raise ValueError(f"Batch not concatenable: custom attribute {attr!r} is not numeric")
                return self.post_process_attribute(attr, None, dense[attr].view(n, -1).to(out_device))
            if attr not in self._host_custom or packed_rows is not None:
                if attr not in self.storage.fields:
                    raise ValueError(f"custom attribute {attr!r} is not numeric and cannot be sampled from a "
                                     "device / sharded buffer")
                src = (self.storage.unpack(packed_rows, names=[attr]) if packed_rows is not None
                       else self.storage.gather(index_tensor, names=[attr]))
                return self.post_process_attribute(attr, None, src[attr].view(-1).tolist())
            if host_idx is None:
                host_idx = np.asarray(handles) if handles is not None else index_tensor.cpu().numpy()
            return self.post_process_attribute(attr, None, list(self._host_custom[attr][host_idx]))

        result, used = [], []
        for attr in sample_attrs:
            if attr in self._major_attr:
                result.append({f.key: finish(attr, f.key, dense[f.name])
                               for f in fields.values() if f.attr == attr and f.kind == "major"})
                used.append(attr)
            elif attr in self._sub_attr:
                result.append(finish(attr, None, dense[attr].view(n, -1), fields[attr].scalar))
                used.append(attr)
            elif attr in self._custom_attr:
                result.append(custom_value(attr))
                used.append(attr)
            elif attr == "*":
                tmp = {}
                for k in self._custom_attr:
                    if k not in used:
                        tmp[k] = custom_value(k)
                        used.append(k)
                result.append(tmp)
        return tuple(result)

    # hooks kept for signature compatibility (buffer.py:355-432)
    def pre_process_attribute(self, attribute: Any, sub_key: Any, values):
        return values

    def make_tensor_from_batch(self, batch, device, concatenate: bool):
        if concatenate and len(batch) != 0:
            if t.is_tensor(batch[0]):
                return t.cat([it.to(device) for it in batch], dim=0).to(device)
            try:
                return t.tensor(batch, device=device).view(len(batch), -1)
            except Exception:
                raise ValueError(f"Batch not concatenable: {batch}")
        return batch

    def post_process_attribute(self, attribute: Any, sub_key: Any, values):
        return values
