"""Fixed-length sequence sampling over the device ring: reference ``RNNBuffer`` / ``RNNPrioritizedBuffer``.

Mirror of machin/frame/buffers/rnn_buffers.py:18-182 (uniform) and :259-413 (prioritized) -- SURVEY §8(f)
rank 4: callers of the same tree / draw / gather kernels, no new device code.  A sample is ``sample_length``
consecutive steps of one episode; a batch of B samples is gathered as B * sample_length packed rows by the
row-gather kernel and viewed as ``[B, sample_length, ...]`` (:169-182; ``sample_dimension`` is accepted and
stored like the reference does, which never reads it).

Which sequences are drawn is host logic (Python ``random`` for the uniform buffer, the PER draw for the
prioritized one) and lives in the module-level functions below so that it can be pinned on the CPU against
the unmodified reference (tests/golden/ref_rnn.npz).
"""
import random
from typing import Dict, List, Sequence, Tuple, Union

import numpy as np
import torch as t

from ..transition import TransitionBase
from .buffer import Buffer
from .prioritized_buffer import PrioritizedBuffer


# ---------------------------------------------------------------------- host logic (CPU-testable)
def _valid_episodes(handles: Dict[int, Sequence[int]], sample_length: int) -> List[int]:
    return [ep for ep, slots in handles.items() if len(slots) >= sample_length]


def _sequences(handles, sample_length, episodes, starts) -> List[int]:
    return [int(s) for ep, pos in zip(episodes, starts) for s in handles[ep][pos:pos + sample_length]]


def rnn_sample_random_unique(handles, sample_length: int, batch_size: int) -> Tuple[int, List[int]]:
    """Distinct episodes, then one start per episode (:78-106): ``random.sample`` followed by one
    ``random.randint`` per selected episode, in that order."""
    valid = _valid_episodes(handles, sample_length)
    batch_size = min(len(valid), batch_size)
    episodes = random.sample(valid, k=batch_size)
    starts = [random.randint(0, len(handles[ep]) - sample_length) for ep in episodes]
    return len(episodes), _sequences(handles, sample_length, episodes, starts)


def rnn_sample_random(handles, sample_length: int, batch_size: int) -> Tuple[int, List[int]]:
    """Episodes with replacement (:108-141)."""
    valid = _valid_episodes(handles, sample_length)
    batch_size = min(len(valid), batch_size)
    episodes = random.choices(valid, k=batch_size)
    starts = [random.randint(0, len(handles[ep]) - sample_length) for ep in episodes]
    return len(episodes), _sequences(handles, sample_length, episodes, starts)


def rnn_sample_all(handles, sample_length: int) -> Tuple[int, List[int]]:
    """Every complete window of every long-enough episode (:143-167)."""
    out, count = [], 0
    for ep in _valid_episodes(handles, sample_length):
        slots = handles[ep]
        for pos in range(len(slots) - sample_length + 1):
            out += [int(s) for s in slots[pos:pos + sample_length]]
            count += 1
    return count, out


def rnn_episode_priorities(n: int, sample_length: int, priorities, default_priority, normalize):
    """Leaf weights of one stored episode (:343-380): steps that cannot start a complete window get 0 so that
    the tree never draws them.  ``priorities``: None (-> ``default_priority`` for the startable steps) or an
    ndarray, which is normalised; like the reference, a short episode's array is zeroed IN PLACE."""
    if priorities is None:
        return [default_priority if i + sample_length <= n else 0 for i in range(n)]
    if n < sample_length:
        priorities[:] = 0
        return priorities
    priorities = normalize(priorities)
    priorities[n - sample_length + 1:] = 0
    return priorities


def _chunk(values, sample_length: int):
    """:169-182 -- lists are cut into windows, tensors get a window axis after the batch axis."""
    if isinstance(values, list):
        return [values[i:i + sample_length] for i in range(0, len(values), sample_length)]
    return values.reshape([int(values.shape[0] / sample_length), sample_length] + list(values.shape)[1:])


# ---------------------------------------------------------------------- buffers
class RNNBuffer(Buffer):
    def __init__(self, sample_length: int, sample_dimension: int = 1, buffer_size: int = 1000000,
                 buffer_device: Union[str, t.device] = "cuda", storage=None, **kwargs):
        if storage is not None:
            kwargs["storage"] = storage               # rejected by Buffer: the ring is the storage
        super().__init__(buffer_size=buffer_size, buffer_device=buffer_device, **kwargs)
        self.sample_length = sample_length
        self.sample_dimension = sample_dimension

    def sample_method_random_unique(self, batch_size: int):
        return rnn_sample_random_unique(self.book.handles, self.sample_length, batch_size)

    def sample_method_random(self, batch_size: int):
        return rnn_sample_random(self.book.handles, self.sample_length, batch_size)

    def sample_method_all(self, _):
        return rnn_sample_all(self.book.handles, self.sample_length)

    def post_process_attribute(self, attribute, sub_key, values):
        return _chunk(values, self.sample_length)


class RNNPrioritizedBuffer(RNNBuffer, PrioritizedBuffer):
    def __init__(self, sample_length: int, sample_dimension: int = 1, buffer_size: int = 1000000,
                 buffer_device: Union[str, t.device] = "cuda", epsilon: float = 1e-2, alpha: float = 0.6,
                 beta: float = 0.4, beta_increment_per_sampling: float = 0.001, **kwargs):
        super().__init__(sample_length=sample_length, sample_dimension=sample_dimension, buffer_size=buffer_size,
                         buffer_device=buffer_device, epsilon=epsilon, alpha=alpha, beta=beta,
                         beta_increment_per_sampling=beta_increment_per_sampling, **kwargs)

    def store_episode(self, episode: List[Union[TransitionBase, Dict]], priorities=None,
                      required_attrs=("state", "action", "next_state", "reward", "terminal")):
        positions = Buffer.store_episode(self, episode, required_attrs)
        default = self._default_priority() if priorities is None else None
        self.wt_tree.update_leaf_batch(
            rnn_episode_priorities(len(episode), self.sample_length, priorities, default, self._normalize_priority),
            positions)
        return positions

    def store_episode_tensors(self, major, sub, custom=None, priorities=None):
        """Tensor-native store; ``priorities``: None or a host array (copied, not modified)."""
        positions = Buffer.store_episode_tensors(self, major, sub, custom)
        default = self._default_priority() if priorities is None else None
        pri = None if priorities is None else np.array(priorities, dtype=np.float64)
        leaf = rnn_episode_priorities(len(positions), self.sample_length, pri, default, self._normalize_priority)
        self.wt_tree.update_leaf_batch(np.asarray(leaf, dtype=np.float64), positions)
        return positions

    def _window_index(self, d_index: t.Tensor) -> t.Tensor:
        # :391-395 reads storage[idx .. idx + L - 1]; an episode that wrapped around the end of the ring
        # continues at slot 0 (the reference raises IndexError there) -- hence the modulo
        steps = t.arange(self.sample_length, device=d_index.device, dtype=t.int64)
        return ((d_index.view(-1, 1) + steps) % self.storage.max_size).reshape(-1)

    def sample_batch(self, batch_size: int, concatenate: bool = True, device: Union[str, t.device] = None,
                     sample_attrs: List[str] = None, additional_concat_custom_attrs: List[str] = None, *_, **__):
        if batch_size <= 0 or self.size() == 0:
            return 0, None, None, None
        d_index, d_isw = self._draw(batch_size)
        result = self.post_process_batch(None, device, concatenate, sample_attrs, additional_concat_custom_attrs,
                                         index_tensor=self._window_index(d_index))
        index, is_weight = d_index.cpu().numpy(), d_isw.cpu().numpy()
        return len(index), result, index, is_weight

    def sample_device(self, batch_size: int, uniform: t.Tensor = None, names=None, use_device_params=False):
        """Device path: fields come back as ``[n, sample_length, ...]``; index / is_weight per window."""
        d_index, d_isw = self._draw(batch_size, uniform=uniform, use_device_params=use_device_params)
        fields = self.storage.gather(self._window_index(d_index), names=names)
        return batch_size, {k: _chunk(v, self.sample_length) for k, v in fields.items()}, d_index, d_isw
