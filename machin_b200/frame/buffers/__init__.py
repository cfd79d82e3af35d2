from .buffer import Buffer, EpisodeBook
from .buffer_d import DistributedBuffer, IMPALABuffer
from .device_ring import DeviceRing
from .prioritized_buffer import PrioritizedBuffer
from .prioritized_buffer_d import DistributedPrioritizedBuffer
from .rnn_buffers import RNNBuffer, RNNPrioritizedBuffer
from .weight_tree import WeightTree

__all__ = ["Buffer", "EpisodeBook", "DistributedBuffer", "IMPALABuffer", "DeviceRing", "PrioritizedBuffer",
           "DistributedPrioritizedBuffer", "RNNBuffer", "RNNPrioritizedBuffer", "WeightTree"]
