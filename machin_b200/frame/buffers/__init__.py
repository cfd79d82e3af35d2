from .buffer import Buffer
from .device_ring import DeviceRing
from .prioritized_buffer import PrioritizedBuffer
from .weight_tree import WeightTree

__all__ = ["Buffer", "DeviceRing", "PrioritizedBuffer", "WeightTree"]
