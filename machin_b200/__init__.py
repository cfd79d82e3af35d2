"""machin_b200 -- B200-native learner hot path behind iffiX/machin's class surface.

Sub-packages mirror the reference's layout for the path only:
``machin_b200.frame.buffers``  (WeightTree, PrioritizedBuffer, ...),
``machin_b200.frame.algorithms`` (DQNPer / DQNApex / DDPGPer / DDPGApex / IMPALA / A3C value math),
``machin_b200.parallel`` (NCCL data plane replacing torch.rpc on this path).
"""
from . import _lib  # noqa: F401  (fails loudly if libmachin_b200.so is missing)

__version__ = "0.1.0"
