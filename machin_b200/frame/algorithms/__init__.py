from .apex import DQNApex
from .dqn_per import DQNPer

__all__ = ["DQNPer", "DQNApex"]
