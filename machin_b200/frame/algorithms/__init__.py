from .dqn_per import DQNPer

__all__ = ["DQNPer"]
