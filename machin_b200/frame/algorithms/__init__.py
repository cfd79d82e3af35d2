from .a2c import A2C, A3C
from .apex import DDPGApex, DQNApex
from .ddpg_per import DDPGPer
from .dqn import DQN
from .dqn_per import DQNPer
from .impala import IMPALA
from .rainbow import RAINBOW

__all__ = ["DQN", "DQNPer", "DQNApex", "DDPGPer", "DDPGApex", "IMPALA", "A2C", "A3C", "RAINBOW"]
