"""``parl.algorithms`` for the hot path: IMPALA, A2C, PPO, DQN, DDQN, PolicyGradient, and the continuous-control
family DDPG / TD3 / SAC.

Signatures follow the reference; the network forward/backward goes through the
user's ``parl.Model`` (torch autograd), everything after the network outputs — returns scan,
losses, gradient w.r.t. the outputs, clipping, Adam — runs in libparl_b200.so."""
from .impala import IMPALA
from .a2c import A2C
from .ppo import PPO
from .dqn import DQN
from .ddqn import DDQN
from .policy_gradient import PolicyGradient
from .ddpg import DDPG
from .td3 import TD3
from .sac import SAC

__all__ = ['IMPALA', 'A2C', 'PPO', 'DQN', 'DDQN', 'PolicyGradient', 'DDPG', 'TD3', 'SAC']
