"""DDQN — parl/algorithms/torch/ddqn.py:29-78: greedy action from the online network, value from
the target network; everything else as DQN."""
from .dqn import DQN

__all__ = ['DDQN']


class DDQN(DQN):
    double_q = True
