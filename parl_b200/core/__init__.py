from .model import Model
from .algorithm import Algorithm
from .agent import Agent

__all__ = ['Model', 'Algorithm', 'Agent']
