"""A few PPO minibatch updates without graph capture (for an ncu launch list): python tools/ppo_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parl_b200.engine.ppo import PPOEngine  # noqa

dev = torch.device('cuda', 0)
eng = PPOEngine(num_envs=2048, step_nums=2048, device=dev, vec_normalize=True, use_graph=False)
eng.rollout()
eng.compute_returns()
perm = torch.randperm(eng.N, device=dev, dtype=torch.int32)
for mb in range(4):
    eng.learn_minibatch(perm[mb * eng.M:(mb + 1) * eng.M], 3e-4)
torch.cuda.synchronize()
print('done')
