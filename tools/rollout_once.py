"""One fused MLP rollout launch at the C4 shape with a short T (for ncu): python tools/rollout_once.py [T]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parl_b200.engine.ppo import PPOEngine  # noqa

T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
eng = PPOEngine(num_envs=2048, step_nums=T, num_minibatches=4, device=torch.device('cuda', 0), vec_normalize=True)
for _ in range(3):
    eng.rollout()
torch.cuda.synchronize()
print('done')
