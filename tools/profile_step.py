"""Kernel-time breakdown of one bench step with torch.profiler (no ncu replay overhead).
    python tools/profile_step.py [envs]
"""
import os
import sys

import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parl_b200.engine.impala import ImpalaEngine  # noqa

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device('cuda', 0)
torch.manual_seed(0)
eng = ImpalaEngine(num_envs=B, sample_batch_steps=50, act_dim=18, seed=1, device=dev)
for _ in range(3):
    eng.rollout()
    eng.learn(1e-3, -0.01)
torch.cuda.synchronize()
for name, fn in (('rollout', eng.rollout), ('learn', lambda: eng.learn(1e-3, -0.01))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    print('%s: %.2f ms' % (name, e0.elapsed_time(e1)))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=18, max_name_column_width=70))
