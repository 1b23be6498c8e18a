"""Two launches of every K2-K4 / replay / MLP / rollout kernel at the BASELINE config shapes (for ONE ncu capture that
covers them all): python tools/kernels_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import bench_kernels as BK  # noqa

def _two(fn, iters=2, warmup=1):
    fn(0)
    fn(1)
    torch.cuda.synchronize()
    return 1e-3


BK.time_fn = _two
for fn in (BK.bench_a2c, BK.bench_gae, lambda: BK.bench_gae(T=20, B=256, segments=True), BK.bench_ppo, BK.bench_td,
           BK.bench_per, BK.bench_replay_gather, BK.bench_mlp):
    fn()
torch.cuda.synchronize()
print('done')
