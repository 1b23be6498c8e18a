"""Uncontended phase times of the IMPALA engine (sequential mode): rollout (graph replay of T actor steps) and learn,
for several per-GPU env counts — shows which side bounds the small-batch (8-GPU strong-scaling) regime."""
import json
import sys

import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from parl_b200.engine.impala import ImpalaEngine  # noqa: E402

dev = torch.device('cuda', 0)
for envs in [int(a) for a in sys.argv[1:]] or [512, 1024, 4096]:
    eng = ImpalaEngine(num_envs=envs, device=dev, pipeline=False)
    eng.reset()
    for _ in range(3):
        eng.rollout()
        eng.learn()
    res = {}
    for name, fn in (('rollout', eng.rollout), ('learn', eng.learn)):
        ev = []
        for _ in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            ev.append((a, b))
        torch.cuda.synchronize()
        ms = sorted(x.elapsed_time(y) for x, y in ev)
        res[name + '_ms'] = ms[len(ms) // 2]
    print(json.dumps(dict(envs=envs, T=eng.T, **res)))
    del eng
    torch.cuda.empty_cache()
