"""Launch K1 a few times per kernel path (for ncu captures): python tools/k1_once.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parl_b200 import kernels as K, _lib  # noqa

T, A = 50, 18
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
modes = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 4]
dev = 'cuda:0'
tl = 2 * torch.randn(T * B, A, device=dev)
bl = tl + 0.5 * torch.randn(T * B, A, device=dev)
acts = torch.randint(0, A, (T * B,), device=dev, dtype=torch.int32)
rew = (torch.rand(T * B, device=dev) < 0.5).float()
dones = (torch.rand(T * B, device=dev) < 0.1).to(torch.uint8)
vals = torch.randn(T * B, device=dev)
out = dict(d_logits=torch.empty_like(tl), d_values=torch.empty_like(vals), losses=torch.empty(8, device=dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = _lib.load()
for mode in modes:
    lib.rl_debug_set_vtrace_path(mode)
    for i in range(3):
        flush.fill_(i)
        K.vtrace_loss_fwd_bwd(tl, bl, acts, rew, dones, vals, T, B, 0.99, 0.5, -0.01, out=out)
torch.cuda.synchronize()
print('done')
