"""conv1 forward / weight gradient on uint8 and bf16 observations at the round-1 capture size (51 200 samples):
timed with CUDA events, or launched a few times for an ncu capture (`--once`)."""
import json
import sys

import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from parl_b200 import kernels as K  # noqa: E402

dev = torch.device('cuda', 0)
n = 51200
once = '--once' in sys.argv
torch.manual_seed(0)
obs = torch.randint(0, 256, (2048, 4, 84, 84), dtype=torch.uint8, device=dev)
u8 = torch.empty((n, 21, 21, 64), dtype=torch.uint8, device=dev)
bf = torch.empty((n, 21, 21, 64), dtype=torch.bfloat16, device=dev)
for s in range(0, n, 2048):
    K.obs_stack_gather(obs, None, 0, 1, u8[s:s + 2048], s2d=True)
    K.obs_stack_gather(obs, None, 0, 1, bf[s:s + 2048], scale=1.0 / 255.0, s2d=True)
w = (torch.randn(32, 256, device=dev) * 0.05).to(torch.bfloat16)
b = torch.randn(32, device=dev) * 0.1
a1 = torch.zeros((n, 12, 12, 128), dtype=torch.bfloat16, device=dev)
dout = torch.zeros((n, 21, 21, 32), dtype=torch.bfloat16, device=dev)
dout[:, :20, :20] = (torch.randn(n, 20, 20, 32, device=dev) * 0.1).to(torch.bfloat16)
dw, db = torch.empty((32, 256), device=dev), torch.empty(32, device=dev)


def fwd(x):
    K.conv2d_s1_nhwc_bf16_fwd(x, w, b, 2, 2, relu=True, out=a1, out_mode=1)


def wg(x):
    K.conv2d_s1_nhwc_bf16_wgrad(dout, x, 2, 2, dw_krsc=dw, db=db)


def gather(x):
    K.obs_stack_gather(obs, None, 0, 1, x[:2048], scale=1.0 / 255.0, s2d=True)


cases = [('conv1_fwd', fwd), ('conv1_wgrad', wg), ('gather_2048', gather)]
for name, fn in cases:
    for tag, x in (('u8', u8), ('bf16', bf)):
        reps = 2 if once else 12
        ev = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(x)
            e1.record()
            ev.append((e0, e1))
        torch.cuda.synchronize()
        ms = sorted(p.elapsed_time(q) for p, q in ev[1:])
        m = n if name != 'gather_2048' else 2048
        in_b = 21 * 21 * 64 * (1 if tag == 'u8' else 2)
        moved = {'conv1_fwd': in_b + 25600, 'conv1_wgrad': in_b + 21 * 21 * 32 * 2, 'gather_2048': 28224 + in_b}[name] * m
        print(json.dumps(dict(kernel=name, input=tag, samples=m, us=ms[len(ms) // 2] * 1e3,
                              gbps_as_built=moved / (ms[len(ms) // 2] * 1e-3) / 1e9)))
