"""TMA-window conv tiles, per-tap form vs column-tap-fused form: forward of the three Atari layers and the two data
gradients at 51 200 samples, CUDA events (or `--once` for an ncu capture of the fused form)."""
import json
import sys

import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from parl_b200 import kernels as K  # noqa: E402

dev = torch.device('cuda', 0)
n = 51200
once = '--once' in sys.argv
bf = torch.bfloat16
torch.manual_seed(0)
r = lambda *s: (torch.randn(*s, device=dev) * 0.1).to(bf)
x0, a1, a2 = r(n, 21, 21, 64), r(n, 12, 12, 128), r(n, 11, 11, 64)
o1, o2, o3 = torch.zeros(n, 12, 12, 128, device=dev, dtype=bf), torch.empty(n, 11, 11, 64, device=dev, dtype=bf), \
    torch.empty(n, 9, 9, 64, device=dev, dtype=bf)
w1, w2, w3 = r(32, 256), r(64, 512), r(64, 576)
b1, b2, b3 = torch.zeros(32, device=dev), torch.zeros(64, device=dev), torch.zeros(64, device=dev)
w3T, w2T = r(64, 576), r(128, 256)
da3g, da2g, da1g = r(n, 11, 11, 64), torch.zeros(n, 12, 12, 64, device=dev, dtype=bf), torch.zeros(n, 21, 21, 32, device=dev, dtype=bf)
cases = [
    ('conv1_fwd', lambda: K.conv2d_s1_nhwc_bf16_fwd(x0, w1, b1, 2, 2, relu=True, out=o1, out_mode=1)),
    ('conv2_fwd', lambda: K.conv2d_s1_nhwc_bf16_fwd(a1, w2, b2, 2, 2, relu=True, out=o2)),
    ('conv3_fwd', lambda: K.conv2d_s1_nhwc_bf16_fwd(a2, w3, b3, 3, 3, relu=True, out=o3)),
    ('conv3_dgrad', lambda: K.conv2d_s1_nhwc_bf16_dgrad(da3g, w3T, 3, 3, da2g, act_mask=a2)),
    ('conv2_dgrad', lambda: K.conv2d_s1_nhwc_bf16_dgrad(da2g, w2T, 2, 2, da1g, act_mask=a1, out_mode=2)),
]
for form in ((1, ) if once else (0, 1)):
    K.set_shiftconv_form(form)
    for name, fn in cases:
        ev = []
        for _ in range(2 if once else 10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            ev.append((e0, e1))
        torch.cuda.synchronize()
        ms = sorted(p.elapsed_time(q) for p, q in ev[1:])
        print(json.dumps(dict(kernel=name, form='fused' if form else 'per_tap', samples=n, us=ms[len(ms) // 2] * 1e3)))
K.set_shiftconv_form(0)
