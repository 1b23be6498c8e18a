import sys, torch
sys.path.insert(0, '.')
from parl_b200 import kernels as K, _lib
DEV = 'cuda:0'
for lm in (0, 1):
    _lib.load().rl_debug_set_wgrad_lane_map(lm)
    for (N, H, Cin, Cout, k) in [(5, 11, 64, 64, 3), (9, 12, 128, 64, 2), (150, 11, 64, 64, 3), (7, 21, 64, 32, 2), (160, 21, 64, 32, 2)]:
        g = torch.Generator(device=DEV).manual_seed(1)
        Ho = H - k + 1
        x = torch.randn(N, H, H, Cin, device=DEV, generator=g).to(torch.bfloat16)
        dout = torch.randn(N, Ho, Ho, Cout, device=DEV, generator=g).to(torch.bfloat16)
        w = torch.zeros(Cout, Cin, k, k, device=DEV, requires_grad=True)
        torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w).backward(dout.float().permute(0, 3, 1, 2))
        ref = w.grad.permute(0, 2, 3, 1).reshape(Cout, -1)
        dgrid = torch.zeros(N, H, H, Cout, device=DEV, dtype=torch.bfloat16)
        dgrid[:, :Ho, :Ho] = dout
        dw = K.conv2d_s1_nhwc_bf16_wgrad(dgrid, x, k, k)
        torch.cuda.synchronize()
        print('lane_map', lm, (N, H, Cin, Cout, k), 'max err', (dw - ref).abs().max().item(), 'ref max', ref.abs().max().item())
    x = torch.randn(1000, 64, device=DEV).to(torch.bfloat16)
    print('colsum err', (K.colsum_bf16(x) - x.float().sum(0)).abs().max().item())
