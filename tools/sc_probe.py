import sys, torch
sys.path.insert(0, '.')
from parl_b200 import kernels as K, _lib
DEV = 'cuda:0'
for bo in (1, 0):
    _lib.load().rl_debug_set_shiftconv_base_offset(bo)
    for (N, H, Cin, Cout, k) in [(3, 21, 64, 32, 2), (5, 12, 128, 64, 2), (7, 11, 64, 64, 3)]:
        g = torch.Generator(device=DEV).manual_seed(1)
        x = torch.randn(N, H, H, Cin, device=DEV, generator=g).to(torch.bfloat16)
        w = (torch.randn(Cout, Cin, k, k, device=DEV, generator=g) / (Cin * k * k) ** 0.5).to(torch.bfloat16)
        b = torch.randn(Cout, device=DEV, generator=g)
        w_krsc = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous()
        out = K.conv2d_s1_nhwc_bf16_fwd(x, w_krsc, b, k, k, relu=True)
        torch.cuda.synchronize()
        ref = torch.relu(torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b)).permute(0, 2, 3, 1)
        print('base_offset', bo, (N, H, Cin, Cout, k), 'max err', (out.float() - ref).abs().max().item())
