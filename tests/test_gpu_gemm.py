"""GPU parity: the tcgen05/TMA bf16 GEMM against a float32 torch matmul of the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (256, 512, 5184), (1000, 512, 576), (4096, 32, 256), (300, 19, 512),
                                   (128, 256, 128), (129, 130, 72)])
@pytest.mark.parametrize('relu', [False, True])
def test_gemm_bf16_tn(M, N, K, relu):
    from parl_b200 import kernels as K_
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    a = torch.randn(M, K, device=DEV, generator=g).to(torch.bfloat16)
    b = (torch.randn(N, K, device=DEV, generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV, generator=g)
    ref = a.float() @ b.float().t() + bias
    if relu:
        ref = torch.relu(ref)
    for dt, tol in ((torch.float32, 2e-3), (torch.bfloat16, 1.5e-2)):
        out = K_.gemm_bf16_tn(a, b, bias, relu=relu, out_dtype=dt)
        torch.cuda.synchronize()
        err = (out.float() - ref).abs().max().item()
        assert err < tol * max(1.0, ref.abs().max().item()), (dt, err)


@pytest.mark.parametrize('N,H,Cin,Cout,k,stride,pad', [(3, 21, 64, 32, 2, 1, 0), (5, 20, 32, 64, 4, 2, 2),
                                                        (7, 11, 64, 64, 3, 1, 0), (300, 11, 64, 64, 3, 1, 0),
                                                        (64, 20, 32, 64, 4, 2, 2)])
def test_conv2d_nhwc_bf16_fwd(N, H, Cin, Cout, k, stride, pad):
    """tcgen05 implicit-GEMM conv vs torch conv2d in float32 on the same bf16 operands (the three layer shapes of
    the Atari actor-critic: s2d conv1, conv2 with padding, conv3)."""
    from parl_b200 import kernels as K_
    g = torch.Generator(device=DEV).manual_seed(N + H + Cin)
    x = torch.randn(N, H, H, Cin, device=DEV, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, k, k, device=DEV, generator=g) / (Cin * k * k) ** 0.5).to(torch.bfloat16)
    b = torch.randn(Cout, device=DEV, generator=g)
    w_krsc = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).contiguous()
    out = K_.conv2d_nhwc_bf16_fwd(x, w_krsc, b, k, k, stride, pad, relu=True)
    torch.cuda.synchronize()
    ref = torch.relu(torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, stride=stride, padding=pad))
    ref = ref.permute(0, 2, 3, 1)
    err = (out.float() - ref).abs().max().item()
    assert out.shape == ref.shape and err < 2e-2 * max(1.0, ref.abs().max().item()), err
