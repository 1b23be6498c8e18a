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
