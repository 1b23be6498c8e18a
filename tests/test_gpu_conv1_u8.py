"""conv1 on the uint8 observation (rl_obs_stack_gather out_dtype 4, rl_conv2d_s1_u8in_bf16_{fwd,wgrad}).

The uint8 path widens the bytes to bf16(byte/255) in shared memory with the same arithmetic as the bf16 gather, so
every result must be BIT-IDENTICAL to the bf16-input kernels; the layout itself is checked against a plain tensor
restatement of the space-to-depth transform of the reference's first conv (8x8 / stride 4 / pad 1 on 4x84x84,
benchmark/torch/a2c/atari_model.py:26-27)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from parl_b200 import kernels as K
    from parl_b200.engine.nets import AtariActorCritic
    from parl_b200.engine.train_net import AtariTrainNet
    DEV = torch.device('cuda', 0)


def _s2d_reference(obs):
    """obs [n,4,84,84] uint8 -> [n,21,21,64] uint8, channel (dy*4+dx)*4+c = pixel (4Y+dy-1, 4X+dx-1), 0 outside."""
    n = obs.shape[0]
    pad = torch.zeros((n, 4, 85, 85), dtype=torch.uint8, device=obs.device)
    pad[:, :, 1:, 1:] = obs
    blk = pad[:, :, :84, :84].reshape(n, 4, 21, 4, 21, 4)              # (n, c, Y, dy, X, dx)
    return blk.permute(0, 2, 4, 3, 5, 1).reshape(n, 21, 21, 64).contiguous()


@pytest.mark.parametrize('n', [1, 37])
def test_gather_u8_space_to_depth_layout(n):
    torch.manual_seed(n)
    obs = torch.randint(0, 256, (n, 4, 84, 84), dtype=torch.uint8, device=DEV)
    out = torch.empty((n, 21, 21, 64), dtype=torch.uint8, device=DEV)
    K.obs_stack_gather(obs, None, 0, 1, out, s2d=True)
    assert torch.equal(out, _s2d_reference(obs))
    bf = torch.empty((n, 21, 21, 64), dtype=torch.bfloat16, device=DEV)
    K.obs_stack_gather(obs, None, 0, 1, bf, scale=1.0 / 255.0, s2d=True)
    assert torch.equal((out.float() * (1.0 / 255.0)).to(torch.bfloat16), bf)


def test_gather_u8_from_frame_ring_matches_bf16_gather():
    torch.manual_seed(3)
    T, B = 6, 19
    planes = torch.randint(0, 256, (T + 4, B, 84 * 84), dtype=torch.uint8, device=DEV)
    ages = torch.randint(0, 5, (T + 1, B), dtype=torch.uint8, device=DEV)
    for layout in (K.TIME_MAJOR, K.ENV_MAJOR):
        u8 = torch.empty((T * B, 21, 21, 64), dtype=torch.uint8, device=DEV)
        bf = torch.empty((T * B, 21, 21, 64), dtype=torch.bfloat16, device=DEV)
        K.obs_stack_gather(planes, ages, 0, T, u8, layout=layout, s2d=True)
        K.obs_stack_gather(planes, ages, 0, T, bf, layout=layout, scale=1.0 / 255.0, s2d=True)
        assert torch.equal((u8.float() * (1.0 / 255.0)).to(torch.bfloat16), bf)


def _operands(n, seed):
    torch.manual_seed(seed)
    obs = torch.randint(0, 256, (n, 4, 84, 84), dtype=torch.uint8, device=DEV)
    u8 = torch.empty((n, 21, 21, 64), dtype=torch.uint8, device=DEV)
    bf = torch.empty((n, 21, 21, 64), dtype=torch.bfloat16, device=DEV)
    K.obs_stack_gather(obs, None, 0, 1, u8, s2d=True)
    K.obs_stack_gather(obs, None, 0, 1, bf, scale=1.0 / 255.0, s2d=True)
    w = (torch.randn(32, 256, device=DEV) * 0.05).to(torch.bfloat16)
    b = torch.randn(32, device=DEV) * 0.1
    return u8, bf, w, b


# 1 sample: fewer tiles than SMs; 160: ragged last tile, 3-4 tiles per CTA (staging ring wraps); 512: 12 tiles per
# CTA (the 8-deep window ring wraps too)
@pytest.mark.parametrize('n', [1, 160, 512])
@pytest.mark.parametrize('out_mode', [0, 1])
def test_conv1_forward_u8_bit_identical_to_bf16_input(n, out_mode):
    u8, bf, w, b = _operands(n, 10 + n)
    shape = (n, 20, 20, 32) if out_mode == 0 else (n, 12, 12, 128)
    o_u8 = torch.zeros(shape, dtype=torch.bfloat16, device=DEV)
    o_bf = torch.zeros(shape, dtype=torch.bfloat16, device=DEV)
    K.conv2d_s1_nhwc_bf16_fwd(u8, w, b, 2, 2, relu=True, out=o_u8, out_mode=out_mode)
    K.conv2d_s1_nhwc_bf16_fwd(bf, w, b, 2, 2, relu=True, out=o_bf, out_mode=out_mode)
    torch.cuda.synchronize()
    assert o_bf.float().abs().sum().item() > 0
    assert torch.equal(o_u8, o_bf)


def test_conv1_forward_u8_against_float32_conv():
    n = 64
    torch.manual_seed(5)
    obs = torch.randint(0, 256, (n, 4, 84, 84), dtype=torch.uint8, device=DEV)
    u8 = torch.empty((n, 21, 21, 64), dtype=torch.uint8, device=DEV)
    K.obs_stack_gather(obs, None, 0, 1, u8, s2d=True)
    w4 = torch.randn(32, 4, 8, 8, device=DEV) * 0.05
    b = torch.randn(32, device=DEV) * 0.1
    w = w4.view(32, 4, 2, 4, 2, 4).permute(0, 2, 4, 3, 5, 1).reshape(32, 256).to(torch.bfloat16)
    out = K.conv2d_s1_nhwc_bf16_fwd(u8, w, b, 2, 2, relu=True)
    x = (obs.float() * (1.0 / 255.0)).to(torch.bfloat16).float()
    ref = torch.relu(torch.nn.functional.conv2d(x, w4.to(torch.bfloat16).float(), b, stride=4, padding=1))
    err = (out.float().permute(0, 3, 1, 2) - ref).abs().max().item()
    assert err < 0.03, err          # bf16 output rounding of O(1) activations


@pytest.mark.parametrize('n', [1, 160, 512])
def test_conv1_wgrad_u8_bit_identical_to_bf16_input(n):
    u8, bf, _, _ = _operands(n, 20 + n)
    dout = torch.zeros((n, 21, 21, 32), dtype=torch.bfloat16, device=DEV)
    dout[:, :20, :20] = (torch.randn(n, 20, 20, 32, device=DEV) * 0.1).to(torch.bfloat16)
    dw_u8, db_u8 = torch.empty((32, 256), device=DEV), torch.empty(32, device=DEV)
    dw_bf, db_bf = torch.empty((32, 256), device=DEV), torch.empty(32, device=DEV)
    K.conv2d_s1_nhwc_bf16_wgrad(dout, u8, 2, 2, dw_krsc=dw_u8, db=db_u8)
    K.conv2d_s1_nhwc_bf16_wgrad(dout, bf, 2, 2, dw_krsc=dw_bf, db=db_bf)
    torch.cuda.synchronize()
    assert dw_bf.abs().sum().item() > 0
    assert torch.equal(dw_u8, dw_bf) and torch.equal(db_u8, db_bf)
    # accumulate form
    K.conv2d_s1_nhwc_bf16_wgrad(dout, u8, 2, 2, dw_krsc=dw_u8, db=db_u8, accumulate=True)
    assert torch.allclose(dw_u8, 2 * dw_bf, rtol=1e-6, atol=0) and torch.allclose(db_u8, 2 * db_bf, rtol=1e-6, atol=0)


def test_train_net_uint8_observations_give_identical_gradients():
    N, A = 160, 18
    grads = []
    obs = torch.randint(0, 256, (N, 4, 84, 84), dtype=torch.uint8, device=DEV, generator=torch.Generator(DEV).manual_seed(1))
    d_logits = torch.randn(N, A, device=DEV, generator=torch.Generator(DEV).manual_seed(2)) * 0.1
    d_values = torch.randn(N, device=DEV, generator=torch.Generator(DEV).manual_seed(3)) * 0.1
    for dt in (torch.uint8, torch.bfloat16):
        torch.manual_seed(0)
        model = AtariActorCritic(A).to(DEV)
        for p in model.parameters():
            p.grad = torch.zeros_like(p)
        net = AtariTrainNet(model, N, DEV, obs_dtype=dt)
        assert net.x0.dtype == dt
        K.obs_stack_gather(obs, None, 0, 1, net.x0, scale=1.0 / 255.0, s2d=True)
        logits, values = net.forward_from_x0()
        net.backward(d_logits, d_values)
        torch.cuda.synchronize()
        grads.append((logits.clone(), values.clone(), [p.grad.clone() for p in model.parameters()]))
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])
    for a, b in zip(grads[0][2], grads[1][2]):
        assert torch.equal(a, b)
