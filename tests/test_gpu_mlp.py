"""GPU parity of the fused fp32 MLP kernels (rl_mlp_fwd / rl_mlp_bwd) against a plain PyTorch fp32 network of the
reference's model shapes (benchmark/torch/ppo/mujoco_model.py:27-53, benchmark/torch/QuickStart/cartpole_model.py:21-38,
examples/DQN/cartpole_model.py:21-41).  Tolerance: 1e-5 relative on outputs, 1e-4 on gradients (fp32 sums over up to
131 072 samples, different summation order)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _net(dims, heads, act):
    torch.manual_seed(sum(dims) + len(heads))
    layers = [nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)]
    hs = [nn.Linear(dims[-1], h) for h in heads]
    for m in layers + hs:
        m.cuda()
    f = torch.tanh if act == 1 else torch.relu

    def fwd(x):
        for m in layers:
            x = f(m(x))
        return torch.cat([h(x) for h in hs], 1)
    return layers, hs, fwd


@pytest.mark.parametrize('dims,heads,act,n', [
    ((17, 64, 64), (6, 1), 1, 2048),          # PPO MuJoCo actor-critic: mean + value heads
    ((17, 64, 64), (6, 1), 1, 131072),        # one C4 minibatch
    ((4, 20), (2, ), 1, 777),                 # QuickStart CartPole policy (ragged tile)
    ((4, 128, 128), (2, ), 0, 4096),          # DQN CartPole MLP
    ((4, 64), (2, 1), 1, 5120),               # A2C CartPole actor-critic (C2: 256 envs x 20 steps)
    ((9, ), (3, ), 2, 100),                   # a single linear layer
])
def test_mlp_fwd_bwd_matches_torch(dims, heads, act, n):
    from parl_b200 import kernels
    layers, hs, fwd = _net(dims, heads, act)
    x = torch.randn(n, dims[0], device='cuda')
    plan = kernels.MlpPlan([[(m.weight.detach(), m.bias.detach())] for m in layers] +
                           [[(h.weight.detach(), h.bias.detach()) for h in hs]], act)
    out = plan.forward(x)
    ref = fwd(x)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5), (out - ref).abs().max().item()
    d_out = torch.randn_like(ref) / n
    ref.backward(d_out)
    params = [m for m in layers] + [h for h in hs]
    grads = [(torch.empty_like(m.weight), torch.empty_like(m.bias)) for m in params]
    plan.backward(x, d_out, grads=grads)
    for m, (gw, gb) in zip(params, grads):
        sw = m.weight.grad.abs().max().item() + 1e-12
        assert (gw - m.weight.grad).abs().max().item() <= 1e-4 * sw + 1e-7, ('dw', tuple(m.weight.shape))
        sb = m.bias.grad.abs().max().item() + 1e-12
        assert (gb - m.bias.grad).abs().max().item() <= 1e-4 * sb + 1e-7, ('db', tuple(m.bias.shape))
    # accumulate=True adds on top; two identical passes are bit-identical (deterministic reduction)
    g2 = [(g[0].clone(), g[1].clone()) for g in grads]
    plan.backward(x, d_out, grads=g2, accumulate=True)
    for (gw, gb), (hw, hb) in zip(grads, g2):
        assert torch.allclose(hw, 2 * gw, rtol=1e-6, atol=1e-9) and torch.allclose(hb, 2 * gb, rtol=1e-6, atol=1e-9)
    g3 = [(torch.empty_like(g[0]), torch.empty_like(g[1])) for g in grads]
    plan.backward(x, d_out, grads=g3)
    for (gw, gb), (hw, hb) in zip(grads, g3):
        assert torch.equal(gw, hw) and torch.equal(gb, hb)


def test_mlp_rejects_bad_shapes():
    from parl_b200 import kernels
    w = torch.zeros(200, 4, device='cuda')
    plan = kernels.MlpPlan([[(w, None)]], 0)
    with pytest.raises(RuntimeError):
        plan.forward(torch.zeros(8, 4, device='cuda'))
