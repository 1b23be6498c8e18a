import sys, time, torch
sys.path.insert(0, '.')
from parl_b200 import kernels as K
from parl_b200.engine.nets import AtariActorCritic
from parl_b200.engine.train_net import AtariTrainNet
DEV = 'cuda:0'
rel = lambda a, b: ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()
torch.manual_seed(0)
N, A = 160, 18
t0 = time.time()
model = AtariActorCritic(A).to(DEV)
with torch.no_grad():
    for p in model.parameters():
        if p.dim() == 1:
            p.normal_(0, 0.1)
for p in model.parameters():
    p.grad = torch.zeros_like(p)
net = AtariTrainNet(model, N, DEV)
torch.cuda.synchronize(); print('init %.2fs' % (time.time() - t0)); t0 = time.time()
obs = torch.randint(0, 255, (N, 4, 84, 84), dtype=torch.uint8, device=DEV)
K.obs_stack_gather(obs, None, 0, 1, net.x0, scale=1.0 / 255.0, s2d=True)
logits, values = net.forward_from_x0()
torch.cuda.synchronize(); print('fwd %.2fs' % (time.time() - t0)); t0 = time.time()
d_logits = torch.randn(N, A, device=DEV) * 0.1
d_values = torch.randn(N, device=DEV) * 0.1
net.backward(d_logits, d_values)
torch.cuda.synchronize(); print('bwd %.2fs' % (time.time() - t0)); t0 = time.time()
got = {n: p.grad.clone() for n, p in model.named_parameters()}
for p in model.parameters():
    p.grad = None
rl, rv = model.policy_and_value(net.x0)
torch.autograd.backward([rl, rv], [d_logits, d_values])
torch.cuda.synchronize(); print('ref %.2fs' % (time.time() - t0))
print('logits', rel(logits, rl), 'values', rel(values.view(-1), rv))
for n, p in model.named_parameters():
    print('%-16s rel %.4f  |g| %.4g |ref| %.4g' % (n, rel(got[n], p.grad), got[n].norm().item(), p.grad.norm().item()))
ref_bf16 = {n: p.grad.clone() for n, p in model.named_parameters()}
m32 = AtariActorCritic(A, compute_dtype=torch.float32).to(DEV)
m32.load_state_dict(model.state_dict())
l32, v32 = m32.policy_and_value(obs)
torch.autograd.backward([l32, v32], [d_logits, d_values])
print('vs float32 reference-form network:   ours      torch-bf16')
for (n, p), (_, q) in zip(model.named_parameters(), m32.named_parameters()):
    print('%-16s %.4f    %.4f' % (n, rel(got[n], q.grad), rel(ref_bf16[n], q.grad)))
