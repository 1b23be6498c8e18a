"""Tiny launches of the non-tensor-core kernels for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_kernels.py
K1 (all three paths, both layouts), flat losses, GAE scans, PER, replay gathers, fused MLP forward / backward,
the fused rollout kernel with VecNormalize, the A2C / PPO / PG engines at toy sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parl_b200 import kernels as K, _lib  # noqa

dev = torch.device('cuda', 0)
torch.manual_seed(0)
lib = _lib.load()
T, B, A = 50, 16, 18
tl = torch.randn(T * B, A, device=dev)
bl = tl + 0.5 * torch.randn(T * B, A, device=dev)
acts = torch.randint(0, A, (T * B, ), device=dev, dtype=torch.int32)
rew = (torch.rand(T * B, device=dev) < 0.5).float()
dones = (torch.rand(T * B, device=dev) < 0.1).to(torch.uint8)
vals = torch.randn(T * B, device=dev)
for mode in (0, 4):
    lib.rl_debug_set_vtrace_path(mode)
    r = K.vtrace_loss_fwd_bwd(tl, bl, acts, rew, dones, vals, T, B, 0.99, 0.5, -0.01, want_returns=True)
lib.rl_debug_set_vtrace_path(0)
K.vtrace_loss_fwd_bwd(tl, bl, acts.long(), rew, dones, vals, T, B, 0.99, 0.5, -0.01, layout=K.ENV_MAJOR)
torch.cuda.synchronize()
print('K1 ok', r['losses'][:3].tolist())

N = 300
K.a2c_loss_fwd_bwd(torch.randn(N, 2, device=dev), torch.randn(N, device=dev),
                   torch.randint(0, 2, (N, ), device=dev, dtype=torch.int32), torch.randn(N, device=dev),
                   torch.randn(N, device=dev), 0.5, -0.01)
K.ppo_loss_fwd_bwd(torch.randn(N, device=dev), torch.randn(N, 6, device=dev), torch.randn(N, device=dev),
                   torch.randn(N, device=dev), torch.randn(N, device=dev), torch.randn(N, device=dev),
                   mean=torch.randn(N, 6, device=dev), logstd=torch.zeros(6, device=dev))
K.td_loss_fwd_bwd(torch.randn(N, 6, device=dev), torch.randn(N, 6, device=dev),
                  torch.randint(0, 6, (N, ), device=dev, dtype=torch.int32), torch.randn(N, device=dev),
                  torch.zeros(N, device=dev), 0.99, q_online_next=torch.randn(N, 6, device=dev))
K.gae_scan(torch.randn(20, 33, device=dev), torch.randn(20, 33, device=dev), torch.zeros(20, 33, device=dev),
           torch.randn(33, device=dev), torch.zeros(33, device=dev))
tree = K.DeviceSumTree(256, dev)
tree.store(0, 200, 0.6, 0.01)
ti, ei, w = tree.sample(32, 0.5, 200.0, seed=1, draw=0)
tree.update(ti, torch.rand(32, device=dev), 0.6, 0.01)
torch.cuda.synchronize()
print('losses / scans / PER ok')

from parl_b200.engine.a2c import A2CEngine      # noqa
from parl_b200.engine.ppo import PPOEngine      # noqa
from parl_b200.engine.pg import PolicyGradientEngine  # noqa
e = A2CEngine(num_envs=70, sample_batch_steps=6, device=dev)
e.step()
e = PPOEngine(num_envs=40, step_nums=8, num_minibatches=2, update_epochs=2, device=dev, vec_normalize=True,
              max_episode_steps=5, use_graph=False)
e.step()
e = PolicyGradientEngine(num_envs=40, rollout_steps=12, device=dev)
e.step()
torch.cuda.synchronize()
print('mlp engines ok')

# round-2 additions: conv1 on uint8 observations (forward + weight gradient), the fused env step + next-observation
# gather, the one-launch operand refresh, the 2x2-cluster GEMM, the fc + head call
obs = torch.randint(0, 256, (40, 4, 84, 84), dtype=torch.uint8, device=dev)
u8 = torch.empty((40, 21, 21, 64), dtype=torch.uint8, device=dev)
K.obs_stack_gather(obs, None, 0, 1, u8, s2d=True)
w1 = (torch.randn(32, 256, device=dev) * 0.05).to(torch.bfloat16)
a1 = torch.zeros((40, 12, 12, 128), dtype=torch.bfloat16, device=dev)
K.conv2d_s1_nhwc_bf16_fwd(u8, w1, torch.zeros(32, device=dev), 2, 2, relu=True, out=a1, out_mode=1)
dout = torch.zeros((40, 21, 21, 32), dtype=torch.bfloat16, device=dev)
dout[:, :20, :20] = 0.1
K.conv2d_s1_nhwc_bf16_wgrad(dout, u8, 2, 2, db=torch.empty(32, device=dev))
Bq, Tq = 37, 3
planes = torch.zeros((Tq + 4, Bq, 7056), dtype=torch.uint8, device=dev)
ages = torch.zeros((Tq + 1, Bq), dtype=torch.uint8, device=dev)
st = K.EpisodeStats(Bq, dev)
K.env_atari_synth_step(planes[3], None, None, None, ages[0], st, 5, 0, reset=True)
nxt = torch.empty((Bq, 21, 21, 64), dtype=torch.uint8, device=dev)
for t in range(Tq):
    K.env_atari_synth_step_gather(planes, t, torch.empty(Bq, device=dev), torch.empty(Bq, dtype=torch.uint8, device=dev),
                                  ages[t], ages[t + 1], st, 5, nxt, p_done=0.3, logits=torch.randn(Bq, 18, device=dev),
                                  actions_out=torch.empty(Bq, dtype=torch.int32, device=dev), step=t)
flat = torch.randn(5000, device=dev)
idx = torch.randint(-1, 5000, (4099, ), dtype=torch.int32, device=dev)
K.gather_cast(flat, idx, torch.empty(4099, dtype=torch.bfloat16, device=dev))
K.gather_cast(flat, idx, torch.empty(4099, dtype=torch.float32, device=dev))
a = (torch.randn(300, 512, device=dev) * 0.1).to(torch.bfloat16)
b = (torch.randn(512, 512, device=dev) * 0.1).to(torch.bfloat16)
K.gemm_bf16_tn(a, b, torch.zeros(512, device=dev), relu=True)                 # 3 x 4 tiles of 128 x 128: cluster form
K.gemm_bf16_tn_heads(a, b, torch.zeros(512, device=dev), torch.empty(300, 512, dtype=torch.bfloat16, device=dev),
                     (torch.randn(18, 512, device=dev) * 0.1).to(torch.bfloat16), torch.zeros(18, device=dev),
                     torch.empty(300, 18, device=dev))
torch.cuda.synchronize()
print('round-2 kernels ok')
