"""Small engine run for compute-sanitizer (memcheck / racecheck are 10-100x slower: keep it tiny).
    compute-sanitizer --tool memcheck python tools/sanitize_step.py [envs] [T]
Covers: eager rollout, eager learn, graph-captured rollout, pipelined step, host-contract step.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parl_b200.engine.impala import ImpalaEngine  # noqa

B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device('cuda', 0)
torch.manual_seed(0)
eng = ImpalaEngine(num_envs=B, sample_batch_steps=T, act_dim=18, seed=1, device=dev)
for _ in range(2):
    eng.rollout()
    out = eng.learn(1e-3, -0.01)
torch.cuda.synchronize()
print('sequential ok', [float(x) for x in out[:3]] if isinstance(out, (tuple, list)) else out)
eng = ImpalaEngine(num_envs=B, sample_batch_steps=T, act_dim=18, seed=3, device=dev, pipeline=True)
for _ in range(5):
    out = eng.step(1e-3, -0.01)
torch.cuda.synchronize()
print('pipelined ok')
eng = ImpalaEngine(num_envs=B, sample_batch_steps=T, act_dim=18, seed=2, device=dev, pipeline=True)
hosts = [eng.make_host_sample_buffers() for _ in range(2)]
for _ in range(3):
    eng.step_host(hosts, 1e-3, -0.01)
torch.cuda.synchronize()
print('host-contract ok')

# ---- the reference-facing host contract: remote Actor + AtariAgent.learn(numpy) (engine roles actor / learner)
import parl_b200 as parl  # noqa
from parl_b200.engine.impala_host import DeviceImpalaActor, AtariAgent  # noqa
parl.connect('localhost:8010')
cfg = dict(env_num=B, sample_batch_steps=T, act_dim=18, seed=4)
agent = AtariAgent(cfg, device=dev)
actor = parl.remote_class(wait=False)(DeviceImpalaActor)(cfg, device=dev)
actor.set_weights(agent.get_weights()).get()
fut = actor.sample()
for _ in range(2):
    batch = fut.get()
    actor.set_weights(agent.get_weights())
    fut = actor.sample()
    agent.learn(batch['obs'], batch['actions'], batch['behaviour_logits'], batch['rewards'], batch['dones'], 1e-3, -0.01)
fut.get()
actor.destroy()
torch.cuda.synchronize()
print('remote actor / agent ok')
