"""Whole-workload numbers for the BASELINE.json configs next to the IMPALA headline (bench.py):
    python tools/bench_workloads.py [a2c|ppo|dqn|all] [--gpus N via torchrun]
  C2  A2C, 256 vectorised CartPole envs x 20 steps per update                       (configs[1])
  C4  PPO, MuJoCo-shaped continuous control obs 17 / act 6, 2048 envs x 2048 steps,
      32 minibatches x 10 epochs                                                     (configs[3])
  C5  DQN, 1 M-transition HBM frame replay (sharded over the ranks), prioritised sample + TD loss   (configs[4])
One JSON line per workload (rank 0): env-steps/s over >= 3 timed iterations after warm-up, CUDA events on the
launching stream, max over ranks; per-phase times; the fp32 MLP kernels' share.  Under torchrun the env columns /
replay shard by rank with the collectives of parl_b200.engine.dist (gradient all-reduce, PPO advantage statistics,
PER minimum priority).
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parl_b200 import kernels as K           # noqa: E402
from parl_b200.engine import dist as pdist    # noqa: E402


def timed(fn, iters, warmup, dev, world):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    return ms.item() * 1e-3 / iters


def phase(fn, dev, n=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def bench_a2c(rank, world, dev, total_envs=256, T=20):
    from parl_b200.engine.a2c import A2CEngine
    B, off = pdist.shard_envs(total_envs, rank, world)
    torch.manual_seed(0)
    eng = A2CEngine(num_envs=B, sample_batch_steps=T, seed=1, device=dev, env_offset=off)
    pdist.attach_grad_sync(eng.alg)
    K.reset_launch_count()
    sec = timed(lambda: eng.step(0.001, -0.01), 200, 20, dev, world)
    launches = K.launch_count() / 220.0
    r_ms, l_ms = phase(eng.rollout, dev, 20), phase(lambda: eng.learn(0.001, -0.01), dev, 20)
    pdist.check_replicas(eng.model)
    return dict(workload='C2 A2C CartPole (configs[1])', metric='env_steps_per_sec', value=total_envs * T / sec,
                unit='env-steps/s', n_gpus=world, envs=total_envs, T=T, ms_per_iteration=sec * 1e3,
                rollout_ms=r_ms, learn_ms=l_ms, launches_per_iteration=launches,
                model='CartPoleActorCritic 4-64-64-{2,1} fp32 (rl_mlp_fwd/bwd, rl_rollout_mlp)',
                metrics=eng.get_metrics())


def bench_ppo(rank, world, dev, total_envs=2048, T=2048, epochs=10, minibatches=32, vec_normalize=True):
    from parl_b200.engine.ppo import PPOEngine
    B, off = pdist.shard_envs(total_envs, rank, world)
    torch.manual_seed(0)
    eng = PPOEngine(num_envs=B, step_nums=T, num_minibatches=minibatches, update_epochs=epochs, seed=2, device=dev,
                    env_offset=off, vec_normalize=vec_normalize)
    pdist.attach_grad_sync(eng.alg, reduction='sum')
    pdist.attach_adv_stats_sync(eng.alg)
    eng.grad_world = world                               # mean losses: SUM all-reduce then / world inside the Adam step
    K.reset_launch_count()
    sec = timed(lambda: eng.step(), 3, 1, dev, world)
    launches = K.launch_count() / 4.0
    r_ms = phase(eng.rollout, dev, 2)
    g_ms = phase(eng.compute_returns, dev, 3)
    l_ms = phase(lambda: eng.learn(), dev, 1)
    # kernel-level: one minibatch forward / backward of the fp32 MLP (11.3 kFLOP forward per sample)
    M = eng.M
    x = eng.obs.view(-1, eng.D)[:M].contiguous()
    f_ms = phase(lambda: eng.plan.forward(x, out=eng.mean_mb, out2=eng.val_mb, split=eng.AD), dev, 10)
    dm, dv = torch.randn_like(eng.mean_mb), torch.randn_like(eng.val_mb)
    b_ms = phase(lambda: eng.plan.backward(x, dm, d_out2=dv, split=eng.AD), dev, 10)
    flop_f = 2.0 * (17 * 64 + 64 * 64 + 64 * 7) * M
    pdist.check_replicas(eng.model)
    return dict(workload='C4 PPO MuJoCo-shaped (configs[3])', metric='env_steps_per_sec', value=total_envs * T / sec,
                unit='env-steps/s', n_gpus=world, envs=total_envs, T=T, minibatch=M * world, epochs=epochs,
                ms_per_iteration=sec * 1e3, rollout_ms=r_ms, gae_ms=g_ms, update_ms=l_ms,
                launches_per_iteration=launches, vec_normalize=vec_normalize,
                mlp_fwd_ms_per_minibatch=f_ms, mlp_bwd_ms_per_minibatch=b_ms,
                mlp_fwd_tflops=flop_f / (f_ms * 1e-3) / 1e12, mlp_bwd_tflops=3 * flop_f / (b_ms * 1e-3) / 1e12,
                gae_gbps=17.0 * T * B / (g_ms * 1e-3) / 1e9,
                model='MujocoModel 17-64-64-{6,1} fp32 (rl_mlp_fwd/bwd, rl_rollout_mlp)', metrics=eng.get_metrics())


def bench_dqn(rank, world, dev, memory_size=1000000, total_envs=1024, batch=4096, prioritized=True):
    from parl_b200.engine.dqn import DQNEngine
    B, off = pdist.shard_envs(total_envs, rank, world)
    torch.manual_seed(0)
    eng = DQNEngine(memory_size=memory_size // world, num_envs=B, batch_size=batch // world, act_dim=18,
                    prioritized=prioritized, double_q=True, seed=3, device=dev, env_offset=off, update_freq=4)
    pdist.attach_grad_sync(eng.alg)
    pdist.attach_per_min_sync(eng)
    eng.warmup(64)
    sec = timed(lambda: eng.step(), 20, 5, dev, world)
    e_ms, l_ms = phase(eng.env_step, dev, 8), phase(eng.learn, dev, 8)
    rows = eng.rpm.sample_uniform_rows(eng.batch_size)
    g_ms = phase(lambda: eng.rpm.gather(rows), dev, 10)
    out = {}
    s_ms = phase(lambda: out.__setitem__('s', eng.tree.sample(eng.batch_size, 0.5, float(eng.rpm.size()), seed=1, draw=0)),
                 dev, 10) if prioritized else None
    gather_bytes = eng.batch_size * 5 * 7056
    pdist.check_replicas(eng.model)
    return dict(workload='C5 DQN 1M-transition HBM replay + PER + TD (configs[4])', metric='learner_samples_per_sec',
                value=batch / (l_ms * 1e-3), unit='samples/s', n_gpus=world, replay_capacity=eng.rpm.max_size * world,
                frames_bytes_per_gpu=eng.rpm.frames.numel(), lanes_per_gpu=B, batch=batch, prioritized=prioritized,
                env_steps_per_sec=total_envs * eng.update_freq / sec, ms_per_iteration=sec * 1e3, env_step_ms=e_ms,
                learn_ms=l_ms, gather_ms=g_ms, gather_gbps=gather_bytes / (g_ms * 1e-3) / 1e9, per_sample_ms=s_ms,
                model='AtariQModel (benchmark/torch/dqn/model.py) torch bf16 autocast; replay / PER / TD loss native',
                metrics=eng.get_metrics())


if __name__ == '__main__':
    which = [a for a in sys.argv[1:] if not a.startswith('-')] or ['all']
    rank, world, dev = pdist.init('nccl')
    t0 = time.time()
    for name, fn in (('a2c', bench_a2c), ('ppo', bench_ppo), ('dqn', bench_dqn)):
        if name in which or 'all' in which:
            try:
                r = fn(rank, world, dev)
            except Exception as e:                        # a tool: report and carry on with the next workload
                import traceback
                r = dict(workload=name, error=repr(e), trace=traceback.format_exc()[-1500:])
            if rank == 0:
                print(json.dumps(r))
                sys.stdout.flush()
            torch.cuda.empty_cache()
    if world > 1:
        torch.distributed.destroy_process_group()
