"""Micro-benchmarks of individual kernels (CUDA events, rotating >L2 buffers).
    python tools/bench_kernels.py [vtrace|env|all]
"""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parl_b200 import kernels as K  # noqa


def time_fn(fn, iters=50, warmup=5):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for i in range(iters):
        fn(i)
    en.record()
    torch.cuda.synchronize()
    return st.elapsed_time(en) / iters * 1e-3


def bench_vtrace(T=50, B=4096, A=18, nbuf=8, iters=64):
    dev = 'cuda:0'
    bufs = []
    for i in range(nbuf):
        g = torch.Generator(device=dev).manual_seed(i)
        tl = 2 * torch.randn(T * B, A, device=dev, generator=g)
        bl = tl + 0.5 * torch.randn(T * B, A, device=dev, generator=g)
        acts = torch.randint(0, A, (T * B,), device=dev, dtype=torch.int32, generator=g)
        rew = (torch.rand(T * B, device=dev, generator=g) < 0.5).float()
        dones = (torch.rand(T * B, device=dev, generator=g) < 0.1).to(torch.uint8)
        vals = torch.randn(T * B, device=dev, generator=g)
        out = dict(d_logits=torch.empty_like(tl), d_values=torch.empty_like(vals),
                   losses=torch.empty(8, device=dev))
        bufs.append((tl, bl, acts, rew, dones, vals, out))

    def run(i):
        tl, bl, acts, rew, dones, vals, out = bufs[i % nbuf]
        K.vtrace_loss_fwd_bwd(tl, bl, acts, rew, dones, vals, T, B, 0.99, 0.5, -0.01, out=out)
    sec = time_fn(run, iters=iters)
    alg_bytes = (T - 1) * B * (12 * A + 17) + 4 * B
    return dict(kernel='vtrace_loss_fwd_bwd', T=T, B=B, A=A, us=sec * 1e6, alg_bytes=alg_bytes,
                gbps=alg_bytes / sec / 1e9, footprint_mb=nbuf * alg_bytes / 1e6)


def bench_env(B=4096, HW=84 * 84, nplanes=16, iters=64):
    dev = 'cuda:0'
    planes = torch.zeros(nplanes, B, HW, dtype=torch.uint8, device=dev)
    rew = torch.zeros(B, device=dev)
    done = torch.zeros(B, dtype=torch.uint8, device=dev)
    age = torch.zeros(B, dtype=torch.uint8, device=dev)
    st = K.EpisodeStats(B, dev)
    logits = torch.randn(B, 18, device=dev)
    acts = torch.zeros(B, dtype=torch.int32, device=dev)

    def run(i):
        K.env_atari_synth_step(planes[i % nplanes], rew, done, age, age, st, 1, i, logits=logits, actions_out=acts)
    sec = time_fn(run, iters=iters)
    alg = B * HW
    return dict(kernel='env_atari_synth_step', B=B, us=sec * 1e6, alg_bytes=alg, gbps=alg / sec / 1e9,
                env_steps_per_s=B / sec)


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if which == 'vtrace_cpasync':
        from parl_b200 import _lib
        _lib.load().rl_debug_set_tma(1)
        r = bench_vtrace()
        r['path'] = 'cp.async'
        print(json.dumps(r))
    if which == 'vtrace_c3':
        print(json.dumps(bench_vtrace()))
    if which in ('vtrace', 'all'):
        print(json.dumps(bench_vtrace()))
        print(json.dumps(bench_vtrace(B=65536, nbuf=2, iters=16)))
        print(json.dumps(bench_vtrace(B=512, nbuf=64)))
    if which in ('env', 'all'):
        print(json.dumps(bench_env()))
        print(json.dumps(bench_env(B=512)))
