"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: env sharding by env_offset reproduces the
single-pool streams, and a SUM all-reduce of per-shard IMPALA gradients equals the global-batch gradient
(the reference's losses are SUMs over the batch: impala.py:67-79 — SURVEY.md §8e)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import envs as oenv
from oracle import vtrace as ovt


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    B, T, A, seed = 8, 6, 5, 3
    Bl = B // world
    # rank-local shard of the actor pool
    env = oenv.AtariSynthVec(Bl, seed, hw=64, p_done=0.3, env_offset=rank * Bl)
    env.reset()
    rews, dones = [], []
    for _ in range(T):
        _, r, d = env.step()
        rews.append(r), dones.append(d)
    rews, dones = np.stack(rews), np.stack(dones)
    # tiny "network": logits = W[a], values = w_v * x ; shared parameters on every rank
    g = torch.Generator().manual_seed(0)
    W = torch.randn(A, requires_grad=True, generator=g)
    wv = torch.randn(1, requires_grad=True, generator=g)
    rng = np.random.RandomState(100 + 0)
    feats_all = rng.randn(T, B).astype(np.float32)
    bl_all = rng.randn(T, B, A).astype(np.float32)
    acts_all = rng.randint(0, A, (T, B))
    sl = slice(rank * Bl, (rank + 1) * Bl)
    feats = torch.tensor(feats_all[:, sl])
    tl = feats[..., None] * W[None, None, :]
    vals = feats * wv
    o = ovt.impala_loss_time_major(tl.detach().numpy(), bl_all[:, sl], acts_all[:, sl], rews, dones, vals.detach().numpy(),
                                   0.99, 0.5, -0.01)
    torch.autograd.backward([tl, vals], [torch.tensor(o['d_logits']), torch.tensor(o['d_values'])])
    flat = torch.cat([W.grad.reshape(-1), wv.grad.reshape(-1)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    loss = torch.tensor([o['total_loss']])
    dist.all_reduce(loss, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put((flat.numpy(), loss.item(), rews, dones))
    else:
        out.put((None, None, rews, dones))
    dist.destroy_process_group()


def test_sharded_envs_and_sum_allreduce_match_single_process():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    flat = next(r[0] for r in res if r[0] is not None)
    loss = next(r[1] for r in res if r[1] is not None)
    # single-process global batch
    B, T, A, seed = 8, 6, 5, 3
    env = oenv.AtariSynthVec(B, seed, hw=64, p_done=0.3)
    env.reset()
    rews, dones = [], []
    for _ in range(T):
        _, r, d = env.step()
        rews.append(r), dones.append(d)
    rews, dones = np.stack(rews), np.stack(dones)
    shard_rews = sorted((r[2] for r in res), key=lambda x: -1)           # order-free check below
    got = np.concatenate([r[2] for r in sorted(res, key=lambda r: r[0] is None)], axis=1)
    assert np.array_equal(np.sort(got, axis=1), np.sort(rews, axis=1))   # same streams, sharded by env_offset
    g = torch.Generator().manual_seed(0)
    W = torch.randn(A, requires_grad=True, generator=g)
    wv = torch.randn(1, requires_grad=True, generator=g)
    rng = np.random.RandomState(100)
    feats = torch.tensor(rng.randn(T, B).astype(np.float32))
    bl = rng.randn(T, B, A).astype(np.float32)
    acts = rng.randint(0, A, (T, B))
    tl = feats[..., None] * W[None, None, :]
    vals = feats * wv
    o = ovt.impala_loss_time_major(tl.detach().numpy(), bl, acts, rews, dones, vals.detach().numpy(), 0.99, 0.5, -0.01)
    torch.autograd.backward([tl, vals], [torch.tensor(o['d_logits']), torch.tensor(o['d_values'])])
    want = torch.cat([W.grad.reshape(-1), wv.grad.reshape(-1)]).numpy()
    np.testing.assert_allclose(flat, want, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(loss, o['total_loss'], rtol=1e-5)
