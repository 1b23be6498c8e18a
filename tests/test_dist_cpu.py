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


# ---------------------------------------------------------------------------------------------------------------
# product code: parl_b200.engine.dist (the multi-GPU wiring of the engines) on the gloo backend
def _dist_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from parl_b200.engine import dist as pdist
    r, w, dev = pdist.init('gloo')
    assert (r, w, dev.type) == (rank, world, 'cpu')
    n, off = pdist.shard_envs(4096, rank, world)
    res = dict(shard=(n, off))

    class Alg(object):                      # the attribute surface the helpers touch
        grad_sync = None
        adv_stats_sync = None

    class IMPALA(Alg):
        pass

    class PPO(Alg):
        pass

    g = torch.arange(6, dtype=torch.float32) * (rank + 1)
    a = pdist.attach_grad_sync(IMPALA())                    # SUM-loss algorithm -> SUM all-reduce
    x = g.clone()
    a.grad_sync(x)
    res['sum'] = x.numpy()
    p = pdist.attach_grad_sync(PPO())                       # MEAN-loss algorithm -> SUM / world
    y = g.clone()
    p.grad_sync(y)
    res['mean'] = y.numpy()
    # global advantage statistics over a sharded minibatch
    rng = np.random.RandomState(7)
    adv_all = rng.randn(world * 50).astype(np.float32) * 3 + 1
    pdist.attach_adv_stats_sync(p)
    st = p.adv_stats_sync(torch.tensor(adv_all[rank * 50:(rank + 1) * 50]))
    res['stats'] = st.numpy()
    # sharded PER: global minimum priority

    class Eng(object):
        min_sync = None
    e = pdist.attach_per_min_sync(Eng())
    state = torch.tensor([0.5 + rank, 2.0], dtype=torch.float64)
    e.min_sync(state)
    res['per_state'] = state.numpy()
    # replica fingerprints
    m = torch.nn.Linear(3, 2)
    with torch.no_grad():
        m.weight.fill_(0.25), m.bias.fill_(1.0)
    res['spread_same'] = pdist.check_replicas(m)
    with torch.no_grad():
        m.bias.add_(float(rank))
    try:
        pdist.check_replicas(m)
        res['diverged_detected'] = False
    except RuntimeError:
        res['diverged_detected'] = True
    out.put((rank, res))
    dist.destroy_process_group()


def test_engine_dist_helpers_world2():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dist_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert res[0]['shard'] == (2048, 0) and res[1]['shard'] == (2048, 2048)
    base = np.arange(6, dtype=np.float32)
    for r in range(world):
        np.testing.assert_allclose(res[r]['sum'], base * 3)            # (1 + 2) x
        np.testing.assert_allclose(res[r]['mean'], base * 1.5)
        adv_all = np.random.RandomState(7).randn(world * 50).astype(np.float32) * 3 + 1
        t = torch.tensor(adv_all)
        want = np.array([t.mean().item(), 1.0 / (t.std().item() + 1e-8)])      # ppo.py:115-117 (unbiased std)
        np.testing.assert_allclose(res[r]['stats'], want, rtol=1e-5)
        np.testing.assert_allclose(res[r]['per_state'], [0.5, 2.0])
        assert res[r]['spread_same'] == 0.0 and res[r]['diverged_detected']
