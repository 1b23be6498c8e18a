"""Multi-GPU wiring of the engines — one process per GPU, ``torch.distributed`` (NCCL over NVLink; gloo on CPU for
the host-logic tests).  The data path shards by env column: no collective in the rollout; per
learner update

  * one all-reduce of the FLAT gradient buffer (``FlatAdam.grad``).  Reduction semantics follow the reference's
    global-batch loss: IMPALA / A2C losses are SUMS over the batch (parl/algorithms/paddle/impala/impala.py:67-79,
    parl/algorithms/torch/a2c.py:48-60) -> SUM; PPO / DQN / PG losses are MEANS (ppo.py:120-138, dqn.py:66) ->
    SUM then / world (equal shard sizes);
  * PPO ``norm_adv`` over a sharded minibatch: a (sum, sum of squares, n) all-reduce, then mean and the UNBIASED
    std exactly as ``(adv - adv.mean()) / (adv.std() + 1e-8)`` (parl/algorithms/torch/ppo.py:115-117);
  * prioritised replay sharded by rank: the importance weights divide by the GLOBAL minimum priority
    (benchmark/fluid/Prioritized_DQN/proportional_per.py:154-156) -> a 1-double all-reduce(MIN).
Launch: ``python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... script.py``.
"""
import os

import torch
import torch.distributed as dist

SUM_LOSS_ALGOS = ('IMPALA', 'A2C')            # losses summed over the batch
MEAN_LOSS_ALGOS = ('PPO', 'DQN', 'DDQN', 'PolicyGradient')


def init(backend=None):
    """Join the process group described by the torchrun environment; returns (rank, world, device)."""
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    device = torch.device('cuda', local_rank) if backend == 'nccl' else torch.device('cpu')
    if backend == 'nccl':
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if backend == 'nccl':
            dist.init_process_group(backend, device_id=device)
        else:
            dist.init_process_group(backend)
    return rank, world, device


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_initialized() else 1


def shard_envs(total_envs, rank, world):
    """Env columns of this rank: (num_envs, env_offset) — the offset keeps the GLOBAL Philox streams, so an N-GPU run
    steps exactly the envs a 1-GPU run of the same seed steps."""
    assert total_envs % world == 0, 'env columns must divide evenly over the ranks'
    n = total_envs // world
    return n, rank * n


def attach_grad_sync(alg, reduction=None, group=None):
    """alg.grad_sync = all-reduce of the flat gradient with the reference's loss semantics ('sum' | 'mean';
    default chosen from the algorithm class)."""
    if world_size(group) == 1:
        alg.grad_sync = None
        return alg
    if reduction is None:
        name = alg.__class__.__name__
        reduction = 'sum' if name in SUM_LOSS_ALGOS else 'mean'
    w = float(world_size(group))

    def sync(flat_grad):
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        if reduction == 'mean':
            flat_grad.div_(w)
    alg.grad_sync = sync
    return alg


def global_adv_stats(local_adv, group=None):
    """{mean, 1/(std_unbiased + 1e-8)} of the advantages of ALL ranks' minibatch shards (float32 [2] on the
    advantages' device) — the stats tensor rl_ppo_loss_fwd_bwd takes."""
    a = local_adv.double().reshape(-1)
    # torch.full (a fill kernel), not torch.tensor (a pageable H2D copy): the call may sit inside a captured CUDA graph
    mom = torch.stack([a.sum(), (a * a).sum(), torch.full((), float(a.numel()), dtype=torch.float64, device=a.device)])
    if world_size(group) > 1:
        dist.all_reduce(mom, op=dist.ReduceOp.SUM, group=group)
    n = mom[2]
    mean = mom[0] / n
    var = (mom[1] - n * mean * mean) / (n - 1.0)                      # unbiased, torch.Tensor.std default
    std = torch.sqrt(torch.clamp(var, min=0.0))
    return torch.stack([mean, 1.0 / (std + 1e-8)]).float()


def attach_adv_stats_sync(ppo_alg, group=None):
    ppo_alg.adv_stats_sync = (lambda adv: global_adv_stats(adv, group)) if world_size(group) > 1 else None
    return ppo_alg


def attach_per_min_sync(dqn_engine, group=None):
    """Sharded prioritised replay: make state[0] (the running minimum priority) global before each sample."""
    if world_size(group) == 1:
        dqn_engine.min_sync = None
        return dqn_engine

    def sync(state):
        dist.all_reduce(state[0:1], op=dist.ReduceOp.MIN, group=group)
    dqn_engine.min_sync = sync
    return dqn_engine


def replica_checksum(model):
    """One float64 per parameter tensor (sum of |w|) — cheap replica-consistency fingerprint."""
    return torch.stack([p.detach().double().abs().sum() for p in model.parameters()])


def check_replicas(model, group=None, rtol=0.0):
    """Assert that every rank holds the same weights (data-parallel learners apply identical updates to identical
    replicas; there is no parameter broadcast on the path).  Returns the max relative spread."""
    if world_size(group) == 1:
        return 0.0
    cs = replica_checksum(model)
    lo, hi = cs.clone(), cs.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    spread = ((hi - lo) / hi.clamp(min=1e-30)).max().item()
    if spread > rtol:
        raise RuntimeError('learner replicas diverged: max relative checksum spread %.3e' % spread)
    return spread
