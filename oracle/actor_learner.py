"""CPU restatement of the reference IMPALA actor-learner loop (TEST INFRASTRUCTURE / CPU baseline).

Follows examples/IMPALA/train.py:34-252 (Learner: sample queue, train batches, Adam + global-norm
clip 40), examples/IMPALA/actor.py:27-105 (Actor.sample: 50 steps x 5 envs, env-major merge),
examples/IMPALA/atari_agent.py:35-42 (batch-5 CPU inference + np.random.choice per row),
parl/env/vector_env.py:41-63 (auto-reset) and parl/env/atari_wrappers.py:270-307 (FrameStack).
The env is the "lean" flavour of SURVEY.md Appendix C: a direct 84x84 uint8 synthetic frame env
(parl/tests/gym.py:163-169 distributions) behind FrameStack(4) — the most favourable case for the
CPU side.  Actors are OS processes (one per core, single-threaded torch, as the reference's xparl
jobs are: parl/core/torch/agent.py:26, parl/remote/job.py:17) that ship their sample dict back by
pickle over a pipe (standing in for cloudpickle+ZeroMQ, parl/remote/communication.py:59-130); the
learner uses the remaining cores.  paddle is absent, so the network is the torch twin of the C3 model.
Used only by bench.py (cpu_baseline / --impl reference) and tests.
"""
import multiprocessing as mp
import os
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import vtrace as ovt


class LeanAtariEnv(object):
    """84x84 uint8 frames ~ U{0..254}, reward {0,1}, done p=.1, FrameStack(4, 'NCHW')."""

    def __init__(self, hw=(84, 84), p_done=0.1):
        self.hw, self.p_done = hw, p_done
        self.frames = None

    def _frame(self):
        return np.random.randint(0, 255, self.hw, dtype=np.uint8)

    def reset(self):
        f = self._frame()
        self.frames = [f] * 4
        return np.stack(self.frames)

    def step(self, action):
        f = self._frame()
        reward = float(np.random.choice([0.0, 1.0]))
        done = bool(np.random.choice([True, False], p=[self.p_done, 1 - self.p_done]))
        self.frames = self.frames[1:] + [f]
        return np.stack(self.frames), reward, done, {}


class CpuAtariModel(nn.Module):
    """torch twin of benchmark/torch/a2c/atari_model.py:23-96 (the 84x84 actor-critic)."""

    def __init__(self, act_dim):
        super().__init__()
        self.conv1 = nn.Conv2d(4, 32, 8, 4, 1)
        self.conv2 = nn.Conv2d(32, 64, 4, 2, 2)
        self.conv3 = nn.Conv2d(64, 64, 3, 1, 0)
        self.fc = nn.Linear(64 * 9 * 9, 512)
        self.fc_pi = nn.Linear(512, act_dim)
        self.fc_v = nn.Linear(512, 1)

    def trunk(self, x):
        x = x / 255.0
        x = F.relu(self.conv1(x))
        x = F.relu(self.conv2(x))
        x = F.relu(self.conv3(x))
        return F.relu(self.fc(x.flatten(1)))

    def policy(self, x):
        return self.fc_pi(self.trunk(x))

    def value(self, x):
        return self.fc_v(self.trunk(x)).squeeze(1)


def actor_sample(model, envs, obs_batch, T):
    """Actor.sample (examples/IMPALA/actor.py:54-91)."""
    n = len(envs)
    data = [dict(obs=[], actions=[], behaviour_logits=[], rewards=[], dones=[]) for _ in range(n)]
    for _ in range(T):
        with torch.no_grad():
            logits = model.policy(torch.from_numpy(np.stack(obs_batch)).float())
            probs = F.softmax(logits, -1).numpy().astype(np.float64)
        probs /= probs.sum(-1, keepdims=True)
        actions = np.array([np.random.choice(len(p), 1, p=p)[0] for p in probs])     # atari_agent.py:39-40
        nxt = []
        for e in range(n):                                                            # vector_env.py:53-63
            o, r, d, _ = envs[e].step(actions[e])
            if d:
                o = envs[e].reset()
            nxt.append(o)
            data[e]['obs'].append(obs_batch[e])
            data[e]['actions'].append(actions[e])
            data[e]['behaviour_logits'].append(logits[e].numpy())
            data[e]['rewards'].append(r)
            data[e]['dones'].append(d)
        obs_batch = nxt
    out = {k: np.stack([x for e in range(n) for x in data[e][k]]) for k in data[0]}   # env-major merge :79-89
    return out, obs_batch


def _actor_proc(conn, act_dim, env_num, T, seed):
    torch.set_num_threads(1)
    np.random.seed(seed)
    torch.manual_seed(seed)
    model = CpuAtariModel(act_dim)
    envs = [LeanAtariEnv() for _ in range(env_num)]
    obs = [e.reset() for e in envs]
    while True:
        msg = conn.recv()
        if msg is None:
            return
        model.load_state_dict({k: torch.from_numpy(v) for k, v in msg.items()})       # set_weights (train.py:171)
        sample, obs = actor_sample(model, envs, obs, T)
        conn.send(sample)                                                             # pickle over a pipe


def impala_learn(model, optimizer, batch, T, gamma=0.99, vf_coeff=0.5, ent_coeff=-0.01, lr=1e-3):
    """IMPALA.learn (parl/algorithms/paddle/impala/impala.py:134-215) in torch CPU float32."""
    obs = torch.from_numpy(batch['obs']).float()
    actions = torch.from_numpy(batch['actions'].astype(np.int64))
    bl = torch.from_numpy(batch['behaviour_logits'].astype(np.float32))
    values = model.value(obs)                                                         # two passes, as the reference
    tl = model.policy(obs)
    A = tl.shape[-1]
    t_lsm, b_lsm = F.log_softmax(tl, -1), F.log_softmax(bl, -1)
    onehot = F.one_hot(actions, A).float()
    tlp, blp = (t_lsm * onehot).sum(-1), (b_lsm * onehot).sum(-1)
    p = t_lsm.exp()
    entropy = -(p * t_lsm).sum(-1)
    kl = (p * (t_lsm - b_lsm)).sum(-1).mean()
    B = obs.shape[0] // T

    def tm(x):
        return x.reshape(B, T).transpose(0, 1)
    tlp_, blp_, ent_, v_ = tm(tlp), tm(blp), tm(entropy), tm(values)
    rew = tm(torch.from_numpy(batch['rewards'].astype(np.float32)))
    dones = tm(torch.from_numpy(batch['dones'].astype(bool)))
    boot = v_[-1]
    disc = (~dones[:-1]).float() * gamma
    vs, pg = ovt.from_importance_weights(blp_[:-1].detach().numpy(), tlp_[:-1].detach().numpy(), disc.numpy(),
                                         rew[:-1].numpy(), v_[:-1].detach().numpy(), boot.detach().numpy(), 1.0, 1.0)
    pi_loss = -(tlp_[:-1] * torch.from_numpy(pg)).sum()
    vf_loss = 0.5 * ((v_[:-1] - torch.from_numpy(vs)) ** 2).sum()
    total = pi_loss + vf_coeff * vf_loss + ent_coeff * ent_[:-1].sum()
    for g in optimizer.param_groups:
        g['lr'] = lr
    optimizer.zero_grad()
    total.backward()
    gn = torch.sqrt(sum((q.grad ** 2).sum() for q in model.parameters()))
    scale = 40.0 / max(float(gn), 40.0)                                               # ClipGradByGlobalNorm(40)
    for q in model.parameters():
        q.grad.mul_(scale)
    optimizer.step()
    return total.item(), kl.item()


def run_cpu_impala(seconds=15.0, n_actors=None, env_num=5, T=50, act_dim=18, train_batch_size=1000, seed=0):
    """Runs actors + learner for about `seconds`; returns a dict with env-steps/s measured exactly as
    the reference logs it (sample_total_steps / elapsed, examples/IMPALA/train.py:93,227,243)."""
    cores = os.cpu_count() or 1
    if n_actors is None:
        n_actors = max(1, cores - 2)
    learner_threads = max(1, cores - n_actors)
    torch.set_num_threads(learner_threads)
    ctx = mp.get_context('fork')
    model = CpuAtariModel(act_dim)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    conns, procs = [], []
    for i in range(n_actors):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_actor_proc, args=(b, act_dim, env_num, T, seed + 1 + i), daemon=True)
        p.start()
        conns.append(a)
        procs.append(p)

    def weights():
        return {k: v.detach().numpy() for k, v in model.state_dict().items()}
    import queue
    import threading
    state = dict(w=weights(), learn_steps=0, learn_time=0.0, stop=False)
    sample_q = queue.Queue(maxsize=8)                 # sample_queue_max_size (impala_config.py:33)

    def learner_loop():                               # Learner.run_learn thread (train.py:77-79,123-145)
        pending = []
        while not state['stop']:
            try:
                pending.append(sample_q.get(timeout=0.05))
            except queue.Empty:
                continue
            if sum(len(s['actions']) for s in pending) >= train_batch_size:
                batch = {k: np.concatenate([s[k] for s in pending]) for k in pending[0]}
                pending = []
                t1 = time.time()
                impala_learn(model, opt, batch, T)
                state['learn_time'] += time.time() - t1
                state['learn_steps'] += 1
                state['w'] = weights()

    th = threading.Thread(target=learner_loop, daemon=True)
    th.start()
    for c in conns:
        c.send(state['w'])
    t0 = time.time()
    steps = 0
    while time.time() - t0 < seconds:                 # run_remote_sample threads (train.py:165-194), one loop
        got = False
        for c in conns:
            if c.poll(0):
                sample = c.recv()
                steps += env_num * T                  # sample_total_steps += obs.shape[0]  (train.py:93)
                c.send(state['w'])
                try:
                    sample_q.put(sample, timeout=max(0.0, seconds - (time.time() - t0)))
                except queue.Full:
                    pass
                got = True
        if not got:
            time.sleep(0.002)
    state['stop'] = True
    th.join(timeout=30)
    learn_steps, learn_time = state['learn_steps'], state['learn_time']
    elapsed = time.time() - t0
    for c in conns:
        try:
            c.send(None)
        except Exception:
            pass
    for p in procs:
        p.join(timeout=2)
        if p.is_alive():
            p.terminate()
    return dict(env_steps_per_s=steps / elapsed, elapsed_s=elapsed, sample_steps=steps, actors=n_actors,
                env_num=env_num, cores=cores, learner_threads=learner_threads, learn_steps=learn_steps,
                learn_ms_per_batch=(1e3 * learn_time / learn_steps) if learn_steps else None,
                train_batch_size=train_batch_size)
