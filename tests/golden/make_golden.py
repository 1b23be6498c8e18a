"""Generate golden fixtures by RUNNING THE REFERENCE (build container only).

    python tests/golden/make_golden.py

Imports PaddlePaddle/PARL from /root/reference with the stub recipe of
SURVEY.md §8(c) (termcolor stub, pyarrow hidden, PARL_BACKEND=torch) and
records, on seeded inputs:
  * vtrace_kat.npz      — inputs + O(T^2) ground truth of the reference's own
                          known-answer test (vtrace_test_paddle.py:33-144)
  * a2c/ppo/dqn/ddqn/pg — losses returned by parl.algorithms.torch.*.learn and
                          the gradients autograd delivers to the network
                          outputs (captured with tensor hooks on "table" models
                          whose outputs ARE their parameters)
  * gae.npz             — parl.utils.calc_gae on random segments
  * ppo_returns.npz     — benchmark/torch/ppo/storage.py RolloutStorage.compute_returns
  * per.npz             — benchmark/fluid/Prioritized_DQN/proportional_per.py
                          SumTree/ProportionalPER store/sample/update trace
  * atari_replay.npz    — benchmark/torch/dqn/replay_memory.py sample() stacking
/root/reference does not exist on the GPU box; the .npz files are what travels.
"""
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def _setup_reference_import():
    stubs = tempfile.mkdtemp(prefix='parl_stubs_')
    with open(os.path.join(stubs, 'termcolor.py'), 'w') as f:
        f.write('def colored(s, *a, **k):\n    return s\n')
    with open(os.path.join(stubs, 'pyarrow.py'), 'w') as f:
        f.write('raise ImportError("hidden: forces cloudpickle path")\n')
    os.environ['PARL_BACKEND'] = 'torch'
    os.environ['PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION'] = 'python'
    os.environ['HOME'] = tempfile.mkdtemp(prefix='parl_home_')
    sys.path[:0] = [stubs, REF]


def main():
    _setup_reference_import()
    import warnings
    warnings.filterwarnings('ignore')
    import numpy as np
    import torch
    import parl
    from parl.algorithms import A2C, PPO, DQN, DDQN, PolicyGradient
    from parl.utils import calc_gae
    sys.path.insert(0, os.path.join(HERE, '..', '..'))
    from oracle import vtrace as ovt

    assert parl.__version__ == '2.2.1'
    rng = np.random.RandomState(1234)

    # ---- V-trace KAT (reference test's own ground truth) -------------------
    out = {}
    for B in (1, 4):
        k = ovt.kat_inputs(B)
        vs, pg = ovt.ground_truth_o_t2(**k)
        for name, val in k.items():
            out['B%d_%s' % (B, name)] = np.asarray(val, np.float32)
        out['B%d_vs' % B] = vs.astype(np.float32)
        out['B%d_pg_advantages' % B] = pg.astype(np.float32)
    np.savez(os.path.join(HERE, 'vtrace_kat.npz'), **out)

    grads = {}

    def hook(name):
        def _h(g):
            grads[name] = g.detach().clone().numpy()
        return _h

    # ---- A2C ----------------------------------------------------------------
    class A2CTable(parl.Model):
        def __init__(self, logits, values):
            super().__init__()
            self.lg = torch.nn.Parameter(torch.tensor(logits))
            self.v = torch.nn.Parameter(torch.tensor(values))
            self.lg.register_hook(hook('d_logits'))
            self.v.register_hook(hook('d_values'))

        def policy(self, obs):
            return self.lg

        def value(self, obs):
            return self.v

        def policy_and_value(self, obs):
            return self.lg, self.v

    out = {}
    for case, (N, A) in enumerate([(40, 2), (500, 6), (96, 18)]):
        logits = (rng.randn(N, A) * 2).astype(np.float32)
        values = rng.randn(N).astype(np.float32)
        actions = rng.randint(0, A, size=N).astype(np.int64)
        adv = rng.randn(N).astype(np.float32)
        tv = rng.randn(N).astype(np.float32)
        model = A2CTable(logits, values)
        alg = A2C(model, {'vf_loss_coeff': 0.5, 'learning_rate': 0.001})
        total, pi, vf, ent = alg.learn(torch.zeros(N, 1), torch.tensor(actions), torch.tensor(adv),
                                       torch.tensor(tv), 0.001, -0.01)
        p = 'c%d_' % case
        out.update({p + 'logits': logits, p + 'values': values, p + 'actions': actions, p + 'advantages': adv,
                    p + 'target_values': tv, p + 'total_loss': total.item(), p + 'pi_loss': pi.item(),
                    p + 'vf_loss': vf.item(), p + 'entropy': ent.item(),
                    p + 'd_logits': grads['d_logits'], p + 'd_values': grads['d_values']})
    np.savez(os.path.join(HERE, 'a2c.npz'), **out)

    # ---- PPO ----------------------------------------------------------------
    class PPOTable(parl.Model):
        def __init__(self, values, logits=None, mean=None, logstd=None):
            super().__init__()
            self.v = torch.nn.Parameter(torch.tensor(values))
            self.v.register_hook(hook('d_values'))
            if logits is not None:
                self.lg = torch.nn.Parameter(torch.tensor(logits))
                self.lg.register_hook(hook('d_logits'))
                self.cont = False
            else:
                self.mu = torch.nn.Parameter(torch.tensor(mean))
                self.ls = torch.nn.Parameter(torch.tensor(logstd))
                self.mu.register_hook(hook('d_mean'))
                self.ls.register_hook(hook('d_logstd'))
                self.cont = True

        def value(self, obs):
            return self.v

        def policy(self, obs):
            if self.cont:
                return self.mu, torch.exp(self.ls.expand_as(self.mu))   # mujoco_model.py:46-53
            return self.lg

    out = {}
    case = 0
    for cont in (False, True):
        for (M, D, clipv, norm) in [(64, 6, True, True), (257, 18 if not cont else 6, True, False), (128, 4, False, True)]:
            values = rng.randn(M).astype(np.float32)
            old_v = (values + 0.3 * rng.randn(M)).astype(np.float32)
            ret = rng.randn(M).astype(np.float32)
            adv = rng.randn(M).astype(np.float32)
            kw = dict(clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.01, initial_lr=3e-4, eps=1e-5,
                      max_grad_norm=0.5, use_clipped_value_loss=clipv, norm_adv=norm, continuous_action=cont)
            p = 'c%d_' % case
            if cont:
                mean = rng.randn(M, D).astype(np.float32)
                logstd = (0.3 * rng.randn(D)).astype(np.float32)
                act = (mean + np.exp(logstd) * rng.randn(M, D) * 1.3).astype(np.float32)
                model = PPOTable(values, mean=mean, logstd=logstd)
                with torch.no_grad():
                    old_lp = torch.distributions.Normal(torch.tensor(mean), torch.tensor(np.exp(logstd))).log_prob(
                        torch.tensor(act)).sum(1).numpy()
                old_lp = (old_lp + 0.3 * rng.randn(M)).astype(np.float32)
                out.update({p + 'mean': mean, p + 'logstd': logstd})
            else:
                logits = (rng.randn(M, D) * 2).astype(np.float32)
                act = rng.randint(0, D, size=M).astype(np.int64)
                model = PPOTable(values, logits=logits)
                with torch.no_grad():
                    old_lp = torch.distributions.Categorical(logits=torch.tensor(logits)).log_prob(torch.tensor(act)).numpy()
                old_lp = (old_lp + 0.3 * rng.randn(M)).astype(np.float32)
                out.update({p + 'logits': logits})
            alg = PPO(model, **kw)
            alg.model.cpu()
            vl, al, el = alg.learn(torch.zeros(M, 1), torch.tensor(act), torch.tensor(old_v), torch.tensor(ret),
                                   torch.tensor(old_lp), torch.tensor(adv))
            out.update({p + 'values': values, p + 'batch_value': old_v, p + 'batch_return': ret, p + 'batch_adv': adv,
                        p + 'batch_action': act, p + 'batch_logprob': old_lp, p + 'value_loss': vl, p + 'action_loss': al,
                        p + 'entropy_loss': el, p + 'd_values': grads['d_values'],
                        p + 'clipv': clipv, p + 'norm_adv': norm, p + 'continuous': cont})
            if cont:
                out.update({p + 'd_mean': grads['d_mean'], p + 'd_logstd': grads['d_logstd']})
            else:
                out.update({p + 'd_logits': grads['d_logits']})
            case += 1
    out['n_cases'] = case
    np.savez(os.path.join(HERE, 'ppo.npz'), **out)

    # ---- DQN / DDQN -----------------------------------------------------------
    class QTable(parl.Model):
        def __init__(self, q, q_next):
            super().__init__()
            self.q = torch.nn.Parameter(torch.tensor(q))
            self.qn = torch.nn.Parameter(torch.tensor(q_next))
            self.q.register_hook(hook('d_q'))

        def forward(self, obs):
            return self.q if float(obs.flatten()[0]) == 0.0 else self.qn

    out = {}
    for case, (M, A) in enumerate([(32, 2), (32, 6), (256, 18)]):
        q = rng.randn(M, A).astype(np.float32)
        q_next_online = rng.randn(M, A).astype(np.float32)
        q_next_target = rng.randn(M, A).astype(np.float32)
        action = rng.randint(0, A, size=(M, 1)).astype(np.int64)
        reward = np.clip(rng.randn(M, 1), -1, 1).astype(np.float32)
        terminal = (rng.rand(M, 1) < 0.2).astype(np.float32)
        p = 'c%d_' % case
        out.update({p + 'q': q, p + 'q_next_online': q_next_online, p + 'q_next_target': q_next_target,
                    p + 'action': action, p + 'reward': reward, p + 'terminal': terminal})
        for name, cls in (('dqn', DQN), ('ddqn', DDQN)):
            model = QTable(q, q_next_online)
            alg = cls(model, gamma=0.99, lr=1e-3)
            alg.model.cpu(), alg.target_model.cpu()
            with torch.no_grad():
                alg.target_model.qn.copy_(torch.tensor(q_next_target))
            loss = alg.learn(torch.zeros(M, 1), torch.tensor(action), torch.tensor(reward), torch.ones(M, 1),
                             torch.tensor(terminal))
            out.update({p + name + '_loss': loss, p + name + '_d_q': grads['d_q']})
    out['n_cases'] = 3
    np.savez(os.path.join(HERE, 'dqn.npz'), **out)

    # ---- PolicyGradient -------------------------------------------------------
    class ProbTable(parl.Model):
        def __init__(self, prob):
            super().__init__()
            self.p = torch.nn.Parameter(torch.tensor(prob))
            self.p.register_hook(hook('d_prob'))

        def forward(self, obs):
            return self.p

    N, A = 77, 2
    lg = rng.randn(N, A).astype(np.float32)
    prob = (np.exp(lg) / np.exp(lg).sum(-1, keepdims=True)).astype(np.float32)
    action = rng.randint(0, A, size=N).astype(np.int64)
    reward = rng.rand(N).astype(np.float32) * 20
    alg = PolicyGradient(ProbTable(prob), lr=1e-3)
    loss = alg.learn(torch.zeros(N, 1), torch.tensor(action), torch.tensor(reward))
    np.savez(os.path.join(HERE, 'pg.npz'), prob=prob, action=action, reward=reward, loss=loss.item(),
             d_prob=grads['d_prob'])

    # ---- calc_gae -------------------------------------------------------------
    out = {}
    for case, (L, gamma, lam) in enumerate([(1, 0.99, 1.0), (5, 0.99, 1.0), (20, 0.99, 0.95), (13, 0.9, 0.5)]):
        r = rng.rand(L)
        v = rng.randn(L)
        nv = float(rng.randn())
        adv = calc_gae(r, v, nv, gamma, lam)
        p = 'c%d_' % case
        out.update({p + 'rewards': r, p + 'values': v, p + 'next_value': nv, p + 'gamma': gamma, p + 'lam': lam,
                    p + 'adv': np.ascontiguousarray(adv)})
    out['n_cases'] = 4
    np.savez(os.path.join(HERE, 'gae.npz'), **out)

    # ---- PPO RolloutStorage.compute_returns ------------------------------------
    sys.path.insert(0, os.path.join(REF, 'benchmark/torch/ppo'))
    from storage import RolloutStorage
    out = {}
    for case, (T, B, pd) in enumerate([(8, 3, 0.3), (128, 8, 0.05), (33, 17, 0.1)]):
        obs_space = types.SimpleNamespace(shape=(2, ))
        act_space = types.SimpleNamespace(shape=())
        st = RolloutStorage(T, B, obs_space, act_space)
        st.rewards[:] = rng.rand(T, B)
        st.values[:] = rng.randn(T, B)
        st.dones[:] = (rng.rand(T, B) < pd)
        value = rng.randn(B).astype(np.float32)
        done = (rng.rand(B) < pd).astype(np.float32)
        adv, ret = st.compute_returns(value, done, 0.99, 0.95)
        p = 'c%d_' % case
        out.update({p + 'rewards': st.rewards.copy(), p + 'values': st.values.copy(), p + 'dones': st.dones.copy(),
                    p + 'value': value, p + 'done': done, p + 'adv': adv, p + 'ret': ret})
    out['n_cases'] = 3
    np.savez(os.path.join(HERE, 'ppo_returns.npz'), **out)

    # ---- SumTree / ProportionalPER ----------------------------------------------
    sys.path.insert(0, os.path.join(REF, 'benchmark/fluid/Prioritized_DQN'))
    import proportional_per as ref_per
    cap, seg = 64, 8
    per = ref_per.ProportionalPER(alpha=0.6, seg_num=seg, size=cap, framestack=1)
    item = (np.zeros(1), 0, 0.0, np.zeros(1), False)
    store_delta = []
    for i in range(cap):
        d = None if i < 40 else float(rng.rand() * 3)
        store_delta.append(-1.0 if d is None else d)
        per.store(item, d)
    tree_after_store = np.array(per.elements.tree, np.float64)
    us, idxs, ws, upd = [], [], [], []
    draws = iter(())

    def fake_uniform(low, high):
        u = next(draws)
        return low + u * (high - low)

    real_uniform = np.random.uniform
    # numpy>=1.24 refuses the reference's ragged ``np.array(items)`` (proportional_per.py:157);
    # give the reference module a numpy proxy whose array() falls back to dtype=object.
    class _NP(object):
        def __getattr__(self, k):
            return getattr(np, k)

        @staticmethod
        def array(x, *a, **k):
            try:
                return np.array(x, *a, **k)
            except ValueError:
                o = np.empty(len(x), dtype=object)
                o[:] = [tuple(i) for i in x]
                return o
    ref_per.np = _NP()
    for rnd in range(6):
        u = rng.rand(seg)
        draws = iter(u)
        ref_per.np.random.uniform = fake_uniform
        try:
            _, indices, w = per.sample(beta=0.5 + 0.1 * rnd)
        finally:
            ref_per.np.random.uniform = real_uniform
        newp = rng.rand(seg) * 2
        per.update(indices, newp)
        us.append(u), idxs.append(indices), ws.append(w), upd.append(newp)
    np.savez(os.path.join(HERE, 'per.npz'), capacity=cap, seg_num=seg, alpha=0.6, eps=0.01,
             store_delta=np.array(store_delta), tree_after_store=tree_after_store,
             u=np.array(us), indices=np.array(idxs), weights=np.array(ws), new_priorities=np.array(upd),
             tree_final=np.array(per.elements.tree, np.float64), min_final=per.elements._min,
             max_priority_final=per._max_priority)

    # ---- Atari replay stacking ---------------------------------------------------
    sys.path.insert(0, os.path.join(REF, 'benchmark/torch/dqn'))
    import importlib
    ref_rpm = importlib.import_module('replay_memory')
    Experience = ref_rpm.Experience
    size, ctx, shape = 50, 4, (3, 2)
    rpm = ref_rpm.ReplayMemory(size, shape, ctx)
    frames, acts, rews, overs = [], [], [], []
    for i in range(73):                      # wraps around
        f = rng.randint(0, 255, size=shape).astype(np.uint8)
        a, r, o = int(rng.randint(0, 6)), float(rng.rand()), bool(rng.rand() < 0.15)
        rpm.append(Experience(f, a, r, o))
        frames.append(f), acts.append(a), rews.append(r), overs.append(o)
    raw = rng.randint(rpm.size() - ctx - 1, size=24)
    batch_idx = (rpm._curr_pos + raw) % rpm._curr_size
    exps = [rpm.sample(i) for i in batch_idx]
    np.savez(os.path.join(HERE, 'atari_replay.npz'), size=size, ctx=ctx, frames=np.array(frames), actions=np.array(acts),
             rewards=np.array(rews, np.float32), overs=np.array(overs), raw=raw,
             obs=np.array([e[0] for e in exps]), reward=np.array([e[1] for e in exps], np.float32),
             action=np.array([e[2] for e in exps]), isOver=np.array([e[3] for e in exps]))
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()
