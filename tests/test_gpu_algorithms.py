"""GPU parity through the reference-facing API: parl_b200.algorithms.*.learn (fused loss kernels + FlatAdam)
against a plain-torch restatement of the reference ``learn`` bodies (same initial weights, one update),
comparing returned losses AND the updated parameters."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions import Categorical, Normal

import parl_b200 as parl
from oracle import vtrace as ovt

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


class ACModel(parl.Model):
    def __init__(self, obs_dim=8, act_dim=5):
        super().__init__()
        self.fc = nn.Linear(obs_dim, 32)
        self.pi = nn.Linear(32, act_dim)
        self.v = nn.Linear(32, 1)

    def policy(self, x):
        return self.pi(torch.tanh(self.fc(x)))

    def value(self, x):
        return self.v(torch.tanh(self.fc(x))).squeeze(1)

    def policy_and_value(self, x):
        h = torch.tanh(self.fc(x))
        return self.pi(h), self.v(h).squeeze(1)


def _params(m):
    return torch.cat([p.detach().reshape(-1).cpu() for p in m.parameters()])


def _clip_torch(params, max_norm):
    torch.nn.utils.clip_grad_norm_(params, max_norm)


def test_impala_learn_matches_reference_form():
    torch.manual_seed(0)
    T, B, A, D = 10, 12, 5, 8
    model = ACModel(D, A).to(DEV)
    ref = copy.deepcopy(model)
    alg = parl.algorithms.IMPALA(model, sample_batch_steps=T, gamma=0.99, vf_loss_coeff=0.5, clip_rho_threshold=1.0,
                                 clip_pg_rho_threshold=1.0)
    rng = np.random.RandomState(0)
    obs = rng.randn(B * T, D).astype(np.float32)                 # env-major flat, like Actor.sample()
    actions = rng.randint(0, A, B * T).astype(np.int64)
    bl = rng.randn(B * T, A).astype(np.float32)
    rewards = (rng.rand(B * T) < 0.5).astype(np.float32)
    dones = rng.rand(B * T) < 0.1
    loss, kl = alg.learn(obs, actions, bl, rewards, dones, 0.001, -0.01)
    # reference form (impala.py:134-215) in torch
    o = torch.tensor(obs, device=DEV)
    values, tl = ref.value(o), ref.policy(o)
    t_lsm, b_lsm = F.log_softmax(tl, -1), F.log_softmax(torch.tensor(bl, device=DEV), -1)
    oh = F.one_hot(torch.tensor(actions, device=DEV), A).float()
    tlp, blp = (t_lsm * oh).sum(-1), (b_lsm * oh).sum(-1)
    ent = -(t_lsm.exp() * t_lsm).sum(-1)
    ref_kl = (t_lsm.exp() * (t_lsm - b_lsm)).sum(-1).mean()
    tm = lambda x: x.reshape(B, T).transpose(0, 1)
    tlp_, blp_, ent_, v_ = tm(tlp), tm(blp), tm(ent), tm(values)
    rew, dn = tm(torch.tensor(rewards, device=DEV)), tm(torch.tensor(dones, device=DEV))
    disc = (~dn[:-1]).float() * 0.99
    vs, pg = ovt.from_importance_weights(blp_[:-1].detach().cpu().numpy(), tlp_[:-1].detach().cpu().numpy(),
                                         disc.cpu().numpy(), rew[:-1].cpu().numpy(), v_[:-1].detach().cpu().numpy(),
                                         v_[-1].detach().cpu().numpy(), 1.0, 1.0)
    pi_loss = -(tlp_[:-1] * torch.tensor(pg, device=DEV)).sum()
    vf_loss = 0.5 * ((v_[:-1] - torch.tensor(vs, device=DEV)) ** 2).sum()
    total = pi_loss + 0.5 * vf_loss + (-0.01) * ent_[:-1].sum()
    opt = torch.optim.Adam(ref.parameters(), lr=0.001)
    total.backward()
    gn = torch.sqrt(sum((p.grad ** 2).sum() for p in ref.parameters()))
    for p in ref.parameters():
        p.grad.mul_(40.0 / max(gn.item(), 40.0))                 # paddle ClipGradByGlobalNorm(40)
    opt.step()
    np.testing.assert_allclose(loss.total_loss.item(), total.item(), rtol=1e-4)
    np.testing.assert_allclose(loss.pi_loss.item(), pi_loss.item(), rtol=1e-4)
    np.testing.assert_allclose(loss.vf_loss.item(), vf_loss.item(), rtol=1e-4)
    np.testing.assert_allclose(loss.entropy.item(), ent_[:-1].sum().item(), rtol=1e-4)
    np.testing.assert_allclose(kl.item(), ref_kl.item(), rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(_params(model).numpy(), _params(ref).numpy(), rtol=1e-3, atol=2e-5)
    # the numpy-dict weight contract still works on the flat-buffer parameters
    w = alg.get_weights()
    alg.set_weights(w)
    assert all(isinstance(v, np.ndarray) for v in w.values())


def test_a2c_learn_matches_reference_form():
    torch.manual_seed(1)
    N, A, D = 500, 5, 8
    model = ACModel(D, A).to(DEV)
    ref = copy.deepcopy(model)
    alg = parl.algorithms.A2C(model, {'vf_loss_coeff': 0.5, 'learning_rate': 0.001})
    rng = np.random.RandomState(1)
    obs = rng.randn(N, D).astype(np.float32)
    actions = rng.randint(0, A, N).astype(np.int64)
    adv, tv = rng.randn(N).astype(np.float32), rng.randn(N).astype(np.float32)
    # large gradients so that clip_grad_norm_(40) is active
    total, pi, vf, ent = alg.learn(obs, actions, adv * 30, tv * 30, 0.002, -0.01)
    o = torch.tensor(obs, device=DEV)
    logits, values = ref.policy(o), ref.value(o)
    logp = (F.log_softmax(logits, 1) * F.one_hot(torch.tensor(actions, device=DEV), A)).sum(-1)
    r_pi = -(logp * torch.tensor(adv * 30, device=DEV)).sum()
    r_vf = 0.5 * ((values - torch.tensor(tv * 30, device=DEV)) ** 2).sum()
    r_ent = Categorical(logits=logits).entropy().sum()
    r_total = r_pi + 0.5 * r_vf - 0.01 * r_ent
    opt = torch.optim.Adam(ref.parameters(), lr=0.002)
    r_total.backward()
    _clip_torch(ref.parameters(), 40.0)
    opt.step()
    for got, want in ((total, r_total), (pi, r_pi), (vf, r_vf), (ent, r_ent)):
        np.testing.assert_allclose(got.item(), want.item(), rtol=1e-4)
    np.testing.assert_allclose(_params(model).numpy(), _params(ref).numpy(), rtol=1e-3, atol=2e-5)
    acts, vals = alg.sample(obs)
    assert acts.dtype == torch.int64 and acts.shape == (N, ) and vals.shape == (N, )
    assert alg.predict(obs).shape == (N, )


class MujocoLike(parl.Model):
    def __init__(self, obs_dim=17, act_dim=6):
        super().__init__()
        self.v1, self.v2 = nn.Linear(obs_dim, 32), nn.Linear(32, 1)
        self.p1, self.p2 = nn.Linear(obs_dim, 32), nn.Linear(32, act_dim)
        self.fc_pi_std = nn.Parameter(0.1 * torch.randn(1, act_dim))

    def value(self, x):
        return self.v2(torch.tanh(self.v1(x)))

    def policy(self, x):
        mean = self.p2(torch.tanh(self.p1(x)))
        return mean, torch.exp(self.fc_pi_std.expand_as(mean))


@pytest.mark.parametrize('continuous', [True, False])
def test_ppo_learn_matches_reference_form(continuous):
    torch.manual_seed(2)
    M, D, A = 512, 17, 6
    model = (MujocoLike(D, A) if continuous else ACModel(D, A)).to(DEV)
    ref = copy.deepcopy(model)
    kw = dict(clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.01, initial_lr=3e-4, eps=1e-5, max_grad_norm=0.5,
              use_clipped_value_loss=True, norm_adv=True, continuous_action=continuous)
    alg = parl.algorithms.PPO(model, **kw)
    rng = np.random.RandomState(2)
    obs = rng.randn(M, D).astype(np.float32)
    o = torch.tensor(obs, device=DEV)
    with torch.no_grad():
        if continuous:
            mean, std = ref.policy(o)
            act = (mean + std * torch.randn_like(mean)).cpu().numpy()
            oldlp = Normal(mean, std).log_prob(torch.tensor(act, device=DEV)).sum(1).cpu().numpy()
        else:
            act = rng.randint(0, A, M).astype(np.int64)
            oldlp = Categorical(logits=ref.policy(o)).log_prob(torch.tensor(act, device=DEV)).cpu().numpy()
        oldv = ref.value(o).reshape(-1).cpu().numpy()
    oldlp = (oldlp + 0.2 * rng.randn(M)).astype(np.float32)
    oldv = (oldv + 0.2 * rng.randn(M)).astype(np.float32)
    ret, adv = rng.randn(M).astype(np.float32), rng.randn(M).astype(np.float32)
    vl, al, el = alg.learn(obs, act, oldv, ret, oldlp, adv, lr=3e-4)
    # reference form (ppo.py:102-147)
    t = lambda x: torch.tensor(x, device=DEV)
    values = ref.value(o)
    if continuous:
        mean, std = ref.policy(o)
        dist = Normal(mean, std)
        lp, ent = dist.log_prob(t(act)).sum(1), dist.entropy().sum(1)
    else:
        dist = Categorical(logits=ref.policy(o))
        lp, ent = dist.log_prob(t(act)), dist.entropy()
    a = t(adv)
    a = (a - a.mean()) / (a.std() + 1e-8)
    ratio = torch.exp(lp - t(oldlp))
    r_al = -torch.min(ratio * a, torch.clamp(ratio, 0.8, 1.2) * a).mean()
    v = values.view(-1)
    vclip = t(oldv) + torch.clamp(v - t(oldv), -0.2, 0.2)
    r_vl = 0.5 * torch.max((v - t(ret)).pow(2), (vclip - t(ret)).pow(2)).mean()
    loss = r_vl * 0.5 + r_al - ent.mean() * 0.01
    opt = torch.optim.Adam(ref.parameters(), lr=3e-4, eps=1e-5)
    loss.backward()
    _clip_torch(ref.parameters(), 0.5)
    opt.step()
    np.testing.assert_allclose(vl, r_vl.item(), rtol=1e-4)
    np.testing.assert_allclose(al, r_al.item(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(el, ent.mean().item(), rtol=1e-4)
    np.testing.assert_allclose(_params(model).numpy(), _params(ref).numpy(), rtol=1e-3, atol=2e-5)
    value, action, logp, entropy = alg.sample(obs)
    assert action.shape[0] == M and logp.shape == (M, )


class QNet(parl.Model):
    def __init__(self, obs_dim=4, act_dim=3):
        super().__init__()
        self.f1, self.f2 = nn.Linear(obs_dim, 32), nn.Linear(32, act_dim)

    def forward(self, x):
        return self.f2(torch.relu(self.f1(x)))


@pytest.mark.parametrize('double_q', [False, True])
def test_dqn_ddqn_learn_matches_reference_form(double_q):
    torch.manual_seed(3)
    M, D, A = 32, 4, 3
    model = QNet(D, A).to(DEV)
    ref = copy.deepcopy(model)
    cls = parl.algorithms.DDQN if double_q else parl.algorithms.DQN
    alg = cls(model, gamma=0.99, lr=1e-3)
    with torch.no_grad():                       # make the target network differ from the online one
        for p in alg.target_model.parameters():
            p.add_(0.05 * torch.randn_like(p))
    ref_tgt = copy.deepcopy(alg.target_model)
    rng = np.random.RandomState(3)
    obs, nobs = rng.randn(M, D).astype(np.float32), rng.randn(M, D).astype(np.float32)
    act = rng.randint(0, A, (M, 1)).astype(np.int64)
    rew = np.clip(rng.randn(M, 1), -1, 1).astype(np.float32)
    term = (rng.rand(M, 1) < 0.2).astype(np.float32)
    loss = alg.learn(obs, act, rew, nobs, term)
    t = lambda x: torch.tensor(x, device=DEV)
    pred = ref(t(obs)).gather(1, t(act))
    with torch.no_grad():
        if double_q:
            greedy = ref(t(nobs)).max(1, keepdim=True)[1]
            max_v = ref_tgt(t(nobs)).gather(1, greedy)
        else:
            max_v = ref_tgt(t(nobs)).max(1, keepdim=True)[0]
        target = t(rew) + (1 - t(term)) * 0.99 * max_v
    r_loss = F.mse_loss(pred, target)
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    r_loss.backward()
    opt.step()
    np.testing.assert_allclose(loss, r_loss.item(), rtol=1e-5)
    np.testing.assert_allclose(_params(model).numpy(), _params(ref).numpy(), rtol=1e-4, atol=1e-6)
    alg.sync_target()
    np.testing.assert_array_equal(_params(alg.target_model).numpy(), _params(model).numpy())


def test_policy_gradient_and_replay_memory():
    torch.manual_seed(4)

    class P(parl.Model):
        def __init__(self):
            super().__init__()
            self.f = nn.Linear(4, 2)

        def forward(self, x):
            return F.softmax(self.f(x), -1)
    model = P().to(DEV)
    ref = copy.deepcopy(model)
    alg = parl.algorithms.PolicyGradient(model, lr=1e-3)
    rng = np.random.RandomState(4)
    obs = rng.randn(50, 4).astype(np.float32)
    act = rng.randint(0, 2, 50).astype(np.int64)
    rew = (rng.rand(50) * 10).astype(np.float32)
    loss = alg.learn(obs, act, rew)
    prob = ref(torch.tensor(obs, device=DEV))
    r_loss = torch.mean(-Categorical(prob).log_prob(torch.tensor(act, device=DEV)) * torch.tensor(rew, device=DEV))
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    r_loss.backward()
    opt.step()
    np.testing.assert_allclose(loss.item(), r_loss.item(), rtol=1e-5)
    np.testing.assert_allclose(_params(model).numpy(), _params(ref).numpy(), rtol=1e-4, atol=1e-6)
    # HBM replay memory with the reference interface
    rpm = parl.utils.ReplayMemory(100, 4, 0)
    for i in range(130):
        rpm.append(np.full(4, i, np.float32), i % 2, float(i), np.full(4, i + 1, np.float32), i % 7 == 0)
    assert rpm.size() == 100
    o, a, r, no, t = rpm.sample_batch(32)
    assert o.shape == (32, 4) and a.shape == (32, ) and t.dtype == bool
    assert np.all(no[:, 0] == o[:, 0] + 1) and np.all(r == o[:, 0]) and np.all(a == (o[:, 0].astype(int) % 2))
    import tempfile, os
    path = os.path.join(tempfile.mkdtemp(), 'rpm.npz')
    rpm.save(path)
    rpm2 = parl.utils.ReplayMemory(100, 4, 0)
    rpm2.load(path)
    assert rpm2.size() == 100 and torch.equal(rpm2.obs, rpm.obs)


def test_device_vector_envs_and_remote_actor_contract():
    """@parl.remote_class Actor hosting a device env pool: set_weights / sample() -> numpy dict / get_metrics."""
    from parl_b200.env import AtariSynthVectorEnv, CartPoleVectorEnv
    parl.connect('localhost:8010')
    env = AtariSynthVectorEnv(16, seed=3, device=DEV)
    o = env.reset()
    assert o.shape == (16, 4, 84, 84) and o.dtype == torch.uint8
    for _ in range(30):
        o, r, d, _ = env.step()
    assert len(env.next_episode_results()) > 0

    @parl.remote_class(wait=False)
    class Actor(object):
        def __init__(self, n):
            self.env = CartPoleVectorEnv(n, seed=1, device=DEV)
            self.obs = self.env.reset()
            self.model = ACModel(4, 2).to(DEV)

        def set_weights(self, w):
            self.model.set_weights(w)

        def sample(self):
            data = dict(obs=[], actions=[], rewards=[], dones=[])
            for _ in range(5):
                logits = self.model.policy(self.obs)
                a = logits.argmax(-1).int()
                nobs, r, d, _ = self.env.step(a)
                data['obs'].append(self.obs.cpu().numpy()), data['actions'].append(a.cpu().numpy())
                data['rewards'].append(r.cpu().numpy()), data['dones'].append(d.cpu().numpy())
                self.obs = nobs
            return {k: np.stack(v) for k, v in data.items()}
    actors = [Actor(8) for _ in range(2)]
    w = ACModel(4, 2).get_weights()
    for a in actors:
        a.set_weights(w)
    outs = [f.get() for f in [a.sample() for a in actors]]
    assert outs[0]['obs'].shape == (5, 8, 4) and outs[0]['dones'].dtype == bool
    np.testing.assert_array_equal(outs[0]['obs'], outs[1]['obs'])      # same seed, same weights -> same rollout
