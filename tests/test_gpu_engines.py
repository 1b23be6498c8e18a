"""GPU tests of the fused MLP rollout kernel (rl_rollout_mlp) and the on-device A2C / PPO / DQN engines:
  * the fused rollout is BIT-IDENTICAL to stepping the stand-alone kernels (rl_mlp_fwd, rl_sample_*, rl_env_*_step),
    which are themselves pinned bit-exactly to the CPU env twin (tests/test_gpu_envs.py);
  * the engines' returns / learner updates equal the reference-form computation (oracle GAE scans, the
    torch-autograd `parl.algorithms.*.learn` path on the same batch) within fp32 tolerance;
  * the lane-interleaved replay ring equals one reference ring per lane (oracle AtariReplay)."""
import numpy as np
import pytest
import torch

from oracle import returns as oret
from oracle import replay as orp

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


@pytest.mark.parametrize('B,T', [(200, 37), (9600, 6)])       # 16-env latency tiles / 64-env throughput tiles
def test_fused_rollout_cartpole_equals_unfused_kernels(B, T):
    from parl_b200 import kernels as K
    from parl_b200.engine.nets import CartPoleActorCritic
    torch.manual_seed(3)
    seed, off = 11, 5
    model = CartPoleActorCritic(4, 2).to(DEV)
    layers, act = model.native_layers()
    plan = K.MlpPlan([[(w.detach(), b.detach()) for (w, b) in seg] for seg in layers], act)
    # --- fused
    st_f = K.EpisodeStats(B, DEV)
    obs_cur = torch.zeros((B, 4), device=DEV)
    K.env_cartpole_step(torch.zeros((B, 4), device=DEV), obs_cur, None, None, None, st_f, seed, 0, max_episode_steps=30,
                        env_offset=off, reset=True)
    f = dict(obs=torch.empty((T, B, 4), device=DEV), act=torch.empty((T, B), dtype=torch.int32, device=DEV),
             rew=torch.empty((T, B), device=DEV), done=torch.empty((T, B), dtype=torch.uint8, device=DEV),
             val=torch.empty((T + 1, B), device=DEV), logp=torch.empty((T, B), device=DEV),
             logits=torch.empty((T, B, 2), device=DEV))
    plan.rollout(K.ENV_CARTPOLE, K.POLICY_CATEGORICAL, T, obs_cur, st_f, seed, 0, f['obs'], f['act'], f['rew'],
                 f['done'], logp_out=f['logp'], val_out=f['val'], logits_out=f['logits'], env_offset=off,
                 max_episode_steps=30)
    # --- unfused: the stand-alone kernels, one step at a time
    st_u = K.EpisodeStats(B, DEV)
    state, obs = torch.zeros((B, 4), device=DEV), torch.zeros((B, 4), device=DEV)
    K.env_cartpole_step(state, obs, None, None, None, st_u, seed, 0, max_episode_steps=30, env_offset=off, reset=True)
    rew, done = torch.zeros(B, device=DEV), torch.zeros(B, dtype=torch.uint8, device=DEV)
    for t in range(T):
        assert torch.equal(f['obs'][t], obs), t
        logits, val = plan.forward(obs.clone(), split=2)
        a, lp = K.sample_categorical(logits, seed, t, env_offset=off, want_logp=True)
        assert torch.equal(f['logits'][t], logits) and torch.equal(f['val'][t], val.view(-1)), t
        assert torch.equal(f['act'][t], a), t
        assert torch.allclose(f['logp'][t], lp, rtol=1e-6, atol=1e-6), t
        K.env_cartpole_step(state, obs, rew, done, a, st_u, seed, t, max_episode_steps=30, env_offset=off)
        assert torch.equal(f['rew'][t], rew) and torch.equal(f['done'][t], done), t
    assert torch.equal(obs_cur, obs)
    _, val = plan.forward(obs.clone(), split=2)
    assert torch.equal(f['val'][T], val.view(-1))
    assert torch.equal(st_f.totals, st_u.totals) and torch.equal(st_f.ep_len, st_u.ep_len)
    assert T < 30 or f['done'].sum().item() > 0           # episodes did end (30-step limit / pole falls)


def test_fused_rollout_mujoco_gaussian_equals_unfused_kernels():
    from parl_b200 import kernels as K
    from parl_b200.engine.nets import MujocoModel
    torch.manual_seed(4)
    B, T, seed, off = 130, 25, 21, 64
    model = MujocoModel(17, 6).to(DEV)
    with torch.no_grad():
        model.fc_pi_std.copy_(torch.linspace(-0.5, 0.3, 6).view(1, 6))
    layers, act = model.native_layers()
    plan = K.MlpPlan([[(w.detach(), b.detach()) for (w, b) in seg] for seg in layers], act)
    logstd = model.fc_pi_std.detach().reshape(-1).contiguous()
    st_f, st_u = K.EpisodeStats(B, DEV), K.EpisodeStats(B, DEV)
    obs_cur = torch.zeros((B, 17), device=DEV)
    K.env_mujoco_synth_step(obs_cur, None, None, st_f, seed, 0, env_offset=off, reset=True)
    f = dict(obs=torch.empty((T, B, 17), device=DEV), act=torch.empty((T, B, 6), device=DEV),
             rew=torch.empty((T, B), device=DEV), done=torch.empty((T, B), dtype=torch.uint8, device=DEV),
             val=torch.empty((T + 1, B), device=DEV), logp=torch.empty((T, B), device=DEV))
    plan.rollout(K.ENV_MUJOCO_SYNTH, K.POLICY_GAUSSIAN, T, obs_cur, st_f, seed, 100, f['obs'], f['act'], f['rew'],
                 f['done'], logp_out=f['logp'], val_out=f['val'], logstd=logstd, env_offset=off, p_done=0.05,
                 max_episode_steps=12)
    obs = torch.zeros((B, 17), device=DEV)
    K.env_mujoco_synth_step(obs, None, None, st_u, seed, 0, env_offset=off, reset=True)
    # the fused rollout started at global step 100: observations after the reset come from counter step0+t+1
    rew, done = torch.zeros(B, device=DEV), torch.zeros(B, dtype=torch.uint8, device=DEV)
    for t in range(T):
        if t > 0:
            assert torch.equal(f['obs'][t], obs), t
        mean, val = plan.forward(f['obs'][t].clone(), split=6)
        a, lp = K.sample_gaussian(mean, logstd, seed, 100 + t, env_offset=off)
        assert torch.equal(f['act'][t], a) and torch.equal(f['val'][t], val.view(-1)), t
        assert torch.allclose(f['logp'][t], lp, rtol=1e-6, atol=1e-6), t
        K.env_mujoco_synth_step(obs, rew, done, st_u, seed, 100 + t, p_done=0.05, max_episode_steps=12, env_offset=off)
        assert torch.equal(f['rew'][t], rew) and torch.equal(f['done'][t], done), t
    assert torch.equal(obs_cur, obs)
    assert torch.equal(st_f.totals, st_u.totals)


def test_a2c_engine_returns_and_update_match_reference_form():
    from parl_b200 import kernels as K
    from parl_b200.algorithms import A2C
    from parl_b200.engine.a2c import A2CEngine
    from parl_b200.engine.nets import CartPoleActorCritic
    torch.manual_seed(0)
    eng = A2CEngine(num_envs=256, sample_batch_steps=20, seed=5, device=DEV, max_episode_steps=15)
    ref_model = CartPoleActorCritic(4, 2).to(DEV)
    ref_model.load_state_dict(eng.model.state_dict())
    ref = A2C(ref_model, dict(vf_loss_coeff=0.5, learning_rate=0.001))
    eng.rollout()
    T, B = eng.T, eng.B
    # returns: calc_gae per episode segment (fp64) as benchmark/torch/a2c/actor.py:82-102
    adv, tgt = K.gae_scan_segments(eng.rewards, eng.values[:T], eng.dones, eng.values[T], 0.99, 1.0)
    oadv, otgt = oret.a2c_segment_gae_time_major(eng.rewards.cpu().numpy(), eng.values[:T].cpu().numpy(),
                                                 eng.dones.cpu().numpy(), eng.values[T].cpu().numpy(), 0.99, 1.0)
    np.testing.assert_allclose(adv.cpu().numpy(), oadv, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(tgt.cpu().numpy(), otgt, rtol=1e-5, atol=1e-5)
    # one update: engine (fused MLP forward/backward) vs the torch-autograd algorithm on the same batch
    losses = eng.learn(0.001, -0.01).cpu().numpy()
    rl = ref.learn(eng.obs.view(T * B, 4), eng.actions.view(-1), adv.view(-1), tgt.view(-1), 0.001, -0.01)
    np.testing.assert_allclose(losses, np.array([float(x) for x in rl]), rtol=1e-4, atol=1e-3)
    for (n, p), (_, q) in zip(eng.model.named_parameters(), ref_model.named_parameters()):
        assert torch.allclose(p, q, rtol=1e-4, atol=1e-5), (n, (p - q).abs().max().item())
    for _ in range(3):
        out = eng.step()
    assert torch.isfinite(out).all() and eng.get_metrics()['episodes'] > 0


def test_ppo_engine_returns_and_minibatch_update_match_reference_form():
    from parl_b200.algorithms import PPO
    from parl_b200.engine.ppo import PPOEngine
    from parl_b200.engine.nets import MujocoModel
    torch.manual_seed(1)
    eng = PPOEngine(num_envs=64, step_nums=32, num_minibatches=4, update_epochs=2, seed=9, device=DEV, p_done=0.05,
                    max_episode_steps=20, num_updates=10)
    ref_model = MujocoModel(17, 6).to(DEV)
    ref_model.load_state_dict(eng.model.state_dict())
    ref = PPO(ref_model, clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.0, initial_lr=3e-4, continuous_action=True)
    for k in range(2):                                    # second rollout exercises the carried done flags
        eng.rollout()
        prev_last = eng.step_dones[eng.T - 1].clone()
    T, B, N = eng.T, eng.B, eng.N
    # storage.dones[t] is the done flag BEFORE observation t (train.py:97-98)
    assert torch.equal(eng.dones[1:], eng.step_dones[:T - 1].float())
    adv, ret = eng.compute_returns()
    oadv, oretn = oret.compute_returns(eng.rewards.cpu().numpy(), eng.values[:T].cpu().numpy(), eng.dones.cpu().numpy(),
                                       eng.values[T].cpu().numpy(), eng.last_done.cpu().numpy(), 0.99, 0.95)
    assert np.array_equal(adv.cpu().numpy(), oadv) and np.array_equal(ret.cpu().numpy(), oretn)     # bit-exact scan
    # sampled log-probs equal Normal(mean, std).log_prob(a).sum(1) of the reference model
    with torch.no_grad():
        mean, std = ref_model.policy(eng.obs.view(N, 17))
        lp = torch.distributions.Normal(mean, std).log_prob(eng.actions.view(N, 6)).sum(1)
        v = ref_model.value(eng.obs.view(N, 17)).view(-1)
    assert torch.allclose(eng.logprobs.view(-1), lp, rtol=1e-4, atol=1e-4)
    assert torch.allclose(eng.values[:T].reshape(-1), v, rtol=1e-5, atol=1e-5)
    # one minibatch update vs the torch-autograd algorithm
    idx = torch.randperm(N, device=DEV)[:eng.M].to(torch.int32)
    li = idx.long()
    losses = eng.learn_minibatch(idx, 3e-4).cpu().numpy()
    rl = ref.learn(eng.obs.view(N, 17)[li], eng.actions.view(N, 6)[li], eng.values[:T].reshape(-1)[li],
                   ret.view(-1)[li], eng.logprobs.view(-1)[li], adv.view(-1)[li], 3e-4)
    np.testing.assert_allclose(losses[:3], np.array(rl), rtol=1e-4, atol=1e-5)
    for (n, p), (_, q) in zip(eng.model.named_parameters(), ref_model.named_parameters()):
        assert torch.allclose(p, q, rtol=1e-4, atol=1e-5), (n, (p - q).abs().max().item())
    out, lr = eng.learn()
    assert torch.isfinite(out).all() and 0 < lr < 3e-4


def test_lane_replay_equals_one_reference_ring_per_lane():
    from parl_b200.engine.dqn import DeviceAtariReplay
    L, cap_q, ctx, hw = 3, 24, 4, (4, 4)
    rpm = DeviceAtariReplay(L * cap_q, hw, ctx, lanes=L, device=DEV)
    from collections import deque
    refs = [orp.AtariReplay(cap_q, hw, ctx) for _ in range(L)]
    ctxq = [deque(maxlen=ctx - 1) for _ in range(L)]
    rng = np.random.RandomState(0)
    frames = rng.randint(0, 255, (60, L) + hw).astype(np.uint8)
    cur = torch.as_tensor(frames[0]).to(DEV)
    rpm.current_plane().copy_(cur.view(L, -1))
    for t in range(55):                                   # wraps the 24-position ring twice
        act = rng.randint(0, 6, L).astype(np.int32)
        rew = rng.randn(L).astype(np.float32)
        over = rng.rand(L) < 0.15
        # recent_obs == rpm.recent_obs() + [obs] of every lane BEFORE the append (train.py:58-60,
        # replay_memory.py:36-56: a deque of the episode's last ctx-1 frames, cleared when isOver)
        got = rpm.recent_obs().cpu().numpy()
        for l in range(L):
            pad = [np.zeros(hw, np.uint8)] * (ctx - 1 - len(ctxq[l]))
            want = np.stack(pad + list(ctxq[l]) + [frames[t, l]])
            np.testing.assert_array_equal(got[l], want, err_msg='t=%d lane=%d' % (t, l))
            refs[l].append(frames[t, l], act[l], rew[l], over[l])
            if over[l]:
                ctxq[l].clear()
            else:
                ctxq[l].append(frames[t, l])
        pos = rpm.pos
        rpm.reward[pos].copy_(torch.as_tensor(rew).to(DEV))
        rpm.is_over[pos].copy_(torch.as_tensor(over.astype(np.uint8)).to(DEV))
        rpm.next_plane().copy_(torch.as_tensor(frames[t + 1]).to(DEV).view(L, -1))
        rpm.commit(torch.as_tensor(act).to(DEV))
    # every valid row: the 5-frame window, action, reward, terminal equal the reference ring's sample()
    rows = torch.arange(L * cap_q, dtype=torch.int32, device=DEV)
    ok = rpm.valid_rows(rows).cpu().numpy()
    assert ok.sum() >= L * (cap_q - 2 * ctx - 2)
    # the newest transition's next frame already sits in the device ring (written in place by the env step) while
    # the reference ring only receives it with the next append: compare from the second-newest position on
    age = (rpm.pos - 1 - rows.cpu().numpy() // L) % cap_q
    ok &= age >= 1
    sel = rows[torch.as_tensor(ok).to(DEV)]
    obs, act, rew, nobs, term = rpm.gather(sel)
    obs, nobs = obs.cpu().numpy(), nobs.cpu().numpy()
    for i, r in enumerate(sel.cpu().numpy()):
        q, l = divmod(int(r), L)
        o, rr, aa, tt = refs[l].sample((q - (ctx - 1)) % cap_q)
        np.testing.assert_array_equal(obs[i], o[:ctx])
        np.testing.assert_array_equal(nobs[i], o[1:])
        assert act[i].item() == aa and rew[i].item() == rr and bool(term[i].item()) == bool(tt)


@pytest.mark.parametrize('prioritized,double_q', [(True, False), (False, True)])
def test_dqn_engine_runs_and_priorities_follow_td(prioritized, double_q):
    from parl_b200.engine.dqn import DQNEngine
    torch.manual_seed(2)
    eng = DQNEngine(memory_size=64 * 40, num_envs=64, batch_size=32, act_dim=6, prioritized=prioritized,
                    double_q=double_q, seed=3, device=DEV, update_freq=2)
    eng.warmup(12)
    assert eng.rpm.size() == 12 * 64
    w0 = [p.detach().clone() for p in eng.model.parameters()]
    for _ in range(6):
        loss = eng.step()
    assert np.isfinite(float(loss))
    assert any(not torch.equal(a, b) for a, b in zip(w0, eng.model.parameters()))
    if prioritized:
        tree = eng.tree.tree.cpu().numpy()
        cap = eng.rpm.max_size
        leaves = tree[cap - 1:]
        assert abs(tree[0] - leaves.sum()) <= 1e-9 * max(tree[0], 1.0)          # sum-tree invariant after updates
        assert (leaves > 0).sum() == eng.rpm.size()
        assert 0.5 < eng.beta <= 1.0
    m = eng.get_metrics()
    assert m['learn_steps'] == 6 and m['sample_steps'] == (12 + 12) * 64


def test_device_vecnormalize_matches_reference_trace(golden):
    """rl_vecnormalize_step vs the trace recorded from parl.env.mujoco_wrappers.VecNormalizeEnv (float32 I/O)."""
    from parl_b200 import kernels as K
    g = golden('vecnormalize')
    T, B, D = g['term_seq'].shape
    vn = K.VecNormalize(B, D, DEV)
    f32 = lambda x: torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(DEV)
    ob0 = vn.reset(f32(g['obs_seq'][0]))
    np.testing.assert_allclose(ob0.cpu().numpy(), g['ob0'], rtol=2e-6, atol=2e-6)
    for t in range(T):
        rew = f32(g['rew_seq'][t])
        done = torch.as_tensor(g['done_seq'][t].astype(np.uint8)).to(DEV)
        ob = vn.step(f32(g['obs_seq'][t + 1]), rew, done, terminal_obs=f32(g['term_seq'][t]))
        np.testing.assert_allclose(ob.cpu().numpy(), g['ob_out'][t], rtol=1e-5, atol=1e-5, err_msg='t=%d' % t)
        np.testing.assert_allclose(rew.cpu().numpy(), g['rew_out'][t], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(vn.ob_count.cpu().numpy(), g['ob_count'], rtol=1e-12)
    np.testing.assert_allclose(vn.ob_var.cpu().numpy(), g['ob_var'], rtol=1e-5)        # float32 inputs
    np.testing.assert_allclose(vn.ret_var.cpu().numpy(), g['ret_var'], rtol=1e-5)


def test_fused_rollout_with_vecnormalize_equals_unfused_kernels():
    from parl_b200 import kernels as K
    from parl_b200.engine.nets import MujocoModel
    torch.manual_seed(6)
    B, T, seed = 70, 30, 4
    model = MujocoModel(17, 6).to(DEV)
    layers, act = model.native_layers()
    plan = K.MlpPlan([[(w.detach(), b.detach()) for (w, b) in seg] for seg in layers], act)
    logstd = model.fc_pi_std.detach().reshape(-1).contiguous()
    st_f, st_u = K.EpisodeStats(B, DEV), K.EpisodeStats(B, DEV)
    vn_f, vn_u = K.VecNormalize(B, 17, DEV), K.VecNormalize(B, 17, DEV)
    obs_cur = torch.zeros((B, 17), device=DEV)
    K.env_mujoco_synth_step(obs_cur, None, None, st_f, seed, 0, reset=True)
    vn_f.reset(obs_cur)
    f = dict(obs=torch.empty((T, B, 17), device=DEV), act=torch.empty((T, B, 6), device=DEV),
             rew=torch.empty((T, B), device=DEV), done=torch.empty((T, B), dtype=torch.uint8, device=DEV),
             val=torch.empty((T + 1, B), device=DEV))
    plan.rollout(K.ENV_MUJOCO_SYNTH, K.POLICY_GAUSSIAN, T, obs_cur, st_f, seed, 0, f['obs'], f['act'], f['rew'],
                 f['done'], val_out=f['val'], logstd=logstd, p_done=0.1, vecnorm=vn_f)
    obs = torch.zeros((B, 17), device=DEV)
    K.env_mujoco_synth_step(obs, None, None, st_u, seed, 0, reset=True)
    vn_u.reset(obs)
    rew, done = torch.zeros(B, device=DEV), torch.zeros(B, dtype=torch.uint8, device=DEV)
    for t in range(T):
        assert torch.equal(f['obs'][t], obs), t
        mean, _ = plan.forward(obs.clone(), split=6)
        a, _ = K.sample_gaussian(mean, logstd, seed, t)
        assert torch.equal(f['act'][t], a), t
        K.env_mujoco_synth_step(obs, rew, done, st_u, seed, t, p_done=0.1)
        vn_u.step(obs, rew, done)
        assert torch.equal(f['rew'][t], rew) and torch.equal(f['done'][t], done), t
    assert torch.equal(obs_cur, obs)
    for a, b in zip(vn_f._state, vn_u._state):
        assert torch.equal(a, b)
    assert torch.equal(st_f.totals, st_u.totals)           # raw rewards in the episode statistics


def test_rollout_storage_facade_matches_reference_fixture(golden):
    """parl_b200.utils.RolloutStorage (append / compute_returns / sample_batch) on the reference-recorded
    compute_returns fixture (tests/golden/ppo_returns.npz, from benchmark/torch/ppo/storage.py, gamma .99 / lambda .95)."""
    from parl_b200.utils import RolloutStorage
    g = golden('ppo_returns')
    for c in range(int(g['n_cases'])):
        pre = 'c%d_' % c
        rew, val, don = g[pre + 'rewards'], g[pre + 'values'], g[pre + 'dones']
        T, B = rew.shape
        st = RolloutStorage(T, B, 17, 6, device=DEV)
        rng = np.random.RandomState(c)
        obs, act = rng.randn(T, B, 17).astype(np.float32), rng.randn(T, B, 6).astype(np.float32)
        lp = rng.randn(T, B).astype(np.float32)
        for t in range(T):
            st.append(obs[t], act[t], lp[t], rew[t], don[t], val[t])
        adv, ret = st.compute_returns(g[pre + 'value'], g[pre + 'done'])
        assert np.array_equal(adv.cpu().numpy(), g[pre + 'adv']) and np.array_equal(ret.cpu().numpy(), g[pre + 'ret'])
        idx = rng.permutation(T * B)[:min(37, T * B)]
        o, a, l, ad, r, v = st.sample_batch(idx, as_numpy=True)
        assert np.array_equal(o, obs.reshape(T * B, 17)[idx]) and np.array_equal(a, act.reshape(T * B, 6)[idx])
        assert np.array_equal(l, lp.reshape(-1)[idx]) and np.array_equal(ad, g[pre + 'adv'].reshape(-1)[idx])
        assert np.array_equal(r, g[pre + 'ret'].reshape(-1)[idx]) and np.array_equal(v, val.reshape(-1)[idx])


def test_atari_replay_memory_facade_matches_reference_fixture(golden):
    """parl_b200.utils.atari_replay_memory.ReplayMemory vs the reference-recorded sample() stacking."""
    from parl_b200.utils.atari_replay_memory import ReplayMemory, Experience
    g = golden('atari_replay')
    size, ctx = int(g['size']), int(g['ctx'])
    shp = tuple(g['frames'].shape[1:])
    # frames of the fixture are 3x2 bytes: widen to 16 bytes per frame (the device gather's alignment contract)
    rpm = ReplayMemory(size, (4, 4), ctx, device=DEV)
    ref_ctx = []
    for f, a, r, o in zip(g['frames'], g['actions'], g['rewards'], g['overs']):
        wide = np.zeros((4, 4), np.uint8)
        wide.reshape(-1)[:f.size] = f.reshape(-1)
        rpm.append(Experience(wide, int(a), float(r), bool(o)))
    assert rpm.size() == min(len(g['frames']), size)
    ref = orp.AtariReplay(size, shp, ctx)
    for f, a, r, o in zip(g['frames'], g['actions'], g['rewards'], g['overs']):
        ref.append(f, a, r, o)
    idx = ref.batch_indices(g['raw'])
    obs, act, rew, over = rpm.sample_batch_by_index(idx)
    got = obs.reshape(len(idx), ctx + 1, 16)[:, :, :g['frames'][0].size].reshape((len(idx), ctx + 1) + shp)
    np.testing.assert_array_equal(got, g['obs'])
    np.testing.assert_array_equal(act.astype(np.int32), g['action'])
    assert len(rpm.recent_obs()) == ctx - 1


def test_policy_gradient_engine_matches_algorithm_and_learns_cartpole():
    """QuickStart on the device: one update equals parl.algorithms.PolicyGradient.learn on the same batch, and a
    short training run raises the mean CartPole episode length (benchmark/torch/QuickStart/train.py)."""
    from parl_b200 import kernels as K
    from parl_b200.algorithms import PolicyGradient
    from parl_b200.engine.pg import PolicyGradientEngine
    from parl_b200.engine.nets import CartPolePolicy
    torch.manual_seed(0)
    eng = PolicyGradientEngine(num_envs=128, rollout_steps=64, lr=1e-3, seed=3, device=DEV)
    ref_model = CartPolePolicy(4, 2).to(DEV)
    ref_model.load_state_dict(eng.model.state_dict())
    ref = PolicyGradient(ref_model, lr=1e-3)
    eng.rollout()
    T, B = eng.T, eng.B
    togo, _ = K.gae_scan_segments(eng.rewards, eng.zeros_tb, eng.dones, eng.zeros_b, 1.0, 1.0)
    # reward-to-go of CartPole (reward 1 per step) inside a segment = number of remaining steps of the segment
    d = eng.dones.cpu().numpy()
    tg = togo.cpu().numpy()
    for b in range(0, B, 17):
        cnt = 0.0
        for t in range(T - 1, -1, -1):
            cnt = 1.0 if d[t, b] else cnt + 1.0
            assert tg[t, b] == cnt
    loss = eng.learn().item()
    rloss = float(ref.learn(eng.obs.view(T * B, 4), eng.actions.view(-1), togo.view(-1)))
    assert abs(loss - rloss) <= 1e-4 * max(1.0, abs(rloss))
    for (n, p), (_, q) in zip(eng.model.named_parameters(), ref_model.named_parameters()):
        assert torch.allclose(p, q, rtol=1e-4, atol=1e-5), n
    eng2 = PolicyGradientEngine(num_envs=256, rollout_steps=200, lr=5e-3, seed=1, device=DEV)
    eng2.rollout()
    first = eng2.get_metrics()['mean_episode_steps']
    for _ in range(60):
        eng2.step()
    m = eng2.get_metrics()
    eng2.stats.totals.zero_()
    for _ in range(3):
        eng2.rollout()
    last = eng2.get_metrics()['mean_episode_steps']
    assert last > 1.5 * first and last > 40, (first, last)
