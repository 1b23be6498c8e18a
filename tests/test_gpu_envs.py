"""GPU parity: device env steppers / samplers vs the CPU twins in oracle/envs.py.
uint8 frames, done masks, rewards and action indices must be BIT-EXACT."""
import numpy as np
import pytest
import torch

from oracle import envs as oenv
from oracle import philox as oph

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_atari_synth_bit_exact_and_stack():
    from parl_b200 import kernels as K
    B, HW, T, seed = 67, 84 * 84, 24, 12345
    ref = oenv.AtariSynthVec(B, seed, hw=HW, p_done=0.2)
    obs0 = ref.reset()
    planes = torch.zeros(T + 4, B, HW, dtype=torch.uint8, device=DEV)
    ages = torch.zeros(T + 1, B, dtype=torch.uint8, device=DEV)
    rew = torch.zeros(T, B, device=DEV)
    done = torch.zeros(T, B, dtype=torch.uint8, device=DEV)
    st = K.EpisodeStats(B, DEV, ring_cap=64)
    K.env_atari_synth_step(planes[3], None, None, None, ages[0], st, seed, 0, reset=True)
    out = torch.empty(B, 4, HW, dtype=torch.uint8, device=DEV)
    K.obs_stack_gather(planes, ages, 0, 1, out)
    assert np.array_equal(out.cpu().numpy(), obs0)
    ref_obs = []
    for t in range(T):
        K.env_atari_synth_step(planes[t + 4], rew[t], done[t], ages[t], ages[t + 1], st, seed, t, p_done=0.2)
        o, r, d = ref.step()
        ref_obs.append(o)
        assert np.array_equal(rew[t].cpu().numpy(), r)
        assert np.array_equal(done[t].cpu().numpy().astype(bool), d)
        assert np.array_equal(ages[t + 1].cpu().numpy(), ref.age)
    allobs = torch.empty(T * B, 4, HW, dtype=torch.uint8, device=DEV)
    K.obs_stack_gather(planes, ages, 1, T, allobs)
    assert np.array_equal(allobs.cpu().numpy().reshape(T, B, 4, HW), np.stack(ref_obs))
    # env-major output order + float32 scaling
    em = torch.empty(T * B, 4, HW, dtype=torch.float32, device=DEV)
    K.obs_stack_gather(planes, ages, 1, T, em, layout=K.ENV_MAJOR, scale=1.0 / 255.0)
    want = np.swapaxes(np.stack(ref_obs), 0, 1).reshape(T * B, 4, HW).astype(np.float32) * np.float32(1 / 255.0)
    np.testing.assert_allclose(em.cpu().numpy(), want, rtol=1e-6)
    # episode statistics gathered with warp ballots
    tot = st.totals.cpu().numpy()
    assert int(tot[0]) == len(ref.completed)
    np.testing.assert_allclose(tot[1], sum(r for r, _ in ref.completed), rtol=1e-6)
    np.testing.assert_allclose(tot[2], sum(l for _, l in ref.completed), rtol=1e-6)
    assert np.array_equal(st.ep_len.cpu().numpy(), ref.ep_len)
    n = min(len(ref.completed), 64)
    ring = sorted(zip(st.ring_ret.cpu().numpy()[:n].tolist(), st.ring_len.cpu().numpy()[:n].tolist()))
    if len(ref.completed) <= 64:
        assert ring == sorted((float(a), int(b)) for a, b in ref.completed)
    # pixel distribution: U{0..254}
    px = planes[4:].cpu().numpy()
    assert px.max() == 254 and px.min() == 0


def test_atari_synth_env_offset_shards_are_consistent():
    from parl_b200 import kernels as K
    B, HW, seed = 64, 256, 9
    full = torch.zeros(B, HW, dtype=torch.uint8, device=DEV)
    half = torch.zeros(B // 2, HW, dtype=torch.uint8, device=DEV)
    st = K.EpisodeStats(B, DEV)
    age = torch.zeros(B, dtype=torch.uint8, device=DEV)
    K.env_atari_synth_step(full, None, None, None, age, st, seed, 5, reset=True)
    K.env_atari_synth_step(half, None, None, None, age[:B // 2], K.EpisodeStats(B // 2, DEV), seed, 5, env_offset=32,
                           reset=True)
    assert torch.equal(full[32:], half)


@pytest.mark.parametrize('A', [2, 6, 18])
def test_fused_action_sampling_bit_exact(A):
    from parl_b200 import kernels as K
    B, seed = 1000, 4242
    rng = np.random.RandomState(A)
    logits = (3 * rng.randn(B, A)).astype(np.float32)
    lg = torch.as_tensor(logits).to(DEV)
    for step in (0, 1, 77):
        a = K.sample_categorical(lg, seed, step, env_offset=11)
        u = oph.action_uniforms(seed, np.arange(B, dtype=np.uint32) + 11, step)
        want = oph.sample_categorical_exact(logits, u)
        assert np.array_equal(a.cpu().numpy(), want)
    # fused into the env step
    HW = 64
    st = K.EpisodeStats(B, DEV)
    plane = torch.zeros(B, HW, dtype=torch.uint8, device=DEV)
    rew = torch.zeros(B, device=DEV)
    done = torch.zeros(B, dtype=torch.uint8, device=DEV)
    age = torch.zeros(B, dtype=torch.uint8, device=DEV)
    acts = torch.zeros(B, dtype=torch.int32, device=DEV)
    K.env_atari_synth_step(plane, rew, done, age, age, st, seed, 3, logits=lg, actions_out=acts)
    u = oph.action_uniforms(seed, np.arange(B, dtype=np.uint32), 3)
    assert np.array_equal(acts.cpu().numpy(), oph.sample_categorical_exact(logits, u))
    # log-prob side output
    a, lp = K.sample_categorical(lg, seed, 5, want_logp=True)
    ref_lp = torch.log_softmax(torch.as_tensor(logits), -1).numpy()[np.arange(B), a.cpu().numpy()]
    np.testing.assert_allclose(lp.cpu().numpy(), ref_lp, rtol=1e-5, atol=1e-5)


def test_sampling_distribution_chi2():
    from parl_b200 import kernels as K
    N = 400000
    logits = torch.tensor([[0.0, 1.0, 2.0, -1.0, 0.5, 0.5]], device=DEV).repeat(N, 1).contiguous()
    a = K.sample_categorical(logits, 1, 0).cpu().numpy()
    p = torch.softmax(logits[0].cpu(), -1).numpy()
    freq = np.bincount(a, minlength=6) / N
    assert np.abs(freq - p).max() < 4e-3


def test_mujoco_synth_matches_twin():
    from parl_b200 import kernels as K
    B, D, T, seed = 300, 17, 40, 5
    ref = oenv.MujocoSynthVec(B, seed, obs_dim=D, p_done=0.05, max_episode_steps=25)
    obs = torch.zeros(B, D, device=DEV)
    rew = torch.zeros(B, device=DEV)
    done = torch.zeros(B, dtype=torch.uint8, device=DEV)
    st = K.EpisodeStats(B, DEV)
    K.env_mujoco_synth_step(obs, None, None, st, seed, 0, reset=True)
    np.testing.assert_allclose(obs.cpu().numpy(), ref.reset(), rtol=1e-4, atol=1e-5)
    for t in range(T):
        K.env_mujoco_synth_step(obs, rew, done, st, seed, t, p_done=0.05, max_episode_steps=25)
        o, r, d = ref.step()
        np.testing.assert_allclose(obs.cpu().numpy(), o, rtol=1e-4, atol=1e-5)
        assert np.array_equal(rew.cpu().numpy(), r)
        assert np.array_equal(done.cpu().numpy().astype(bool), d)
    assert int(st.totals[0].item()) == len(ref.completed)
    x = torch.zeros(100000, D, device=DEV)
    K.env_mujoco_synth_step(x, None, None, K.EpisodeStats(100000, DEV), 3, 0, reset=True)
    assert abs(x.mean().item()) < 5e-3 and abs(x.std().item() - 1) < 5e-3


def test_cartpole_matches_twin():
    from parl_b200 import kernels as K
    B, T, seed = 256, 300, 21
    ref = oenv.CartPoleVec(B, seed, max_episode_steps=200)
    state = torch.zeros(B, 4, device=DEV)
    obs = torch.zeros(B, 4, device=DEV)
    rew = torch.zeros(B, device=DEV)
    done = torch.zeros(B, dtype=torch.uint8, device=DEV)
    st = K.EpisodeStats(B, DEV)
    K.env_cartpole_step(state, obs, None, None, None, st, seed, 0, reset=True)
    np.testing.assert_array_equal(obs.cpu().numpy(), ref.reset())
    rng = np.random.RandomState(0)
    mism = 0
    for t in range(T):
        a = rng.randint(0, 2, B).astype(np.int32)
        K.env_cartpole_step(state, obs, rew, done, torch.as_tensor(a).to(DEV), st, seed, t)
        o, r, d = ref.step(a)
        dd = done.cpu().numpy().astype(bool)
        if not np.array_equal(dd, d):       # sinf/cosf ulp differences may flip a threshold crossing: resync
            mism += int((dd != d).sum())
            ref.state = state.cpu().numpy().copy()
            ref.ep_len = st.ep_len.cpu().numpy().copy()
            ref.ep_ret = st.ep_ret.cpu().numpy().copy()
            continue
        np.testing.assert_allclose(obs.cpu().numpy(), o, rtol=1e-4, atol=1e-5)
    assert mism <= 2
    tot = st.totals.cpu().numpy()
    assert tot[0] > 100 and 15 < tot[2] / tot[0] < 35       # random policy: ~22 steps / episode


def test_gaussian_sampler():
    from parl_b200 import kernels as K
    N, D = 50000, 6
    mean = torch.randn(N, D, device=DEV)
    logstd = torch.tensor([-0.5, 0.0, 0.3, -1.0, 0.1, 0.2], device=DEV)
    a, lp = K.sample_gaussian(mean, logstd, 7, 3)
    ref_lp = torch.distributions.Normal(mean, logstd.exp().expand_as(mean)).log_prob(a).sum(1)
    np.testing.assert_allclose(lp.cpu().numpy(), ref_lp.cpu().numpy(), rtol=1e-4, atol=1e-4)
    z = ((a - mean) / logstd.exp()).cpu().numpy()
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    want = oenv.gaussians(np.arange(N, dtype=np.uint32), 3, D, *oph.split_seed(7), stream=oph.STREAM_GAUSS)
    np.testing.assert_allclose(z, want, rtol=1e-3, atol=1e-4)


def test_obs_gather_nhwc_and_space_to_depth():
    """bf16 NHWC and conv1 space-to-depth gathers against the uint8 NCHW gather (itself checked bit-exactly above)."""
    from parl_b200 import kernels as K
    B, HW, T, seed = 9, 84 * 84, 6, 5
    planes = torch.zeros(T + 4, B, HW, dtype=torch.uint8, device=DEV)
    ages = torch.zeros(T + 1, B, dtype=torch.uint8, device=DEV)
    rew = torch.zeros(B, device=DEV)
    done = torch.zeros(B, dtype=torch.uint8, device=DEV)
    st = K.EpisodeStats(B, DEV)
    K.env_atari_synth_step(planes[3], None, None, None, ages[0], st, seed, 0, reset=True)
    for t in range(T):
        K.env_atari_synth_step(planes[t + 4], rew, done, ages[t], ages[t + 1], st, seed, t, p_done=0.3)
    u8 = torch.empty(T * B, 4, 84, 84, dtype=torch.uint8, device=DEV)
    K.obs_stack_gather(planes, ages, 0, T, u8)
    ref = u8.cpu().float() * np.float32(1.0 / 255.0)                        # [N,4,84,84]
    nhwc = torch.empty(T * B, 84, 84, 4, dtype=torch.bfloat16, device=DEV)
    K.obs_stack_gather(planes, ages, 0, T, nhwc, scale=1.0 / 255.0)
    want = ref.permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(nhwc.cpu(), want)
    s2d = torch.empty(T * B, 21, 21, 64, dtype=torch.bfloat16, device=DEV)
    K.obs_stack_gather(planes, ages, 0, T, s2d, scale=1.0 / 255.0, s2d=True)
    pad = torch.zeros(T * B, 4, 85, 85)
    pad[:, :, 1:, 1:] = ref
    blocks = pad[:, :, :84, :84].reshape(T * B, 4, 21, 4, 21, 4)            # (n, c, Y, dy, X, dx)
    want = blocks.permute(0, 2, 4, 3, 5, 1).reshape(T * B, 21, 21, 64).to(torch.bfloat16)
    assert torch.equal(s2d.cpu(), want)
    # env-major order
    s2d_em = torch.empty(T * B, 21, 21, 64, dtype=torch.bfloat16, device=DEV)
    K.obs_stack_gather(planes, ages, 0, T, s2d_em, layout=K.ENV_MAJOR, scale=1.0 / 255.0, s2d=True)
    assert torch.equal(s2d_em.cpu().reshape(B, T, -1), want.reshape(T, B, -1).transpose(0, 1))


@pytest.mark.parametrize('B', [5, 300, 1200])
def test_fused_step_gather_is_bit_identical_to_separate_kernels(B):
    """rl_env_atari_synth_step_gather == rl_env_atari_synth_step followed by rl_obs_stack_gather(uint8 space-to-depth)
    of the next step: frames, scalars, actions, episode statistics and obs(t+1), over several steps with resets."""
    import torch
    from parl_b200 import kernels as K
    dev = torch.device('cuda', 0)
    T, A, seed = 6, 18, 1234

    def fresh():
        planes = torch.zeros((T + 4, B, 84 * 84), dtype=torch.uint8, device=dev)
        ages = torch.zeros((T + 1, B), dtype=torch.uint8, device=dev)
        st = K.EpisodeStats(B, dev)
        K.env_atari_synth_step(planes[3], None, None, None, ages[0], st, seed, 0, env_offset=7, reset=True)
        return dict(planes=planes, ages=ages, st=st, rew=torch.zeros((T, B), device=dev),
                    done=torch.zeros((T, B), dtype=torch.uint8, device=dev),
                    act=torch.zeros((T, B), dtype=torch.int32, device=dev),
                    obs=torch.zeros((T + 1, B, 21, 21, 64), dtype=torch.uint8, device=dev))

    g = torch.Generator(device=dev).manual_seed(B)
    logits = torch.randn((T, B, A), device=dev, generator=g)
    a, b = fresh(), fresh()
    step_dev = torch.arange(T, dtype=torch.int32, device=dev)
    for s_, fused in ((a, False), (b, True)):
        K.obs_stack_gather(s_['planes'], s_['ages'], 0, 1, s_['obs'][0], s2d=True)
        for t in range(T):
            if fused:
                K.env_atari_synth_step_gather(s_['planes'], t, s_['rew'][t], s_['done'][t], s_['ages'][t], s_['ages'][t + 1],
                                              s_['st'], seed, s_['obs'][t + 1], p_done=0.3, env_offset=7, logits=logits[t],
                                              actions_out=s_['act'][t], step_dev=step_dev[t:])
            else:
                K.env_atari_synth_step(s_['planes'][t + 4], s_['rew'][t], s_['done'][t], s_['ages'][t], s_['ages'][t + 1],
                                       s_['st'], seed, 0, p_done=0.3, env_offset=7, logits=logits[t],
                                       actions_out=s_['act'][t], step_dev=step_dev[t:])
                K.obs_stack_gather(s_['planes'], s_['ages'], t + 1, 1, s_['obs'][t + 1], s2d=True)
    torch.cuda.synchronize()
    for k in ('planes', 'ages', 'rew', 'done', 'act', 'obs'):
        assert torch.equal(a[k], b[k]), k
    assert a['done'].sum().item() > 0 and a['obs'].float().abs().sum().item() > 0
    assert torch.equal(a['st'].totals, b['st'].totals) and torch.equal(a['st'].ep_len, b['st'].ep_len)
