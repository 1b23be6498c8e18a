"""CPU tests of the env layer against fixtures RECORDED FROM THE REFERENCE (tests/golden/make_golden_env.py):
  * oracle/vecnorm.py and the product host wrapper parl_b200.env.mujoco_wrappers.VecNormalizeEnv reproduce the
    reference VecNormalizeEnv trace (observations, rewards, final running statistics);
  * parl_b200.env.atari_wrappers.wrap_deepmind reproduces the reference chain on the scripted raw env (newest frames
    bit for bit, whole stacks by CRC32, rewards, dones, monitor statistics);
  * oracle/envs.py (the CPU twin the device env kernels are bit-exact against) draws from the distributions of the
    reference's mock gym envs (parl/tests/gym.py:105-213): uniform bytes on {0..254}, Bernoulli rewards / dones,
    unit-normal observations — the moments match the recorded reference draws within sampling error."""
import zlib

import numpy as np

from oracle import envs as oenv
from oracle import vecnorm as ovn


def _scripted_env(g, b):
    obs_seq, term_seq, rew_seq, done_seq = g['obs_seq'], g['term_seq'], g['rew_seq'], g['done_seq']

    class Space(object):
        shape = (obs_seq.shape[2], )

    class Scripted(object):
        observation_space = Space()

        def __init__(self):
            self.t = 0

        def reset(self):
            return obs_seq[self.t, b].copy()

        def step(self, action):
            t = self.t
            d = bool(done_seq[t, b])
            ob = term_seq[t, b].copy() if d else obs_seq[t + 1, b].copy()
            self.t += 1
            return ob, float(rew_seq[t, b]), d, {}
    return Scripted()


def test_vecnormalize_oracle_matches_reference_trace(golden):
    g = golden('vecnormalize')
    T, B, D = g['term_seq'].shape
    st = ovn.VecNormState(B, D)
    ob0 = ovn.obfilt(st, g['obs_seq'][0])
    np.testing.assert_allclose(ob0, g['ob0'], rtol=1e-12, atol=1e-12)
    for t in range(T):
        ob, rew = ovn.step(st, g['obs_seq'][t + 1], g['rew_seq'][t], g['done_seq'][t], terminal_ob=g['term_seq'][t])
        np.testing.assert_allclose(ob, g['ob_out'][t], rtol=1e-10, atol=1e-12, err_msg='t=%d' % t)
        np.testing.assert_allclose(rew, g['rew_out'][t], rtol=1e-10, atol=1e-12)
    for k in ('ob_mean', 'ob_var', 'ob_count', 'ret_mean', 'ret_var', 'ret_count'):
        np.testing.assert_allclose(getattr(st, k), g[k], rtol=1e-10, atol=1e-12, err_msg=k)


def test_product_vecnormalize_wrapper_matches_reference_trace(golden):
    from parl_b200.env.mujoco_wrappers import VecNormalizeEnv
    g = golden('vecnormalize')
    T, B, D = g['term_seq'].shape
    envs = [VecNormalizeEnv(_scripted_env(g, b)) for b in range(B)]
    np.testing.assert_allclose(np.stack([e.reset() for e in envs]), g['ob0'], rtol=1e-12)
    for t in range(T):
        for b, e in enumerate(envs):
            ob, r, d, _ = e.step(None)
            if d:
                ob = e.reset()
            np.testing.assert_allclose(ob, g['ob_out'][t, b], rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(np.asarray(r).reshape(-1)[0], g['rew_out'][t, b], rtol=1e-12)
    np.testing.assert_allclose(np.stack([e.ob_rms.var for e in envs]), g['ob_var'], rtol=1e-12)
    np.testing.assert_allclose(np.array([e.ret_rms.count for e in envs]), g['ret_count'], rtol=1e-12)


def test_wrap_deepmind_matches_reference_chain(golden):
    from parl_b200.env import atari_wrappers as aw
    g = golden('wrap_deepmind')
    S = int(g['script_len'])
    rs = np.random.RandomState(int(g['script_seed']))
    raw_frames = rs.randint(0, 255, (S, 210, 160, 3)).astype(np.uint8)
    raw_rew = rs.choice([-2.0, 0.0, 1.0, 3.0], S)
    raw_done = rs.rand(S) < 0.06
    raw_lives = rs.randint(0, 5, S)

    class Space(object):
        shape, dtype = (210, 160, 3), 'uint8'

    class Spec(object):
        id = 'PongNoFrameskip-v4'

    class RawAtari(object):
        observation_space, spec = Space(), Spec()
        _max_episode_steps = 1000

        def __init__(self):
            self.i = 0
            outer = self

            class Lives(object):
                def lives(self):
                    return int(raw_lives[(outer.i - 1) % S])

            class Ale(object):
                ale = Lives()
                np_random = np.random.RandomState(int(g['noop_seed']))

                def get_action_meanings(self):
                    return ['NOOP'] * 6
            self.unwrapped = Ale()

        def _next(self):
            k = self.i % S
            self.i += 1
            return k

        def reset(self):
            return raw_frames[self._next()].copy()

        def step(self, action):
            k = self._next()
            return raw_frames[k].copy(), float(raw_rew[k]), bool(raw_done[k]), {}

        def seed(self, v):
            pass

    env = aw.wrap_deepmind(RawAtari(), dim=84, obs_format='NCHW')
    obs = [env.reset()]
    rew, done = [], []
    for t in range(len(g['rew'])):
        o, r, d, _ = env.step(t % 6)
        if d:
            o = env.reset()
        obs.append(o), rew.append(r), done.append(d)
    obs = np.stack(obs).astype(np.uint8)
    assert obs.shape == (len(g['rew']) + 1, 4, 84, 84)
    np.testing.assert_array_equal(np.array(rew), g['rew'])
    np.testing.assert_array_equal(np.array(done), g['done'])
    np.testing.assert_array_equal(obs[:, -1], g['newest'])
    assert [zlib.crc32(o.tobytes()) for o in obs] == list(g['obs_crc'])
    mon = aw.get_wrapper_by_cls(env, aw.MonitorEnv)
    np.testing.assert_array_equal(np.array(mon.get_episode_rewards(), np.float64), g['episode_rewards'])
    np.testing.assert_array_equal(np.array(mon.get_episode_lengths()), g['episode_lengths'])
    assert aw.get_wrapper_by_cls(env, aw.FireResetEnv) is None        # all-NOOP action meanings: no FIRE wrapper


def test_env_twin_draws_match_the_reference_mock_distributions(golden):
    """oracle/envs.py (and so the device kernels, bit-exact against it) vs draws of the reference mocks."""
    g = golden('mock_env_draws')
    # --- Atari-shaped: bytes uniform on {0..254} (randint(0,255): high exclusive, parl/tests/gym.py:165)
    env = oenv.AtariSynthVec(64, 11, p_done=0.1)
    env.reset()
    pix, rews, dones = [], [], []
    for _ in range(200):
        o, r, d = env.step()
        pix.append(o[:, -1].reshape(-1)), rews.append(r), dones.append(d)
    pix = np.concatenate(pix)
    rews, dones = np.concatenate(rews), np.concatenate(dones)
    ref_hist = g['pong_hist'].astype(np.float64)
    assert ref_hist[255] == 0 and pix.max() <= 254 and pix.min() == 0
    p_ref, p_our = ref_hist / ref_hist.sum(), np.bincount(pix, minlength=256) / pix.size
    assert np.abs(p_ref[:255] - 1 / 255.0).max() < 2e-4
    # KNOWN, DOCUMENTED deviation: one Philox byte per pixel mapped by max(b,1)-1 — 256 byte values
    # onto 255 pixel values, so pixel value 0 has probability 2/256 and every other value 1/256 instead of 1/255 each
    # (total-variation distance 0.0039 from the reference's uniform law, pixel mean 126.5 instead of 127).
    want = np.full(255, 1 / 256.0)
    want[0] = 2 / 256.0
    assert np.abs(p_our[:255] - want).max() < 2e-4
    assert 0.5 * np.abs(p_our - p_ref).sum() < 0.006
    assert abs(pix.mean() - 126.5) < 0.2 and abs(np.arange(256) @ p_ref - 127.0) < 0.2
    n = rews.size
    for ours, ref, p in ((rews.mean(), g['pong_reward_mean'], 0.5), (dones.mean(), g['pong_done_mean'], 0.1)):
        se = np.sqrt(p * (1 - p) * (1.0 / n + 1.0 / float(g['pong_n'])))
        assert abs(ours - ref) < 5 * se and abs(ours - p) < 5 * np.sqrt(p * (1 - p) / n)
    assert set(np.unique(rews)) == set(g['pong_reward_values'])
    # --- MuJoCo-shaped: obs ~ N(0,1)^17, reward in {0,1}, done p = 0.01 (gym.py:196-200)
    mj = oenv.MujocoSynthVec(256, 5, p_done=0.01)
    mj.reset()
    ob, rw, dn = [], [], []
    for _ in range(100):
        o, r, d = mj.step()
        ob.append(o), rw.append(r), dn.append(d)
    ob, rw, dn = np.concatenate(ob), np.concatenate(rw), np.concatenate(dn)
    m = ob.shape[0]
    assert np.abs(ob.mean(0)).max() < 5 / np.sqrt(m) and np.abs(g['cheetah_mean']).max() < 5 / np.sqrt(float(g['cheetah_n']))
    assert np.abs(ob.var(0) - 1).max() < 5 * np.sqrt(2.0 / m) and np.abs(g['cheetah_var'] - 1).max() < 0.06
    kurt = ((ob - ob.mean(0)) ** 4).mean(0) / ob.var(0) ** 2
    assert np.abs(kurt - 3).max() < 0.35 and np.abs(g['cheetah_kurt'] - 3).max() < 0.35
    assert abs(rw.mean() - g['cheetah_reward_mean']) < 0.03 and abs(dn.mean() - g['cheetah_done_mean']) < 0.004
