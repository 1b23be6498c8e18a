"""Golden fixtures for the env layer, recorded by RUNNING THE REFERENCE (build container only):

    python tests/golden/make_golden_env.py

  * vecnormalize.npz   parl/env/mujoco_wrappers.py VecNormalizeEnv (one instance per env, as
                       benchmark/torch/ppo/env_utils.py builds them) driven by scripted envs through ParallelEnv-style
                       stepping with auto-reset: normalised observations / rewards per step, final running statistics
  * mock_env_draws.npz draw statistics of the reference's mock gym envs (parl/tests/gym.py:105-213): per-pixel /
                       per-dim moments, value ranges, reward and done frequencies over N seeded steps — pins the
                       distributions oracle/envs.py and the device env kernels reproduce
  * wrap_deepmind.npz  parl/env/atari_wrappers.py wrap_deepmind(dim=84, 'NCHW') on a scripted raw env: stacked
                       observations, rewards, dones, monitor statistics — pins parl_b200/env/atari_wrappers.py
The reference's own mock gym (parl/tests/gym.py) stands in for gym, as in the reference's unit tests.
"""
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def _setup():
    stubs = tempfile.mkdtemp(prefix='parl_stubs_')
    with open(os.path.join(stubs, 'termcolor.py'), 'w') as f:
        f.write('def colored(s, *a, **k):\n    return s\n')
    with open(os.path.join(stubs, 'pyarrow.py'), 'w') as f:
        f.write('raise ImportError("hidden: forces cloudpickle path")\n')
    os.environ['PARL_BACKEND'] = 'torch'
    os.environ['PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION'] = 'python'
    os.environ['HOME'] = tempfile.mkdtemp(prefix='parl_home_')
    sys.path[:0] = [stubs, os.path.join(REF, 'parl', 'tests'), REF]
    import gym                                   # the reference's mock
    gym.__version__ = '0.21.0'
    return gym


def main():
    gym = _setup()
    import warnings
    warnings.filterwarnings('ignore')
    import numpy as np
    from parl.env.mujoco_wrappers import VecNormalizeEnv
    from parl.env import atari_wrappers as aw

    # ------------------------------------------------------------------ VecNormalizeEnv
    B, D, T = 5, 3, 40
    rng = np.random.RandomState(42)
    obs_seq = (rng.randn(T + 1, B, D) * np.array([1.0, 5.0, 0.2]) + np.array([0.0, 3.0, -1.0])).astype(np.float64)
    term_seq = (rng.randn(T, B, D) * 2.0).astype(np.float64)          # what step() returns when the episode ends
    rew_seq = rng.randn(T, B) * 3.0 + 1.0
    done_seq = rng.rand(T, B) < 0.2

    class Scripted(object):
        """Raw env of column b: step t returns the scripted terminal observation when done, else the next one."""

        def __init__(self, b):
            self.b, self.t = b, 0
            self.observation_space = gym.Box(high=None, low=None, shape=(D, ), dtype=None)

        def reset(self):
            return obs_seq[self.t, self.b].copy()

        def step(self, action):
            t, b = self.t, self.b
            d = bool(done_seq[t, b])
            ob = term_seq[t, b].copy() if d else obs_seq[t + 1, b].copy()
            self.t += 1
            return ob, float(rew_seq[t, b]), d, {}

    envs = [VecNormalizeEnv(Scripted(b)) for b in range(B)]
    ob0 = np.stack([e.reset() for e in envs])
    ob_out, rew_out = np.zeros((T, B, D)), np.zeros((T, B))
    for t in range(T):
        for b, e in enumerate(envs):
            ob, r, d, _ = e.step(None)
            if d:                                 # ParallelEnv.step: reset on done, hand on the reset observation
                ob = e.reset()
            ob_out[t, b], rew_out[t, b] = ob, np.asarray(r).reshape(-1)[0]
    np.savez(os.path.join(HERE, 'vecnormalize.npz'), obs_seq=obs_seq, term_seq=term_seq, rew_seq=rew_seq,
             done_seq=done_seq, ob0=ob0, ob_out=ob_out, rew_out=rew_out,
             ob_mean=np.stack([e.ob_rms.mean for e in envs]), ob_var=np.stack([e.ob_rms.var for e in envs]),
             ob_count=np.array([e.ob_rms.count for e in envs]),
             ret_mean=np.array([float(e.ret_rms.mean) for e in envs]), ret_var=np.array([float(e.ret_rms.var) for e in envs]),
             ret_count=np.array([e.ret_rms.count for e in envs]))

    # ------------------------------------------------------------------ mock env draw statistics
    np.random.seed(2024)
    pong = gym.make('PongNoFrameskip-v4')
    n = 400
    fr = np.zeros((n, 210, 160, 3), np.uint8)
    rw, dn = np.zeros(n), np.zeros(n, bool)
    for i in range(n):
        fr[i], rw[i], dn[i], _ = pong.step(0)
    hist = np.bincount(fr.reshape(-1), minlength=256).astype(np.int64)
    cheetah = gym.make('HalfCheetah-v1')
    m = 20000
    ob = np.zeros((m, 17))
    rw2, dn2 = np.zeros(m), np.zeros(m, bool)
    for i in range(m):
        ob[i], rw2[i], dn2[i], _ = cheetah.step(np.zeros(6))
    cart = gym.make('CartPole-v0')
    oc = np.zeros((m, 4))
    rw3, dn3 = np.zeros(m), np.zeros(m, bool)
    for i in range(m):
        oc[i], rw3[i], dn3[i], _ = cart.step(0)
    np.savez(os.path.join(HERE, 'mock_env_draws.npz'),
             pong_hist=hist, pong_n_pixels=np.int64(fr.size), pong_reward_mean=rw.mean(), pong_done_mean=dn.mean(),
             pong_reward_values=np.unique(rw), pong_n=n,
             cheetah_mean=ob.mean(0), cheetah_var=ob.var(0), cheetah_kurt=((ob - ob.mean(0)) ** 4).mean(0) / ob.var(0) ** 2,
             cheetah_reward_mean=rw2.mean(), cheetah_done_mean=dn2.mean(), cheetah_n=m,
             cart_min=oc.min(0), cart_max=oc.max(0), cart_mean=oc.mean(0), cart_var=oc.var(0),
             cart_reward_mean=rw3.mean(), cart_done_mean=dn3.mean(), cart_n=m)

    # ------------------------------------------------------------------ wrap_deepmind on a scripted raw env
    S = 400
    rs = np.random.RandomState(7)
    raw_frames = rs.randint(0, 255, (S, 210, 160, 3)).astype(np.uint8)
    raw_rew = rs.choice([-2.0, 0.0, 1.0, 3.0], S)
    raw_done = rs.rand(S) < 0.06
    raw_lives = rs.randint(0, 5, S)

    class RawAtari(object):
        """Scripted stand-in for the mock PongNoFrameskip-v4: consumes one script row per raw step / reset."""

        def __init__(self):
            self.i = 0

            class Lives(object):
                def lives(_):
                    return int(raw_lives[(self.i - 1) % S])

            class Ale(object):
                ale = Lives()
                np_random = np.random.RandomState(5)

                def get_action_meanings(_):
                    return ['NOOP'] * 6
            self.unwrapped = Ale()
            self.observation_space = gym.Box(high=None, low=None, shape=(210, 160, 3), dtype='uint8')
            self.action_space = gym.ActionSpace(n=6, shape=())
            self.spec = gym.Spec('PongNoFrameskip-v4')
            self.metadata = {'render.modes': []}
            self.reward_range = [0, 1]
            self._max_episode_steps = 1000

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

        def close(self):
            pass

    env = aw.wrap_deepmind(RawAtari(), dim=84, obs_format='NCHW')
    obs_l, rew_l, done_l = [env.reset()], [], []
    for t in range(60):
        o, r, d, _ = env.step(t % 6)
        if d:
            o = env.reset()
        obs_l.append(o), rew_l.append(r), done_l.append(d)
    mon = aw.get_wrapper_by_cls(env, aw.MonitorEnv)
    import zlib
    obs_a = np.stack(obs_l).astype(np.uint8)
    # small fixture: the raw script is regenerated from its seed (RandomState(7), same call order) by the test; of the
    # outputs only the newest frame of every stacked observation travels, the full stacks are pinned by their CRC32
    np.savez_compressed(os.path.join(HERE, 'wrap_deepmind.npz'), script_seed=7, script_len=S, noop_seed=5,
                        newest=obs_a[:, -1], obs_crc=np.array([zlib.crc32(o.tobytes()) for o in obs_a], np.int64),
                        rew=np.array(rew_l), done=np.array(done_l),
                        episode_rewards=np.array(mon.get_episode_rewards(), np.float64),
                        episode_lengths=np.array(mon.get_episode_lengths(), np.int64))
    print('golden env fixtures written')


if __name__ == '__main__':
    main()
