"""``parl.env.compat_wrappers`` — surface of parl/env/compat_wrappers.py:16-99: ``CompatWrapper`` gives every gym
version the old 4-tuple ``step`` / bare-observation ``reset`` API and applies the ``_max_episode_steps`` time limit;
``is_gym_version_ge`` compares the installed gym's version.  Works without gym installed (a minimal ``Wrapper`` base
with gym's attribute forwarding stands in), so the real-env bridge and the unmodified examples import it anywhere."""
__all__ = ['CompatWrapper', 'is_gym_version_ge', 'V_GYM_CHANGED', 'V_NPRANDOM_CHANGED', 'Wrapper']

V_GYM_CHANGED = '0.26.0'          # env.seed() / reset() / step() signatures changed
V_NPRANDOM_CHANGED = '0.22.0'     # env.np_random.randint -> .integers

try:                                                    # pragma: no cover - depends on the installation
    import gym as _gym
    _GymWrapper = _gym.Wrapper
except Exception:                                       # gym absent (this image): duck-typed base
    _gym = None

    class _GymWrapper(object):
        """Minimal gym.Wrapper: keeps ``env`` and forwards public attribute look-ups to it."""

        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith('_'):
                raise AttributeError("attempted to get missing private attribute '{}'".format(name))
            return getattr(self.env, name)

        @property
        def unwrapped(self):
            return getattr(self.env, 'unwrapped', self.env)

Wrapper = _GymWrapper


def _version_tuple(text):
    out = []
    for part in str(text).split('.'):
        try:
            out.append(int(part))
        except ValueError:
            out.append(part)
    return out


def is_gym_version_ge(compare_version):
    """True when the installed gym is at least ``compare_version`` (no gym -> treated as an old-API env source)."""
    installed = getattr(_gym, '__version__', None) if _gym is not None else None
    if installed is None:
        return False
    try:
        return _version_tuple(installed) >= _version_tuple(compare_version)
    except TypeError:
        return False


class CompatWrapper(Wrapper):
    def __init__(self, env):
        Wrapper.__init__(self, env)
        if hasattr(env, '_max_episode_steps'):
            self._max_episode_steps = int(env._max_episode_steps)
        if hasattr(env, '_elapsed_steps'):
            self._elapsed_steps = env._elapsed_steps
        self.count_ep_step = 0
        self.random_seed = 'without_setting'

    def reset(self, **kwargs):
        if is_gym_version_ge(V_GYM_CHANGED):
            if self.random_seed != 'without_setting':
                kwargs['seed'] = self.random_seed
            obs, _info = self.env.reset(**kwargs)
            return obs
        return self.env.reset(**kwargs)

    def seed(self, random_seed):
        if is_gym_version_ge(V_GYM_CHANGED):
            self.random_seed = random_seed
        else:
            self.env.seed(random_seed)

    def step(self, action):
        self.count_ep_step += 1
        if is_gym_version_ge(V_GYM_CHANGED):
            obs, reward, done, _truncated, info = self.env.step(action)
        else:
            obs, reward, done, info = self.env.step(action)
        if hasattr(self.env, '_elapsed_steps'):
            self._elapsed_steps = self.env._elapsed_steps
        if hasattr(self, '_max_episode_steps') and self.count_ep_step >= self._max_episode_steps:
            done = True
            self.count_ep_step = 0
        return obs, reward, done, info
