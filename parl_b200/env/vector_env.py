"""VectorEnv for host-side gym-API envs — contract of parl/env/vector_env.py:21-63: ``reset()``
returns the list of first observations, ``step(actions)`` steps every env and, when an env reports
done, replaces the returned observation by ``env.reset()`` while keeping done=True and the terminal
reward.  (The on-device pools in device_envs.py implement the same contract inside one kernel.)"""

__all__ = ['VectorEnv']


class VectorEnv(object):
    def __init__(self, envs):
        self.envs = envs
        self.envs_num = len(envs)

    def reset(self):
        return [env.reset() for env in self.envs]

    def step(self, actions):
        obs_batch, reward_batch, done_batch, info_batch = [], [], [], []
        for env, action in zip(self.envs, actions):
            obs, reward, done, info = env.step(action)
            if done:
                obs = env.reset()
            obs_batch.append(obs)
            reward_batch.append(reward)
            done_batch.append(done)
            info_batch.append(info)
        return obs_batch, reward_batch, done_batch, info_batch
