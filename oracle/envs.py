"""Synthetic vectorised envs — CPU twin of parl_b200/csrc/env.cu (TEST INFRASTRUCTURE).

Distributions follow the reference's mock gym (parl/tests/gym.py:105-213):
  PongEnv        obs U{0..254} uint8 frames, reward in {0,1} p=.5, done p=.1 (:163-169)
  HalfCheetahEnv obs N(0,1)^17, reward in {0,1} p=.5, done p=.01            (:198-203)
  CartPoleEnv    (mock: random) — here the real gym classic-control physics,
                 restated from the published constants.
Auto-reset follows parl/env/vector_env.py:53-63 (on done the returned obs is the
reset obs, done=True and the terminal reward are kept) and the frame stack follows
FrameStack (parl/env/atari_wrappers.py:270-307: reset fills all k slots with the
reset frame).  The reference draws from unseeded np.random; the contract here is
the counter-based Philox stream of oracle/philox.py ("parity unpinned" w.r.t. the
reference, bit-exact w.r.t. the device kernels for uint8/bool/int outputs).
"""
import numpy as np

from . import philox as ph


class AtariSynthVec(object):
    """B envs; frames [HW] uint8; obs = 4-frame stack (oldest first), NCHW."""

    def __init__(self, B, seed, hw=84 * 84, p_done=0.1, stack=4, env_offset=0):
        assert hw % 16 == 0
        self.B, self.hw, self.stack = B, hw, stack
        self.k0, self.k1 = ph.split_seed(seed)
        self.ids = (np.arange(B, dtype=np.uint64) + np.uint64(env_offset)).astype(np.uint32)
        self.thr = ph.prob_threshold(p_done)
        self.step_count = 0
        self.ep_ret = np.zeros(B, np.float32)
        self.ep_len = np.zeros(B, np.int32)
        self.completed = []          # (return, length) in (step, env) order

    def gen_frame(self, n):
        blk = np.arange(self.hw // 16, dtype=np.uint32)
        x = ph.philox4x32(self.ids[:, None], np.uint32(n), blk[None, :], ph.STREAM_FRAME, self.k0, self.k1)
        words = np.stack(x, axis=-1).astype('<u4')                       # [B, nblk, 4]
        b = words.view(np.uint8).reshape(self.B, self.hw)
        return (np.maximum(b, 1) - 1).astype(np.uint8)                   # U{0..254}

    def reset(self):
        f = self.gen_frame(0)
        self.frames = [f] * self.stack
        self.age = np.zeros(self.B, np.uint8)
        return self.obs()

    def obs(self):
        return np.stack(self.frames, axis=1)                             # [B, 4, HW]

    def step(self, actions=None):
        s = self.step_count
        x0, x1, _, _ = ph.philox4x32(self.ids, np.uint32(s), 0, ph.STREAM_REWDONE, self.k0, self.k1)
        reward = (x0 & np.uint32(1)).astype(np.float32)
        done = x1 < np.uint32(self.thr)
        newf = self.gen_frame(s + 1)
        self.ep_ret += reward
        self.ep_len += 1
        for b in np.nonzero(done)[0]:
            self.completed.append((float(self.ep_ret[b]), int(self.ep_len[b])))
        self.ep_ret[done] = 0
        self.ep_len[done] = 0
        # non-done: shift in the new frame; done: all 4 slots = reset frame (atari_wrappers.py:290-294)
        shifted = self.frames[1:] + [newf]
        d = done[:, None]
        self.frames = [np.where(d, newf, fr) for fr in shifted]
        self.age = np.where(done, 0, np.minimum(self.age + 1, self.stack - 1)).astype(np.uint8)
        self.step_count += 1
        return self.obs(), reward, done


def gaussians(ids, n, D, k0, k1, stream=ph.STREAM_OBS):
    """[len(ids), D] float32 N(0,1) via Box-Muller on Philox words (float tolerance contract)."""
    nblk = (D + 3) // 4
    blk = np.arange(nblk, dtype=np.uint32)
    x = ph.philox4x32(np.asarray(ids, np.uint32)[:, None], np.uint32(n), blk[None, :], stream, k0, k1)
    def u(v):
        return ((v >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)
    out = []
    for a, b in ((x[0], x[1]), (x[2], x[3])):
        r = np.sqrt(np.float32(-2.0) * np.log(u(a)))
        th = np.float32(6.283185307179586) * u(b)
        out += [r * np.cos(th), r * np.sin(th)]
    z = np.stack(out, axis=-1).reshape(len(ids), nblk * 4)               # per block: z0 z1 z2 z3
    return z[:, :D].astype(np.float32)


class MujocoSynthVec(object):
    def __init__(self, B, seed, obs_dim=17, p_done=0.01, max_episode_steps=0, env_offset=0):
        self.B, self.D = B, obs_dim
        self.k0, self.k1 = ph.split_seed(seed)
        self.ids = (np.arange(B, dtype=np.uint64) + np.uint64(env_offset)).astype(np.uint32)
        self.thr = ph.prob_threshold(p_done)
        self.max_steps = max_episode_steps
        self.step_count = 0
        self.ep_ret = np.zeros(B, np.float32)
        self.ep_len = np.zeros(B, np.int32)
        self.completed = []

    def reset(self):
        return gaussians(self.ids, 0, self.D, self.k0, self.k1)

    def step(self, actions=None):
        s = self.step_count
        x0, x1, _, _ = ph.philox4x32(self.ids, np.uint32(s), 0, ph.STREAM_REWDONE, self.k0, self.k1)
        reward = (x0 & np.uint32(1)).astype(np.float32)
        done = x1 < np.uint32(self.thr)
        self.ep_ret += reward
        self.ep_len += 1
        if self.max_steps:
            done = done | (self.ep_len >= self.max_steps)
        for b in np.nonzero(done)[0]:
            self.completed.append((float(self.ep_ret[b]), int(self.ep_len[b])))
        self.ep_ret[done] = 0
        self.ep_len[done] = 0
        obs = gaussians(self.ids, s + 1, self.D, self.k0, self.k1)
        self.step_count += 1
        return obs, reward, done


class CartPoleVec(object):
    """gym classic-control CartPole (v0: 200-step limit, v1: 500), float32, Euler."""
    GRAVITY, MASSCART, MASSPOLE, LENGTH, FORCE_MAG, TAU = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    THETA_LIMIT = np.float32(12 * 2 * np.pi / 360)
    X_LIMIT = np.float32(2.4)

    def __init__(self, B, seed, max_episode_steps=200, env_offset=0):
        self.B = B
        self.k0, self.k1 = ph.split_seed(seed)
        self.ids = (np.arange(B, dtype=np.uint64) + np.uint64(env_offset)).astype(np.uint32)
        self.max_steps = max_episode_steps
        self.step_count = 0
        self.ep_ret = np.zeros(B, np.float32)
        self.ep_len = np.zeros(B, np.int32)
        self.completed = []

    def _reset_state(self, n):
        x = ph.philox4x32(self.ids, np.uint32(n), 0, ph.STREAM_OBS, self.k0, self.k1)
        u = np.stack([ph.u01_24(v) for v in x], axis=-1)
        return ((u - np.float32(0.5)) * np.float32(0.1)).astype(np.float32)      # U(-0.05, 0.05)

    def reset(self):
        self.state = self._reset_state(0)
        return self.state.copy()

    def step(self, actions):
        f32 = np.float32
        a = np.asarray(actions).astype(np.int64)
        x, x_dot, th, th_dot = [self.state[:, i] for i in range(4)]
        force = np.where(a == 1, f32(self.FORCE_MAG), f32(-self.FORCE_MAG)).astype(f32)
        total_mass = f32(self.MASSPOLE + self.MASSCART)
        pml = f32(self.MASSPOLE * self.LENGTH)
        c, s = np.cos(th).astype(f32), np.sin(th).astype(f32)
        temp = (force + pml * th_dot * th_dot * s) / total_mass
        thacc = (f32(self.GRAVITY) * s - c * temp) / (f32(self.LENGTH) * (f32(4.0 / 3.0) - f32(self.MASSPOLE) * c * c / total_mass))
        xacc = temp - pml * thacc * c / total_mass
        tau = f32(self.TAU)
        x = x + tau * x_dot
        x_dot = x_dot + tau * xacc
        th = th + tau * th_dot
        th_dot = th_dot + tau * thacc
        st = np.stack([x, x_dot, th, th_dot], axis=-1).astype(f32)
        done = (x < -self.X_LIMIT) | (x > self.X_LIMIT) | (th < -self.THETA_LIMIT) | (th > self.THETA_LIMIT)
        reward = np.ones(self.B, f32)
        self.ep_ret += reward
        self.ep_len += 1
        if self.max_steps:
            done = done | (self.ep_len >= self.max_steps)
        for b in np.nonzero(done)[0]:
            self.completed.append((float(self.ep_ret[b]), int(self.ep_len[b])))
        self.ep_ret[done] = 0
        self.ep_len[done] = 0
        rs = self._reset_state(self.step_count + 1)
        self.state = np.where(done[:, None], rs, st).astype(f32)
        self.step_count += 1
        return self.state.copy(), reward, done
