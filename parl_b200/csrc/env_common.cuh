// Device functions shared by the env steppers (env.cu) and the fused MLP rollout kernel (rollout_mlp.cu):
// episode bookkeeping, Box-Muller observation blocks, CartPole physics.  Same RNG contract (philox.cuh) and the
// same correctly-rounded single operations everywhere, so every caller produces bit-identical trajectories.
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace rl {

struct EpisodeStats {
  float* ep_ret;      // [B] running return
  int* ep_len;        // [B] running length
  float* totals;      // [4] completed episodes: count, sum return, sum length, (unused)
  float* ring_ret;    // [ring_cap] most recent completed returns
  int* ring_len;      // [ring_cap]
  unsigned* ring_head;
  int ring_cap;
};

// One warp-synchronous update of the episode bookkeeping for 32 envs.
__device__ __forceinline__ void episode_update(const EpisodeStats& s, int b, bool valid, float reward, bool done) {
  float ret = 0.f;
  int len = 0;
  if (valid) {
    ret = s.ep_ret[b] + reward;
    len = s.ep_len[b] + 1;
    s.ep_ret[b] = done ? 0.f : ret;
    s.ep_len[b] = done ? 0 : len;
  }
  const bool fin = valid && done;
  const unsigned mask = __ballot_sync(0xffffffffu, fin);
  if (mask == 0u) return;
  const int lane = threadIdx.x & 31;
  const float sret = warp_sum(fin ? ret : 0.f);
  const float slen = warp_sum(fin ? (float)len : 0.f);
  unsigned base = 0;
  if (lane == 0) {
    atomicAdd(s.totals + 0, (float)__popc(mask));
    atomicAdd(s.totals + 1, sret);
    atomicAdd(s.totals + 2, slen);
    if (s.ring_cap > 0) base = atomicAdd(s.ring_head, (unsigned)__popc(mask));
  }
  if (s.ring_cap > 0) {
    base = __shfl_sync(0xffffffffu, base, 0);
    if (fin) {
      const unsigned slot = (base + __popc(mask & ((1u << lane) - 1u))) % (unsigned)s.ring_cap;
      s.ring_ret[slot] = ret;
      s.ring_len[slot] = len;
    }
  }
}

__device__ __forceinline__ float u_open(uint32_t x) { return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-08f; }

__device__ __forceinline__ void gauss_block(uint32_t env, uint32_t n, uint32_t blk, uint32_t stream, uint32_t k0,
                                            uint32_t k1, float z[4]) {
  const uint4 x = philox4x32_10(env, n, blk, stream, k0, k1);
  const float r0 = sqrtf(-2.0f * logf(u_open(x.x))), th0 = 6.283185307179586f * u_open(x.y);
  const float r1 = sqrtf(-2.0f * logf(u_open(x.z))), th1 = 6.283185307179586f * u_open(x.w);
  z[0] = r0 * cosf(th0), z[1] = r0 * sinf(th0), z[2] = r1 * cosf(th1), z[3] = r1 * sinf(th1);
}

inline EpisodeStats make_episode_stats(float* ep_ret, int* ep_len, float* totals, float* ring_ret, int* ring_len,
                                       unsigned* ring_head, int ring_cap) {
  EpisodeStats s;
  s.ep_ret = ep_ret, s.ep_len = ep_len, s.totals = totals, s.ring_ret = ring_ret, s.ring_len = ring_len;
  s.ring_head = ring_head, s.ring_cap = (ring_ret && ring_len && ring_head) ? ring_cap : 0;
  return s;
}

// probability -> threshold on a uniform 32-bit draw (done = x < thr)
inline uint32_t prob_threshold(float p) {
  double v = (double)p * 4294967296.0;
  if (v < 0) v = 0;
  if (v > 4294967295.0) v = 4294967295.0;
  return (uint32_t)v;
}

// ---------------------------------------------------------------------------
// VecNormalizeEnv on the device (parl/env/mujoco_wrappers.py:95-168): every env keeps ITS OWN running statistics
// (benchmark/torch/ppo/env_utils.py wraps each env separately) fed one sample per step, float64 like the reference.
//   update_from_moments with batch (x, var 0, count 1)                      (:74-92, :191-217)
// ---------------------------------------------------------------------------
struct VecNormState {
  double* ob_mean;     // [B, D]
  double* ob_var;      // [B, D]
  double* ob_count;    // [B]
  double* ret;         // [B] discounted return accumulator
  double* ret_mean;    // [B]
  double* ret_var;     // [B]
  double* ret_count;   // [B]
  double clipob, cliprew, gamma, eps;
  int update;          // training: observation statistics follow the data (env.train() / env.eval())
  int norm_ob, norm_ret;
};

__device__ __forceinline__ void rms_update1(double& mean, double& var, double count, double x) {
  const double delta = x - mean;
  const double tot = count + 1.0;
  mean = mean + delta * 1.0 / tot;
  const double m2 = var * count + 0.0 + delta * delta * count * 1.0 / tot;
  var = m2 / tot;
}

// reward side of VecNormalizeEnv.step for env b: ret = ret*gamma + rew; ret_rms.update(ret);
// rew / sqrt(var + eps) clipped; ret = 0 when the episode ended.  Returns the normalised reward.
__device__ __forceinline__ float vecnorm_reward(const VecNormState& v, int b, float rew, bool done) {
  double ret = v.ret[b] * v.gamma + (double)rew;
  double out = (double)rew;
  if (v.norm_ret) {
    double m = v.ret_mean[b], s = v.ret_var[b];
    const double c = v.ret_count[b];
    rms_update1(m, s, c, ret);
    v.ret_mean[b] = m, v.ret_var[b] = s, v.ret_count[b] = c + 1.0;
    out = fmin(fmax(out / sqrt(s + v.eps), -v.cliprew), v.cliprew);
  }
  v.ret[b] = done ? 0.0 : ret;
  return (float)out;
}

// statistics-only pass over one observation of env b (the terminal observation of a finished episode)
__device__ __forceinline__ void vecnorm_obs_absorb(const VecNormState& v, int b, int D, const float* __restrict__ x,
                                                   int x_stride) {
  if (!(v.norm_ob && v.update)) return;
  const double c = v.ob_count[b];
  for (int d = 0; d < D; ++d) {
    double m = v.ob_mean[(size_t)b * D + d], s = v.ob_var[(size_t)b * D + d];
    rms_update1(m, s, c, (double)x[d * x_stride]);
    v.ob_mean[(size_t)b * D + d] = m, v.ob_var[(size_t)b * D + d] = s;
  }
  v.ob_count[b] = c + 1.0;
}

// one component of _obfilt: statistics update of dimension d with sample xv given the count BEFORE this observation,
// returns the filtered value.  (The caller bumps ob_count[b] once per observation.)
__device__ __forceinline__ float vecnorm_obs_dim(const VecNormState& v, int b, int D, int d, double count_old, float x) {
  double m = v.ob_mean[(size_t)b * D + d], s = v.ob_var[(size_t)b * D + d];
  const double xv = (double)x;
  if (v.update) {
    rms_update1(m, s, count_old, xv);
    v.ob_mean[(size_t)b * D + d] = m, v.ob_var[(size_t)b * D + d] = s;
  }
  return (float)fmin(fmax((xv - m) / sqrt(s + v.eps), -v.clipob), v.clipob);
}

// _obfilt of one observation of env b, in place on x (stride x_stride between its D components)
__device__ __forceinline__ void vecnorm_obs_filter(const VecNormState& v, int b, int D, float* __restrict__ x, int x_stride) {
  if (!v.norm_ob) return;
  const double c = v.ob_count[b];
  const bool upd = v.update != 0;
  for (int d = 0; d < D; ++d) {
    double m = v.ob_mean[(size_t)b * D + d], s = v.ob_var[(size_t)b * D + d];
    const double xv = (double)x[d * x_stride];
    if (upd) {
      rms_update1(m, s, c, xv);
      v.ob_mean[(size_t)b * D + d] = m, v.ob_var[(size_t)b * D + d] = s;
    }
    x[d * x_stride] = (float)fmin(fmax((xv - m) / sqrt(s + v.eps), -v.clipob), v.clipob);
  }
  if (upd) v.ob_count[b] = c + 1.0;
}

// gym classic_control/cartpole.py (third party, restated: SURVEY.md 8c item 4): one Euler step of the state
// s = (x, x_dot, theta, theta_dot) under action a; returns done.
__device__ __forceinline__ bool cartpole_physics(float4& s, int a) {
  const float force = a == 1 ? 10.0f : -10.0f;
  const float total_mass = 1.1f, pml = 0.05f, tau = 0.02f;
  const float c = cosf(s.z), sn = sinf(s.z);
  const float temp = __fdiv_rn(__fadd_rn(force, __fmul_rn(__fmul_rn(__fmul_rn(pml, s.w), s.w), sn)), total_mass);
  const float den = __fmul_rn(0.5f, __fsub_rn(1.3333334f, __fdiv_rn(__fmul_rn(__fmul_rn(0.1f, c), c), total_mass)));
  const float thacc = __fdiv_rn(__fsub_rn(__fmul_rn(9.8f, sn), __fmul_rn(c, temp)), den);
  const float xacc = __fsub_rn(temp, __fdiv_rn(__fmul_rn(__fmul_rn(pml, thacc), c), total_mass));
  s.x = __fadd_rn(s.x, __fmul_rn(tau, s.y));
  s.y = __fadd_rn(s.y, __fmul_rn(tau, xacc));
  s.z = __fadd_rn(s.z, __fmul_rn(tau, s.w));
  s.w = __fadd_rn(s.w, __fmul_rn(tau, thacc));
  return s.x < -2.4f || s.x > 2.4f || s.z < -0.20943951f || s.z > 0.20943951f;
}

// reset state ~ U(-0.05, 0.05)^4 from the observation stream, counter (env, n)
__device__ __forceinline__ float4 cartpole_reset_state(uint32_t env, uint32_t n, uint32_t k0, uint32_t k1) {
  const uint4 x = philox4x32_10(env, n, 0u, STREAM_OBS, k0, k1);
  float4 s;
  s.x = __fmul_rn(__fsub_rn(u01_24(x.x), 0.5f), 0.1f);
  s.y = __fmul_rn(__fsub_rn(u01_24(x.y), 0.5f), 0.1f);
  s.z = __fmul_rn(__fsub_rn(u01_24(x.z), 0.5f), 0.1f);
  s.w = __fmul_rn(__fsub_rn(u01_24(x.w), 0.5f), 0.1f);
  return s;
}

}  // namespace rl
