// K3 (scan part) — GAE / n-step return scans over a (T,B) rollout, backward in time.
//
// Reference arithmetic being replaced (PaddlePaddle/PARL):
//   benchmark/torch/ppo/storage.py:45-64   RolloutStorage.compute_returns (float32,
//                                          pre-step dones: mask with dones[t+1])
//   parl/utils/rl_utils.py:34-51 + benchmark/torch/a2c/actor.py:82-102
//                                          calc_gae per episode segment (float64,
//                                          post-step dones: segment ends AT done)
// One lane per env column, coalesced over b for every row; the recurrence uses the
// reference's operation order with separately rounded multiply/add so the float32
// variant reproduces RolloutStorage bit for bit.  Bytes per (t,b): read 12, write 8.
#include "common.cuh"

namespace rl {

// PPO convention.  adv_t = delta_t + (gamma*lambda) * nnt_t * adv_{t+1},
// delta_t = r_t + gamma * V_{t+1} * nnt_t - V_t,  nnt_t = 1 - dones[t+1] (last: 1 - last_done).
__global__ void __launch_bounds__(128) gae_scan_ppo_kernel(const float* __restrict__ rewards,
                                                          const float* __restrict__ values,
                                                          const float* __restrict__ dones,
                                                          const float* __restrict__ last_value,
                                                          const float* __restrict__ last_done, int T, int B,
                                                          float gamma, float gamma_lambda, float* __restrict__ adv,
                                                          float* __restrict__ ret) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float next_v = last_value[b];
  float nnt = __fsub_rn(1.0f, last_done[b]);
  float last = 0.f;
  constexpr int U = 8;                      // rows of loads in flight per lane
  for (int t1 = T; t1 > 0; t1 -= U) {
    const int n = min(U, t1);
    float r[U], v[U], d[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (k < n) {
        const long long g = (long long)(t1 - 1 - k) * B + b;
        r[k] = __ldcs(rewards + g), v[k] = __ldcs(values + g), d[k] = __ldcs(dones + g);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (k < n) {
        const long long g = (long long)(t1 - 1 - k) * B + b;
        // delta = rewards[t] + gamma * nextvalues * nextnonterminal - values[t]        storage.py:57-58
        const float delta = __fsub_rn(__fadd_rn(r[k], __fmul_rn(__fmul_rn(gamma, next_v), nnt)), v[k]);
        // lastgaelam = delta + gamma * gae_lambda * nextnonterminal * lastgaelam       :59-60
        last = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lambda, nnt), last));
        __stcs(adv + g, last);
        __stcs(ret + g, __fadd_rn(last, v[k]));                                        // :61
        next_v = v[k];
        nnt = __fsub_rn(1.0f, d[k]);        // dones[t] masks the step t-1 -> t transition
      }
    }
  }
}

// A2C convention (segment GAE in float64): a segment ends at done_t (next value 0) or at the
// rollout end (next value = bootstrap).  adv = lfilter recursion y_t = td_t + gamma*lam*y_{t+1}.
__global__ void __launch_bounds__(128) gae_scan_a2c_kernel(const float* __restrict__ rewards,
                                                          const float* __restrict__ values,
                                                          const uint8_t* __restrict__ dones,
                                                          const float* __restrict__ bootstrap, int T, int B,
                                                          double gamma, double lam, float* __restrict__ adv,
                                                          float* __restrict__ target) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double next_v = (double)bootstrap[b];
  double y = 0.0;
  const double gl = gamma * lam;
  for (int t = T - 1; t >= 0; --t) {
    const long long g = (long long)t * B + b;
    const double v = (double)values[g];
    if (dones[g]) next_v = 0.0, y = 0.0;                     // actor.py:84-87: next_value = 0 at done
    const double td = (double)rewards[g] + gamma * next_v - v;   // rl_utils.py:49
    y = td + gl * y;                                         // rl_utils.py:31 (IIR on reversed sequence)
    adv[g] = (float)y;
    target[g] = (float)(y + v);                              // actor.py:94
    next_v = v;
  }
}

}  // namespace rl

using namespace rl;

extern "C" int rl_gae_scan(const float* rewards, const float* values, const float* dones, const float* last_value,
                           const float* last_done, int T, int B, float gamma, float gae_lambda, float* advantages,
                           float* returns, rl_stream_t stream) {
  RL_CHECK_ARG(rewards && values && dones && last_value && last_done && advantages && returns, "gae_scan: null pointer");
  RL_CHECK_ARG(T > 0 && B > 0, "gae_scan: bad shape");
  // the reference forms gamma * gae_lambda in Python float (double) before it meets float32 data
  const float gl = (float)((double)gamma * (double)gae_lambda);
  gae_scan_ppo_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards, values, dones, last_value, last_done,
                                                                         T, B, gamma, gl, advantages, returns);
  RL_CHECK_LAUNCH("rl_gae_scan");
  return RL_OK;
}

extern "C" int rl_gae_scan_segments(const float* rewards, const float* values, const uint8_t* dones,
                                    const float* bootstrap_value, int T, int B, double gamma, double lam,
                                    float* advantages, float* target_values, rl_stream_t stream) {
  RL_CHECK_ARG(rewards && values && dones && bootstrap_value && advantages && target_values,
               "gae_scan_segments: null pointer");
  RL_CHECK_ARG(T > 0 && B > 0, "gae_scan_segments: bad shape");
  gae_scan_a2c_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rewards, values, dones, bootstrap_value, T, B,
                                                                         gamma, lam, advantages, target_values);
  RL_CHECK_LAUNCH("rl_gae_scan_segments");
  return RL_OK;
}
