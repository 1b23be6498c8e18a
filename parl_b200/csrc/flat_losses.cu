// K2/K3/K4 (flat, scan-free parts): A2C loss, PPO clipped-surrogate loss
// (categorical and diagonal-Gaussian), DQN/DDQN/PER TD loss — forward AND the
// gradient w.r.t. the network outputs in one launch each.
//
// Reference arithmetic being replaced (PaddlePaddle/PARL):
//   parl/algorithms/torch/a2c.py:40-60              A2C  (SUM reductions)
//   parl/algorithms/torch/ppo.py:102-138            PPO  (MEAN reductions, adv-norm, clipping)
//   parl/algorithms/torch/dqn.py:64-69, ddqn.py:64-72   TD target + MSE (MEAN)
//   benchmark/fluid/Prioritized_DQN/per_alg.py:48-69     PER-weighted TD loss, |delta|
//   parl/algorithms/torch/policy_gradient.py:54-75   REINFORCE on probabilities (MEAN)
// Layout: a CTA stages a tile of kEPB consecutive rows of the [N,A] logits into
// shared memory (16-byte cp.async, fully coalesced), one thread owns a row, the
// gradient tile is written back coalesced.  Bytes per row: read 4A + O(20),
// write 4A + 4.
#include "common.cuh"
#include "reduce.cuh"

namespace rl {

constexpr int kFT = 128;    // threads per CTA
constexpr int kEPB = 128;   // rows per CTA (one per thread)

enum { MODE_A2C = 0, MODE_PPO = 1 };

struct FlatArgs {
  const float* logits;      // [N,A]
  const float* values;      // [N]
  const void* actions;      // [N] i32 / i64
  const float* adv;         // [N]
  const float* target;      // A2C: target_values ; PPO: batch_return
  const float* old_value;   // PPO
  const float* old_logp;    // PPO
  const float* adv_stats;   // PPO: {mean, 1/(std+1e-8)} or NULL (no normalisation)
  float* d_logits;
  float* d_values;
  float* losses;
  float* partials;
  unsigned* ticket;
  long long N;
  int A, act64, vec, clip_value;
  float vf_coeff, ent_coeff, clip;
};

template <bool TO_SMEM>
__device__ __forceinline__ void flat_tile_copy(float* s, const float* g, float* gout, long long row0, int nrows, int A,
                                               bool vec) {
  const long long base = row0 * A;
  const int n = nrows * A;
  if (vec) {
    for (int i = threadIdx.x; i < (n >> 2); i += kFT) {
      if (TO_SMEM) cp_async16(s + 4 * i, g + base + 4 * i);
      else *reinterpret_cast<float4*>(gout + base + 4 * i) = *reinterpret_cast<const float4*>(s + 4 * i);
    }
  } else {
    for (int i = threadIdx.x; i < n; i += kFT) {
      if (TO_SMEM) cp_async4(s + i, g + base + i);
      else gout[base + i] = s[i];
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kFT) flat_categorical_loss_kernel(const FlatArgs p) {
  extern __shared__ float4 smem4[];
  float* s = reinterpret_cast<float*>(smem4);
  const int A = p.A;
  const long long row0 = (long long)blockIdx.x * kEPB;
  const int nrows = (int)min((long long)kEPB, p.N - row0);
  const bool vec = p.vec && ((nrows * A) & 3) == 0;
  flat_tile_copy<true>(s, p.logits, nullptr, row0, nrows, A, vec);
  cp_async_commit();
  const int tid = threadIdx.x;
  const long long g = row0 + tid;
  const bool valid = tid < nrows;
  int act = 0;
  float adv = 0.f, tgt = 0.f, v = 0.f, oldv = 0.f, oldlp = 0.f;
  if (valid) {
    act = p.act64 ? (int)reinterpret_cast<const long long*>(p.actions)[g] : reinterpret_cast<const int*>(p.actions)[g];
    adv = p.adv[g], tgt = p.target[g], v = p.values[g];
    if (MODE == MODE_PPO) oldv = p.old_value[g], oldlp = p.old_logp[g];
  }
  float mean = 0.f, inv_std = 1.f;
  if (MODE == MODE_PPO && p.adv_stats) mean = p.adv_stats[0], inv_std = p.adv_stats[1];
  cp_async_wait<0>();
  __syncthreads();
  float sums[3] = {0.f, 0.f, 0.f};   // pi, vf, entropy
  if (valid) {
    float* x = s + tid * A;         // row stride A words: conflict-free for odd A, 2-way for even A
    float m = x[0];
    for (int j = 1; j < A; ++j) m = fmaxf(m, x[j]);
    float S = 0.f;
    for (int j = 0; j < A; ++j) S += __expf(x[j] - m);
    const float logS = __logf(S), inv = __fdividef(1.0f, S);
    float H = 0.f, la = 0.f;
    for (int j = 0; j < A; ++j) {
      const float lj = x[j] - m - logS;
      H -= __expf(x[j] - m) * inv * lj;
      if (j == act) la = lj;
    }
    float gl, ce, dv;        // dL/dlogp_a, coefficient of p_j(logp_j + H), dL/dV
    if (MODE == MODE_A2C) {
      sums[0] = -la * adv;                                   // a2c.py:48
      const float d = v - tgt;
      sums[1] = 0.5f * d * d;                                // :52-53
      sums[2] = H;                                           // :58
      gl = -adv;
      ce = -p.ent_coeff;                                     // d(c_e * H)/dz_j = -c_e p_j (logp_j + H)
      dv = p.vf_coeff * d;
    } else {
      const float invM = 1.0f / (float)p.N;
      const float an = (adv - mean) * inv_std;               // ppo.py:115-117
      const float ratio = expf(la - oldlp);                  // :119
      const float lo = 1.0f - p.clip, hi = 1.0f + p.clip;
      const float surr1 = ratio * an;
      const float surr2 = fminf(fmaxf(ratio, lo), hi) * an;
      sums[0] = -fminf(surr1, surr2);                        // :120-123
      const bool in_range = ratio >= lo && ratio <= hi;
      // torch.min splits the gradient evenly at ties; clamp passes gradient on [lo, hi]
      float w = in_range ? 1.0f : (surr1 < surr2 ? 1.0f : (surr1 == surr2 ? 0.5f : 0.0f));
      gl = -an * ratio * w * invM;
      const float d1 = v - tgt;
      float gv;
      if (p.clip_value) {
        const float dvv = v - oldv;
        const float vclip = oldv + fminf(fmaxf(dvv, -p.clip), p.clip);
        const float d2 = vclip - tgt;
        const float l1 = d1 * d1, l2 = d2 * d2;
        sums[1] = 0.5f * fmaxf(l1, l2);                      // :126-135
        const float g2 = (dvv >= -p.clip && dvv <= p.clip) ? d2 : 0.f;
        gv = l1 > l2 ? d1 : (l1 < l2 ? g2 : 0.5f * (d1 + g2));
      } else {
        sums[1] = 0.5f * d1 * d1;                            // :137
        gv = d1;
      }
      sums[2] = H;
      ce = p.ent_coeff * invM;                               // loss -= c_e * mean(H)
      dv = p.vf_coeff * gv * invM;
    }
    p.d_values[g] = dv;
    for (int j = 0; j < A; ++j) {
      const float xm = x[j] - m;
      const float pj = __expf(xm) * inv;
      const float lj = xm - logS;
      float d = ce * pj * (lj + H) - gl * pj;
      if (j == act) d += gl;
      x[j] = d;
    }
  }
  __syncthreads();
  flat_tile_copy<false>(s, nullptr, p.d_logits, row0, nrows, A, vec);
  double tot[3];
  if (grid_reduce<3, kFT>(sums, p.partials, p.ticket, tot)) {
    if (MODE == MODE_A2C) {
      const float pi = (float)tot[0], vf = (float)tot[1], ent = (float)tot[2];
      p.losses[0] = pi + vf * p.vf_coeff + ent * p.ent_coeff;   // a2c.py:60
      p.losses[1] = pi, p.losses[2] = vf, p.losses[3] = ent;
    } else {
      const double M = (double)p.N;
      const float al = (float)(tot[0] / M), vl = (float)(tot[1] / M), el = (float)(tot[2] / M);
      p.losses[0] = vl, p.losses[1] = al, p.losses[2] = el;     // ppo.py:149 return order
      p.losses[3] = vl * p.vf_coeff + al - el * p.ent_coeff;    // :138
    }
  }
}

// ---------------------------------------------------------------------------
// PPO, diagonal Gaussian policy (state-independent log-std).  One thread per row.
// ---------------------------------------------------------------------------
struct GaussArgs {
  const float* mean;        // [N,D]
  const float* logstd;      // [D]
  const float* action;      // [N,D]
  const float* values;
  const float* adv;
  const float* ret;
  const float* old_value;
  const float* old_logp;
  const float* adv_stats;
  float* d_mean;            // [N,D]
  float* d_logstd;          // [D]
  float* d_values;
  float* losses;
  float* partials;          // [grid, 3 + D]
  unsigned* ticket;
  long long N;
  int D, clip_value;
  float vf_coeff, ent_coeff, clip;
};

constexpr int kMaxD = 16;

__global__ void __launch_bounds__(kFT) ppo_gaussian_loss_kernel(const GaussArgs p) {
  __shared__ float s_ls[kMaxD], s_iv[kMaxD];
  __shared__ float s_dls[kMaxD][kFT / 32];
  const int tid = threadIdx.x, D = p.D;
  if (tid < D) {
    const float ls = p.logstd[tid];
    const float sd = expf(ls);
    s_ls[tid] = logf(sd);                 // Normal(mean, exp(logstd)): log(scale) as torch computes it
    s_iv[tid] = 1.0f / (sd * sd);
  }
  __syncthreads();
  const long long g = (long long)blockIdx.x * kFT + tid;
  const bool valid = g < p.N;
  float sums[3] = {0.f, 0.f, 0.f};
  float dls[kMaxD];
#pragma unroll
  for (int d = 0; d < kMaxD; ++d) dls[d] = 0.f;
  if (valid) {
    float diff[kMaxD];
    float lp = 0.f, H = 0.f;
#pragma unroll
    for (int d = 0; d < kMaxD; ++d) {
      if (d < D) {
        diff[d] = p.action[g * D + d] - p.mean[g * D + d];
        lp += -0.5f * diff[d] * diff[d] * s_iv[d] - s_ls[d] - 0.9189385332046727f;   // ppo.py:107
        H += 1.4189385332046727f + s_ls[d];                                          // :108
      }
    }
    float mean = 0.f, inv_std = 1.f;
    if (p.adv_stats) mean = p.adv_stats[0], inv_std = p.adv_stats[1];
    const float invM = 1.0f / (float)p.N;
    const float an = (p.adv[g] - mean) * inv_std;
    const float ratio = expf(lp - p.old_logp[g]);
    const float lo = 1.0f - p.clip, hi = 1.0f + p.clip;
    const float surr1 = ratio * an, surr2 = fminf(fmaxf(ratio, lo), hi) * an;
    sums[0] = -fminf(surr1, surr2);
    const bool in_range = ratio >= lo && ratio <= hi;
    const float w = in_range ? 1.0f : (surr1 < surr2 ? 1.0f : (surr1 == surr2 ? 0.5f : 0.0f));
    const float gl = -an * ratio * w * invM;
    const float v = p.values[g], tgt = p.ret[g];
    const float d1 = v - tgt;
    float gv;
    if (p.clip_value) {
      const float oldv = p.old_value[g];
      const float dvv = v - oldv;
      const float d2 = oldv + fminf(fmaxf(dvv, -p.clip), p.clip) - tgt;
      const float l1 = d1 * d1, l2 = d2 * d2;
      sums[1] = 0.5f * fmaxf(l1, l2);
      const float g2 = (dvv >= -p.clip && dvv <= p.clip) ? d2 : 0.f;
      gv = l1 > l2 ? d1 : (l1 < l2 ? g2 : 0.5f * (d1 + g2));
    } else {
      sums[1] = 0.5f * d1 * d1;
      gv = d1;
    }
    sums[2] = H;
    p.d_values[g] = p.vf_coeff * gv * invM;
#pragma unroll
    for (int d = 0; d < kMaxD; ++d) {
      if (d < D) {
        p.d_mean[g * D + d] = gl * diff[d] * s_iv[d];
        dls[d] = gl * (diff[d] * diff[d] * s_iv[d] - 1.0f);
      }
    }
  }
  // d_logstd partials ride in the same deterministic grid reduction
#pragma unroll
  for (int d = 0; d < kMaxD; ++d) {
    if (d < D) {
      const float w = warp_sum(dls[d]);
      if ((tid & 31) == 0) s_dls[d][tid >> 5] = w;
    }
  }
  __syncthreads();
  float* part = p.partials + (long long)gridDim.x * 3;        // [grid, D] after the [grid,3] loss partials
  if (tid < D) {
    float a = 0.f;
    for (int w = 0; w < kFT / 32; ++w) a += s_dls[tid][w];
    part[(long long)blockIdx.x * D + tid] = a;
  }
  double tot[3];
  const bool fin = grid_reduce<3, kFT>(sums, p.partials, p.ticket, tot);
  // grid_reduce returns true only in thread 0 of the last CTA; all its partial writes are visible (threadfence)
  if (fin) {
    const double M = (double)p.N;
    const float al = (float)(tot[0] / M), vl = (float)(tot[1] / M), el = (float)(tot[2] / M);
    p.losses[0] = vl, p.losses[1] = al, p.losses[2] = el;
    p.losses[3] = vl * p.vf_coeff + al - el * p.ent_coeff;
  }
  // d_logstd: the WHOLE last CTA adds up the [grid, D] partials (fixed order: strided per thread, shuffle tree, warps
  // in order, fp64).  Round 2: a single thread walking grid*D dependent L2 loads cost 1 ms per 131 072-row minibatch.
  __shared__ bool s_fin;
  __shared__ double s_dd[kFT / 32];
  if (tid == 0) s_fin = fin;
  __syncthreads();
  if (s_fin) {
    for (int d = 0; d < D; ++d) {
      double a = 0.0;
      for (unsigned i = tid; i < gridDim.x; i += kFT) a += (double)__ldcg(part + (long long)i * D + d);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if ((tid & 31) == 0) s_dd[tid >> 5] = a;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < kFT / 32; ++w) t += s_dd[w];
        p.d_logstd[d] = (float)t - p.ent_coeff;               // d(-c_e * mean_m sum_d (.. + logstd_d)) = -c_e
      }
      __syncthreads();
    }
  }
}

// mean and 1/(unbiased std + 1e-8) of adv[N] (ppo.py:115-117) -> stats[2]; one CTA, fp64.
__global__ void __launch_bounds__(1024) adv_stats_kernel(const float* __restrict__ adv, long long N,
                                                         float* __restrict__ stats) {
  __shared__ double s1[32], s2[32];
  __shared__ double s_mean;
  double a = 0.0;
  for (long long i = threadIdx.x; i < N; i += blockDim.x) a += (double)adv[i];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) s1[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)blockDim.x / 32; ++w) t += s1[w];
    s_mean = t / (double)N;
  }
  __syncthreads();
  const double mu = s_mean;
  double q = 0.0;
  for (long long i = threadIdx.x; i < N; i += blockDim.x) {
    const double d = (double)adv[i] - mu;
    q += d * d;
  }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if ((threadIdx.x & 31) == 0) s2[threadIdx.x >> 5] = q;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)blockDim.x / 32; ++w) t += s2[w];
    const float sd = (float)sqrt(t / (double)(N - 1));
    stats[0] = (float)mu;
    stats[1] = 1.0f / (sd + 1e-8f);
  }
}

// ---------------------------------------------------------------------------
// TD loss (DQN / DDQN / PER-weighted) and REINFORCE on probabilities.
// ---------------------------------------------------------------------------
struct TdArgs {
  const float* q;           // [M,A] online Q(s)
  const float* q_tgt_next;  // [M,A] target Q(s')
  const float* q_onl_next;  // [M,A] online Q(s') (DDQN) or NULL
  const void* action;
  const float* reward;
  const float* terminal;    // f32 0/1
  const float* weights;     // PER IS weights or NULL
  float* d_q;               // [M,A]
  float* td_abs;            // [M] or NULL
  float* losses;
  float* partials;
  unsigned* ticket;
  long long M;
  int A, act64;
  float gamma;
};

__global__ void __launch_bounds__(kFT) td_loss_kernel(const TdArgs p) {
  const long long g = (long long)blockIdx.x * kFT + threadIdx.x;
  float sums[1] = {0.f};
  if (g < p.M) {
    const int A = p.A;
    const float* qt = p.q_tgt_next + g * A;
    float best;
    if (p.q_onl_next) {                       // ddqn.py:66-70: argmax from the online net (first max on ties)
      const float* qo = p.q_onl_next + g * A;
      int arg = 0;
      float mo = qo[0];
      for (int j = 1; j < A; ++j)
        if (qo[j] > mo) mo = qo[j], arg = j;
      best = qt[arg];
    } else {                                  // dqn.py:66
      best = qt[0];
      for (int j = 1; j < A; ++j) best = fmaxf(best, qt[j]);
    }
    const int a = p.act64 ? (int)reinterpret_cast<const long long*>(p.action)[g]
                          : reinterpret_cast<const int*>(p.action)[g];
    // target = reward + (1 - terminal) * gamma * max_v          dqn.py:67
    const float target = __fadd_rn(p.reward[g], __fmul_rn(__fmul_rn(__fsub_rn(1.0f, p.terminal[g]), p.gamma), best));
    const float pred = p.q[g * A + a];
    const float diff = pred - target;
    const float w = p.weights ? p.weights[g] : 1.0f;
    sums[0] = w * diff * diff;
    if (p.td_abs) p.td_abs[g] = fabsf(target - pred);            // per_alg.py:63
    const float gq = 2.0f * w * diff / (float)p.M;
    for (int j = 0; j < A; ++j) p.d_q[g * A + j] = (j == a) ? gq : 0.f;
  }
  double tot[1];
  if (grid_reduce<1, kFT>(sums, p.partials, p.ticket, tot)) p.losses[0] = (float)(tot[0] / (double)p.M);
}

__global__ void __launch_bounds__(kFT) pg_prob_loss_kernel(const float* __restrict__ prob, const void* action, int act64,
                                                           const float* __restrict__ reward, long long N, int A,
                                                           float* __restrict__ d_prob, float* losses, float* partials,
                                                           unsigned* ticket) {
  const long long g = (long long)blockIdx.x * kFT + threadIdx.x;
  float sums[1] = {0.f};
  if (g < N) {
    // Categorical(probs) normalises: p_j / sum_k p_k   (policy_gradient.py:71-73)
    float S = 0.f;
    for (int j = 0; j < A; ++j) S += prob[g * A + j];
    const int a = act64 ? (int)reinterpret_cast<const long long*>(action)[g] : reinterpret_cast<const int*>(action)[g];
    const float pa = prob[g * A + a];
    const float r = reward[g];
    sums[0] = -(logf(pa) - logf(S)) * r;
    const float c = r / (float)N;
    for (int j = 0; j < A; ++j) d_prob[g * A + j] = c / S - ((j == a) ? c / pa : 0.f);
  }
  double tot[1];
  if (grid_reduce<1, kFT>(sums, partials, ticket, tot)) losses[0] = (float)(tot[0] / (double)N);
}

// Continuous-control critic TD (DDPG / TD3 / SAC): target = r + gamma (1 - terminal) (min(Q1', Q2') - alpha log pi'),
// loss = mse(Q1, target) + mse(Q2, target), gradients dQ = 2 (Q - target) / N.  One launch, thread per sample.
struct TwinQArgs {
  const float* q1;
  const float* q2;          // NULL: single critic (DDPG)
  const float* tq1;
  const float* tq2;         // NULL: single target critic
  const float* next_logp;   // NULL: no entropy term (DDPG / TD3)
  const float* reward;
  const float* terminal;
  float* d_q1;
  float* d_q2;
  float* target_out;        // NULL or [N]
  float* losses;
  float* partials;
  unsigned* ticket;
  long long N;
  float gamma, alpha;
};

__global__ void __launch_bounds__(kFT) twin_q_td_loss_kernel(const TwinQArgs p) {
  const long long g = (long long)blockIdx.x * kFT + threadIdx.x;
  float sums[2] = {0.f, 0.f};
  if (g < p.N) {
    float tq = p.tq1[g];
    if (p.tq2) tq = fminf(tq, p.tq2[g]);                                       // td3.py:88, sac.py:94
    if (p.next_logp) tq = __fsub_rn(tq, __fmul_rn(p.alpha, p.next_logp[g]));   // sac.py:94
    // reward + (1 - terminal) * gamma * target_Q      td3.py:89, ddpg.py:67 (sac.py:95: gamma * (1 - terminal), same product)
    const float target = __fadd_rn(p.reward[g], __fmul_rn(__fmul_rn(__fsub_rn(1.0f, p.terminal[g]), p.gamma), tq));
    if (p.target_out) p.target_out[g] = target;
    const float inv = 2.0f / (float)p.N;
    const float d1 = p.q1[g] - target;
    sums[0] = d1 * d1;
    p.d_q1[g] = d1 * inv;
    if (p.q2) {
      const float d2 = p.q2[g] - target;
      sums[1] = d2 * d2;
      p.d_q2[g] = d2 * inv;
    }
  }
  double tot[2];
  if (grid_reduce<2, kFT>(sums, p.partials, p.ticket, tot)) {
    const float m1 = (float)(tot[0] / (double)p.N), m2 = (float)(tot[1] / (double)p.N);
    p.losses[0] = m1 + m2;                                                     // td3.py:93-94, sac.py:98-99
    p.losses[1] = m1;
    p.losses[2] = m2;
  }
}

static unsigned* ws_ticket(void* ws) { return reinterpret_cast<unsigned*>(ws); }
static float* ws_partials(void* ws) { return reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 256); }

}  // namespace rl

using namespace rl;

extern "C" size_t rl_flat_workspace_bytes(long long n_rows, int extra) {
  const size_t grid = (size_t)((n_rows + kFT - 1) / kFT) + 1;
  return 256 + grid * (size_t)(4 + (extra > 0 ? extra : 0)) * sizeof(float);
}

static int launch_flat(int mode, const FlatArgs& a, cudaStream_t st) {
  const unsigned grid = (unsigned)((a.N + kEPB - 1) / kEPB);
  const size_t smem = (size_t)kEPB * a.A * sizeof(float);
  if (mode == MODE_A2C) {
    RL_SMEM_OPTIN(flat_categorical_loss_kernel<MODE_A2C>);
    flat_categorical_loss_kernel<MODE_A2C><<<grid, kFT, smem, st>>>(a);
  } else {
    RL_SMEM_OPTIN(flat_categorical_loss_kernel<MODE_PPO>);
    flat_categorical_loss_kernel<MODE_PPO><<<grid, kFT, smem, st>>>(a);
  }
  return 0;
}

extern "C" int rl_a2c_loss_fwd_bwd(const float* logits, const float* values, const void* actions, int actions_i64,
                                   const float* advantages, const float* target_values, long long N, int A,
                                   float vf_loss_coeff, float entropy_coeff, float* losses, float* d_logits,
                                   float* d_values, void* workspace, size_t workspace_bytes, rl_stream_t stream) {
  RL_CHECK_ARG(logits && values && actions && advantages && target_values && losses && d_logits && d_values && workspace,
               "a2c_loss: null pointer");
  RL_CHECK_ARG(N > 0 && A >= 1 && A <= 400, "a2c_loss: bad shape N=%lld A=%d", N, A);
  if (workspace_bytes < rl_flat_workspace_bytes(N, 0)) {
    set_error("a2c_loss: workspace too small");
    return RL_ERR_WORKSPACE;
  }
  FlatArgs a = {};
  a.logits = logits, a.values = values, a.actions = actions, a.adv = advantages, a.target = target_values;
  a.d_logits = d_logits, a.d_values = d_values, a.losses = losses;
  a.ticket = ws_ticket(workspace), a.partials = ws_partials(workspace);
  a.N = N, a.A = A, a.act64 = actions_i64, a.vf_coeff = vf_loss_coeff, a.ent_coeff = entropy_coeff;
  a.vec = aligned16(logits) && aligned16(d_logits) && ((kEPB * A) % 4 == 0);
  launch_flat(MODE_A2C, a, (cudaStream_t)stream);
  RL_CHECK_LAUNCH("rl_a2c_loss_fwd_bwd");
  return RL_OK;
}

extern "C" int rl_adv_stats(const float* adv, long long N, float* stats, rl_stream_t stream) {
  RL_CHECK_ARG(adv && stats && N >= 2, "adv_stats: bad argument");
  adv_stats_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(adv, N, stats);
  RL_CHECK_LAUNCH("rl_adv_stats");
  return RL_OK;
}

extern "C" int rl_ppo_loss_fwd_bwd(const float* logits, const float* mean, const float* logstd, const void* actions,
                                   int actions_i64, const float* values, const float* batch_value,
                                   const float* batch_return, const float* batch_logprob, const float* batch_adv,
                                   const float* adv_stats, long long N, int A_or_D, float clip_param,
                                   float value_loss_coef, float entropy_coef, int use_clipped_value_loss, float* losses,
                                   float* d_logits_or_mean, float* d_logstd, float* d_values, void* workspace,
                                   size_t workspace_bytes, rl_stream_t stream) {
  RL_CHECK_ARG((logits != nullptr) != (mean != nullptr), "ppo_loss: pass exactly one of logits / mean");
  RL_CHECK_ARG(actions && values && batch_value && batch_return && batch_logprob && batch_adv && losses &&
                   d_logits_or_mean && d_values && workspace,
               "ppo_loss: null pointer");
  RL_CHECK_ARG(N > 0 && A_or_D >= 1, "ppo_loss: bad shape");
  if (workspace_bytes < rl_flat_workspace_bytes(N, A_or_D)) {
    set_error("ppo_loss: workspace too small");
    return RL_ERR_WORKSPACE;
  }
  if (logits) {
    RL_CHECK_ARG(A_or_D <= 400, "ppo_loss: A too large");
    FlatArgs a = {};
    a.logits = logits, a.values = values, a.actions = actions, a.adv = batch_adv, a.target = batch_return;
    a.old_value = batch_value, a.old_logp = batch_logprob, a.adv_stats = adv_stats;
    a.d_logits = d_logits_or_mean, a.d_values = d_values, a.losses = losses;
    a.ticket = ws_ticket(workspace), a.partials = ws_partials(workspace);
    a.N = N, a.A = A_or_D, a.act64 = actions_i64, a.vf_coeff = value_loss_coef, a.ent_coeff = entropy_coef;
    a.clip = clip_param, a.clip_value = use_clipped_value_loss;
    a.vec = aligned16(logits) && aligned16(d_logits_or_mean) && ((kEPB * A_or_D) % 4 == 0);
    launch_flat(MODE_PPO, a, (cudaStream_t)stream);
  } else {
    RL_CHECK_ARG(logstd && d_logstd && A_or_D <= kMaxD, "ppo_loss: Gaussian needs logstd/d_logstd and D <= %d", kMaxD);
    GaussArgs a = {};
    a.mean = mean, a.logstd = logstd, a.action = reinterpret_cast<const float*>(actions), a.values = values;
    a.adv = batch_adv, a.ret = batch_return, a.old_value = batch_value, a.old_logp = batch_logprob;
    a.adv_stats = adv_stats, a.d_mean = d_logits_or_mean, a.d_logstd = d_logstd, a.d_values = d_values;
    a.losses = losses, a.ticket = ws_ticket(workspace), a.partials = ws_partials(workspace);
    a.N = N, a.D = A_or_D, a.clip_value = use_clipped_value_loss;
    a.vf_coeff = value_loss_coef, a.ent_coeff = entropy_coef, a.clip = clip_param;
    const unsigned grid = (unsigned)((N + kFT - 1) / kFT);
    ppo_gaussian_loss_kernel<<<grid, kFT, 0, (cudaStream_t)stream>>>(a);
  }
  RL_CHECK_LAUNCH("rl_ppo_loss_fwd_bwd");
  return RL_OK;
}

extern "C" int rl_td_loss_fwd_bwd(const float* q, const float* q_target_next, const float* q_online_next,
                                  const void* action, int action_i64, const float* reward, const float* terminal,
                                  const float* weights, long long M, int A, float gamma, float* losses, float* d_q,
                                  float* td_abs, void* workspace, size_t workspace_bytes, rl_stream_t stream) {
  RL_CHECK_ARG(q && q_target_next && action && reward && terminal && losses && d_q && workspace, "td_loss: null pointer");
  RL_CHECK_ARG(M > 0 && A >= 1, "td_loss: bad shape");
  if (workspace_bytes < rl_flat_workspace_bytes(M, 0)) {
    set_error("td_loss: workspace too small");
    return RL_ERR_WORKSPACE;
  }
  TdArgs a = {};
  a.q = q, a.q_tgt_next = q_target_next, a.q_onl_next = q_online_next, a.action = action, a.reward = reward;
  a.terminal = terminal, a.weights = weights, a.d_q = d_q, a.td_abs = td_abs, a.losses = losses;
  a.ticket = ws_ticket(workspace), a.partials = ws_partials(workspace);
  a.M = M, a.A = A, a.act64 = action_i64, a.gamma = gamma;
  td_loss_kernel<<<(unsigned)((M + kFT - 1) / kFT), kFT, 0, (cudaStream_t)stream>>>(a);
  RL_CHECK_LAUNCH("rl_td_loss_fwd_bwd");
  return RL_OK;
}

extern "C" int rl_twin_q_td_loss_fwd_bwd(const float* q1, const float* q2, const float* q1_target_next,
                                         const float* q2_target_next, const float* next_log_prob, const float* reward,
                                         const float* terminal, long long N, float gamma, float alpha, float* losses,
                                         float* d_q1, float* d_q2, float* target_out, void* workspace,
                                         size_t workspace_bytes, rl_stream_t stream) {
  RL_CHECK_ARG(q1 && q1_target_next && reward && terminal && losses && d_q1 && workspace && N > 0,
               "twin_q_td_loss: bad argument");
  RL_CHECK_ARG((q2 == nullptr) == (d_q2 == nullptr), "twin_q_td_loss: q2 and d_q2 go together");
  if (workspace_bytes < rl_flat_workspace_bytes(N, 0)) {
    set_error("twin_q_td_loss: workspace too small");
    return RL_ERR_WORKSPACE;
  }
  TwinQArgs a = {};
  a.q1 = q1, a.q2 = q2, a.tq1 = q1_target_next, a.tq2 = q2_target_next, a.next_logp = next_log_prob;
  a.reward = reward, a.terminal = terminal, a.d_q1 = d_q1, a.d_q2 = d_q2, a.target_out = target_out, a.losses = losses;
  a.ticket = ws_ticket(workspace), a.partials = ws_partials(workspace);
  a.N = N, a.gamma = gamma, a.alpha = alpha;
  twin_q_td_loss_kernel<<<(unsigned)((N + kFT - 1) / kFT), kFT, 0, (cudaStream_t)stream>>>(a);
  RL_CHECK_LAUNCH("rl_twin_q_td_loss_fwd_bwd");
  return RL_OK;
}

extern "C" int rl_pg_loss_fwd_bwd(const float* prob, const void* action, int action_i64, const float* reward,
                                  long long N, int A, float* losses, float* d_prob, void* workspace,
                                  size_t workspace_bytes, rl_stream_t stream) {
  RL_CHECK_ARG(prob && action && reward && losses && d_prob && workspace && N > 0 && A >= 1, "pg_loss: bad argument");
  if (workspace_bytes < rl_flat_workspace_bytes(N, 0)) {
    set_error("pg_loss: workspace too small");
    return RL_ERR_WORKSPACE;
  }
  pg_prob_loss_kernel<<<(unsigned)((N + kFT - 1) / kFT), kFT, 0, (cudaStream_t)stream>>>(
      prob, action, action_i64, reward, N, A, d_prob, losses, ws_partials(workspace), ws_ticket(workspace));
  RL_CHECK_LAUNCH("rl_pg_loss_fwd_bwd");
  return RL_OK;
}
