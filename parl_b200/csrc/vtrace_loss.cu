// K1 — fused IMPALA loss: log-softmax + entropy + KL + V-trace backward scan +
// policy-gradient/value/entropy loss and its gradient, ONE launch.
//
// Reference arithmetic being replaced (PaddlePaddle/PARL):
//   parl/algorithms/paddle/impala/vtrace.py:99-139   V-trace recurrences
//   parl/algorithms/paddle/impala/impala.py:25-79    VTraceLoss (SUM reductions)
//   parl/algorithms/paddle/impala/impala.py:119-132  log-prob = sum(log_softmax * onehot)
//   parl/algorithms/paddle/impala/impala.py:148-208  row dropping / bootstrap / KL
//
// Layout in HBM: logits [T,B,A] f32, per-step scalars [T,B]; a CTA owns BW=4
// adjacent env columns for ALL T rows (the scan never leaves the CTA).  The
// logits tile is staged into shared memory with 16-byte cp.async (coalesced
// 288-byte row segments at A=18), every (t,b) element is then private to one
// thread for the softmax/entropy work, the T-long recurrence runs on one lane
// per column in the reference's exact operation order, and the gradient tile
// is written back from shared memory with 16-byte coalesced stores.
// Algorithmic traffic: (12A+17) bytes per kept (t,b) element.
#include <stdarg.h>

#include "common.cuh"

namespace rl {

constexpr int kBW = 4;     // env columns per CTA
constexpr int kNT = 128;   // threads per CTA
constexpr int kEPT = 2;    // max (t,b) elements per thread per chunk  (TC*BW <= kNT*kEPT)

struct VtraceLossArgs {
  const float* tl;
  const float* bl;
  const void* actions;
  const float* rewards;
  const uint8_t* dones;
  const float* values;
  float* d_logits;
  float* d_values;
  float* vs_out;
  float* pg_out;
  float* losses;
  float* partials;     // [grid, 4]
  unsigned* ticket;    // zero on entry, zero on exit
  int T, B, A, TC;
  int act64, vec;
  float gamma, clip_rho, clip_pg, vf_coeff, ent_coeff;
};

template <bool EM>
__device__ __forceinline__ long long gidx(int t, int b, int T, int B) {
  return EM ? (long long)b * T + t : (long long)t * B + b;
}
template <bool EM>
__device__ __forceinline__ int sidx(int tl, int bl, int TC) {
  return EM ? bl * TC + tl : tl * kBW + bl;
}

// global <-> shared tile copy.  TM: nt row segments of nb*A floats; EM: nb column
// segments of nt*A floats.  `to_smem` selects direction.
template <bool EM, bool TO_SMEM>
__device__ __forceinline__ void copy_tile(float* s, const float* gsrc, float* gdst, int A, int T, int B, int TC,
                                          int t0, int nt, int b0, int nb, bool vec) {
  const int nseg = EM ? nb : nt;
  const int seglen = (EM ? nt : nb) * A;           // floats
  const int sstride = (EM ? TC : kBW) * A;         // floats between segments in smem
  if (vec) {
    const int segv = seglen >> 2;
    const int total = nseg * segv;
    for (int i = threadIdx.x; i < total; i += kNT) {
      const int sg = i / segv, k = i - sg * segv;
      const long long goff = (EM ? ((long long)(b0 + sg) * T + t0) : ((long long)(t0 + sg) * B + b0)) * A + 4 * k;
      float* sp = s + sg * sstride + 4 * k;
      if (TO_SMEM) {
        cp_async16(sp, gsrc + goff);
      } else {
        *reinterpret_cast<float4*>(gdst + goff) = *reinterpret_cast<const float4*>(sp);
      }
    }
  } else {
    const int total = nseg * seglen;
    for (int i = threadIdx.x; i < total; i += kNT) {
      const int sg = i / seglen, k = i - sg * seglen;
      const long long goff = (EM ? ((long long)(b0 + sg) * T + t0) : ((long long)(t0 + sg) * B + b0)) * A + k;
      float* sp = s + sg * sstride + k;
      if (TO_SMEM) {
        cp_async4(sp, gsrc + goff);
      } else {
        gdst[goff] = *sp;
      }
    }
  }
}

// Per-element softmax statistics.  Reads the element's A target / behaviour
// logits from shared memory, overwrites them with p_j and q_j = p_j * logp_j.
template <int A_>
__device__ __forceinline__ void softmax_stats(float* st, float* sb, int A, int act, float& la, float& lma, float& H,
                                              float& KL) {
  if constexpr (A_ > 0) {
    float x[A_], y[A_];
    if constexpr ((A_ & 1) == 0) {
#pragma unroll
      for (int j = 0; j < A_; j += 2) {
        const float2 a = *reinterpret_cast<const float2*>(st + j);
        const float2 b = *reinterpret_cast<const float2*>(sb + j);
        x[j] = a.x, x[j + 1] = a.y, y[j] = b.x, y[j + 1] = b.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < A_; ++j) x[j] = st[j], y[j] = sb[j];
    }
    float m = x[0], my = y[0];
#pragma unroll
    for (int j = 1; j < A_; ++j) m = fmaxf(m, x[j]), my = fmaxf(my, y[j]);
    float S = 0.f, Sy = 0.f;
#pragma unroll
    for (int j = 0; j < A_; ++j) {
      x[j] -= m;
      y[j] -= my;
      S += __expf(x[j]);
      Sy += __expf(y[j]);
    }
    const float logS = __logf(S), logSy = __logf(Sy);
    const float inv = __fdividef(1.0f, S);
    H = 0.f, KL = 0.f, la = 0.f, lma = 0.f;
#pragma unroll
    for (int j = 0; j < A_; ++j) {
      const float lj = x[j] - logS, lmj = y[j] - logSy;
      const float pj = __expf(x[j]) * inv;
      const float qj = pj * lj;
      H -= qj;
      KL = fmaf(pj, lj - lmj, KL);
      if (j == act) la = lj, lma = lmj;
      x[j] = pj, y[j] = qj;
    }
    if constexpr ((A_ & 1) == 0) {
#pragma unroll
      for (int j = 0; j < A_; j += 2) {
        *reinterpret_cast<float2*>(st + j) = make_float2(x[j], x[j + 1]);
        *reinterpret_cast<float2*>(sb + j) = make_float2(y[j], y[j + 1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < A_; ++j) st[j] = x[j], sb[j] = y[j];
    }
  } else {
    float m = st[0], my = sb[0];
    for (int j = 1; j < A; ++j) m = fmaxf(m, st[j]), my = fmaxf(my, sb[j]);
    float S = 0.f, Sy = 0.f;
    for (int j = 0; j < A; ++j) S += __expf(st[j] - m), Sy += __expf(sb[j] - my);
    const float logS = __logf(S), logSy = __logf(Sy);
    const float inv = __fdividef(1.0f, S);
    H = 0.f, KL = 0.f, la = 0.f, lma = 0.f;
    for (int j = 0; j < A; ++j) {
      const float xj = st[j] - m, yj = sb[j] - my;
      const float lj = xj - logS, lmj = yj - logSy;
      const float pj = __expf(xj) * inv;
      const float qj = pj * lj;
      H -= qj;
      KL = fmaf(pj, lj - lmj, KL);
      if (j == act) la = lj, lma = lmj;
      st[j] = pj, sb[j] = qj;
    }
  }
}

template <int A_, bool EM>
__global__ void __launch_bounds__(kNT, 7) vtrace_loss_kernel(const VtraceLossArgs p) {
  extern __shared__ float4 smem4[];
  const int A = A_ > 0 ? A_ : p.A;
  const int T = p.T, B = p.B, TC = p.TC;
  float* s_tl = reinterpret_cast<float*>(smem4);
  float* s_bl = s_tl + TC * kBW * A;
  float* s_acc = s_bl + TC * kBW * A;   // delta, then acc = vs - V
  float* s_kc = s_acc + TC * kBW;       // gamma_t * min(rho, 1)
  __shared__ float s_carry[kBW];        // acc at the first row of the chunk processed before (later in time)
  __shared__ float s_red[4][kNT / 32];
  __shared__ bool s_last;

  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * kBW;
  const int nb = min(kBW, B - b0);
  float sum_pi = 0.f, sum_vf = 0.f, sum_ent = 0.f, sum_kl = 0.f;
  float acc_carry = 0.f;
  if (tid < kBW) s_carry[tid] = 0.f;

  const int nchunks = (T + TC - 1) / TC;
  for (int c = nchunks - 1; c >= 0; --c) {
    const int t0 = c * TC;
    const int nt = min(TC, T - t0);
    const bool vec = p.vec && (((EM ? nt : nb) * A) & 3) == 0;
    copy_tile<EM, true>(s_tl, p.tl, nullptr, A, T, B, TC, t0, nt, b0, nb, vec);
    copy_tile<EM, true>(s_bl, p.bl, nullptr, A, T, B, TC, t0, nt, b0, nb, vec);
    cp_async_commit();

    // ---- per-element scalars straight from global (overlaps the tile copy) ----
    const int nel = nt * nb;
    int e_act[kEPT];
    float e_r[kEPT], e_v[kEPT], e_vn[kEPT], e_g[kEPT];
#pragma unroll
    for (int k = 0; k < kEPT; ++k) {
      const int i = tid + k * kNT;
      e_act[k] = 0, e_r[k] = 0.f, e_v[k] = 0.f, e_vn[k] = 0.f, e_g[k] = 0.f;
      if (i < nel) {
        const int tl_ = EM ? i % nt : i / nb, bl_ = EM ? i / nt : i % nb;
        const int t = t0 + tl_;
        const long long g = gidx<EM>(t, b0 + bl_, T, B);
        e_act[k] = p.act64 ? (int)reinterpret_cast<const long long*>(p.actions)[g]
                           : reinterpret_cast<const int*>(p.actions)[g];
        e_r[k] = p.rewards[g];
        e_v[k] = p.values[g];
        e_g[k] = p.dones[g] ? 0.0f : p.gamma;            // impala.py:59  (~dones) * discount
        if (t + 1 < T) e_vn[k] = p.values[gidx<EM>(t + 1, b0 + bl_, T, B)];
      }
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- phase A: softmax / entropy / KL / rho / delta (element-private) ----
    float e_la[kEPT], e_H[kEPT], e_rpg[kEPT];
#pragma unroll
    for (int k = 0; k < kEPT; ++k) {
      const int i = tid + k * kNT;
      e_la[k] = 0.f, e_H[k] = 0.f, e_rpg[k] = 0.f;
      if (i < nel) {
        const int tl_ = EM ? i % nt : i / nb, bl_ = EM ? i / nt : i % nb;
        const int si = sidx<EM>(tl_, bl_, TC);
        float la, lma, H, KL;
        softmax_stats<A_>(s_tl + si * A, s_bl + si * A, A, e_act[k], la, lma, H, KL);
        sum_kl += KL;                                       // impala.py:160-162: every row
        if (t0 + tl_ < T - 1) {
          const float rho = expf(la - lma);                 // vtrace.py:101-103
          const float rhoc = p.clip_rho >= 0.f ? fminf(rho, p.clip_rho) : rho;
          const float cs = fminf(rho, 1.0f);                // :109
          e_rpg[k] = p.clip_pg >= 0.f ? fminf(rho, p.clip_pg) : rho;
          // deltas = clipped_rhos * (rewards + discounts * values_t_plus_1 - values)   :115
          const float td = __fsub_rn(__fadd_rn(e_r[k], __fmul_rn(e_g[k], e_vn[k])), e_v[k]);
          s_acc[si] = __fmul_rn(rhoc, td);
          s_kc[si] = __fmul_rn(e_g[k], cs);
          e_la[k] = la, e_H[k] = H;
          sum_ent += H;
        }
      }
    }
    __syncthreads();

    // ---- phase B: backward scan, one lane per column, reference op order ----
    if (tid < nb) {
      float acc = acc_carry;
      const int tl_hi = min(nt, T - 1 - t0) - 1;            // skip the bootstrap row
      for (int tl_ = tl_hi; tl_ >= 0; --tl_) {
        const int si = sidx<EM>(tl_, tid, TC);
        acc = __fadd_rn(s_acc[si], __fmul_rn(s_kc[si], acc));   // vtrace.py:120
        s_acc[si] = acc;
      }
      acc_carry = acc;
    }
    __syncthreads();

    // ---- phase C: advantages, losses, gradient tile (in place over p_j) ----
#pragma unroll
    for (int k = 0; k < kEPT; ++k) {
      const int i = tid + k * kNT;
      if (i < nel) {
        const int tl_ = EM ? i % nt : i / nb, bl_ = EM ? i / nt : i % nb;
        const int t = t0 + tl_;
        const int si = sidx<EM>(tl_, bl_, TC);
        float* pt = s_tl + si * A;
        const long long g = gidx<EM>(t, b0 + bl_, T, B);
        if (t < T - 1) {
          const float acc_n = (t + 1 == T - 1) ? 0.f : (tl_ + 1 < nt ? s_acc[sidx<EM>(tl_ + 1, bl_, TC)] : s_carry[bl_]);
          const float vs = __fadd_rn(s_acc[si], e_v[k]);                 // vtrace.py:125
          const float vs_n = __fadd_rn(acc_n, e_vn[k]);                  // :128-129 (bootstrap at the end)
          const float adv =
              __fmul_rn(e_rpg[k], __fsub_rn(__fadd_rn(e_r[k], __fmul_rn(e_g[k], vs_n)), e_v[k]));   // :136-137
          const float dv = e_v[k] - vs;
          sum_pi -= e_la[k] * adv;                                        // impala.py:67-68
          sum_vf += 0.5f * dv * dv;                                       // :71-72
          p.d_values[g] = p.vf_coeff * dv;
          if (p.vs_out) p.vs_out[(long long)t * B + b0 + bl_] = vs;
          if (p.pg_out) p.pg_out[(long long)t * B + b0 + bl_] = adv;
          // dL/dz_j = p_j (adv - c_e H) - c_e p_j logp_j - adv [j == a]
          const float c0 = adv - p.ent_coeff * e_H[k];
          const float* pq = s_bl + si * A;
          if constexpr (A_ > 0) {
#pragma unroll
            for (int j = 0; j < A_; ++j) {
              float d = fmaf(pt[j], c0, -p.ent_coeff * pq[j]);
              if (j == e_act[k]) d -= adv;
              pt[j] = d;
            }
          } else {
            for (int j = 0; j < A; ++j) {
              float d = fmaf(pt[j], c0, -p.ent_coeff * pq[j]);
              if (j == e_act[k]) d -= adv;
              pt[j] = d;
            }
          }
        } else {
          p.d_values[g] = 0.f;                                            // bootstrap row: no gradient
          for (int j = 0; j < A; ++j) pt[j] = 0.f;
        }
      }
    }
    __syncthreads();
    if (tid < nb) s_carry[tid] = acc_carry;
    copy_tile<EM, false>(s_tl, nullptr, p.d_logits, A, T, B, TC, t0, nt, b0, nb, vec);
    __syncthreads();
  }

  // ---- loss reduction: warp -> CTA -> (last CTA) grid, fixed order, fp64 at the end ----
  float sums[4] = {sum_pi, sum_vf, sum_ent, sum_kl};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float w = warp_sum(sums[q]);
    if ((tid & 31) == 0) s_red[q][tid >> 5] = w;
  }
  __syncthreads();
  if (tid < 4) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kNT / 32; ++w) a += s_red[tid][w];
    p.partials[blockIdx.x * 4 + tid] = a;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = tid; i < (int)gridDim.x; i += kNT) {
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] += (double)__ldcg(p.partials + i * 4 + q);
    }
    __shared__ double s_dred[4][kNT / 32];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double v = acc[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0) s_dred[q][tid >> 5] = v;
    }
    __syncthreads();
    if (tid == 0) {
      double r[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        r[q] = 0.0;
        for (int w = 0; w < kNT / 32; ++w) r[q] += s_dred[q][w];
      }
      const float pi = (float)r[0], vf = (float)r[1], ent = (float)r[2];
      p.losses[0] = pi + vf * p.vf_coeff + ent * p.ent_coeff;      // impala.py:78-79
      p.losses[1] = pi;
      p.losses[2] = vf;
      p.losses[3] = ent;
      p.losses[4] = (float)(r[3] / ((double)T * (double)B));
      *p.ticket = 0u;
    }
  }
}

// ---------------------------------------------------------------------------
// a1: plain V-trace on log-probs (exact reference op order; exp in fp64 so the
// float result is correctly rounded).  One lane per env column, coalesced over b.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) vtrace_returns_kernel(const float* __restrict__ blp, const float* __restrict__ tlp,
                                                            const float* __restrict__ disc, const float* __restrict__ rew,
                                                            const float* __restrict__ val, const float* __restrict__ boot,
                                                            int T, int B, float clip_rho, float clip_pg,
                                                            float* __restrict__ vs, float* __restrict__ pg) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float acc = 0.f;
  float v_next = boot[b];
  float vs_next = boot[b];
  for (int t = T - 1; t >= 0; --t) {
    const long long g = (long long)t * B + b;
    const float rho = (float)exp((double)__fsub_rn(tlp[g], blp[g]));
    const float rhoc = clip_rho >= 0.f ? fminf(rho, clip_rho) : rho;
    const float rhopg = clip_pg >= 0.f ? fminf(rho, clip_pg) : rho;
    const float cs = fminf(rho, 1.0f);
    const float d = disc[g], r = rew[g], v = val[g];
    const float delta = __fmul_rn(rhoc, __fsub_rn(__fadd_rn(r, __fmul_rn(d, v_next)), v));
    acc = __fadd_rn(delta, __fmul_rn(__fmul_rn(d, cs), acc));
    const float vs_t = __fadd_rn(acc, v);
    pg[g] = __fmul_rn(rhopg, __fsub_rn(__fadd_rn(r, __fmul_rn(d, vs_next)), v));
    vs[g] = vs_t;
    vs_next = vs_t;
    v_next = v;
  }
}

template <int A_>
static int launch_vtrace_loss(const VtraceLossArgs& a, int layout, int grid, size_t smem, cudaStream_t st) {
  if (layout == RL_LAYOUT_ENV_MAJOR) {
    cudaFuncSetAttribute(vtrace_loss_kernel<A_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    vtrace_loss_kernel<A_, true><<<grid, kNT, smem, st>>>(a);
  } else {
    cudaFuncSetAttribute(vtrace_loss_kernel<A_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    vtrace_loss_kernel<A_, false><<<grid, kNT, smem, st>>>(a);
  }
  return 0;
}

}  // namespace rl

extern "C" size_t rl_loss_workspace_bytes(int n_cols) {
  const size_t grid = (size_t)(n_cols > 0 ? n_cols : 1);   // >= any kernel's CTA count
  return 256 + grid * 8 * sizeof(float);
}

extern "C" int rl_vtrace_from_importance_weights(const float* blp, const float* tlp, const float* discounts,
                                                 const float* rewards, const float* values, const float* bootstrap,
                                                 int T, int B, float clip_rho, float clip_pg, float* vs, float* pg,
                                                 rl_stream_t stream) {
  RL_CHECK_ARG(blp && tlp && discounts && rewards && values && bootstrap && vs && pg, "vtrace: null pointer");
  RL_CHECK_ARG(T > 0 && B > 0, "vtrace: T=%d B=%d must be positive", T, B);
  rl::vtrace_returns_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(blp, tlp, discounts, rewards, values,
                                                                                bootstrap, T, B, clip_rho, clip_pg, vs, pg);
  RL_CHECK_LAUNCH("rl_vtrace_from_importance_weights");
  return RL_OK;
}

extern "C" int rl_vtrace_loss_fwd_bwd(const float* target_logits, const float* behaviour_logits, const void* actions,
                                      int actions_i64, const float* rewards, const uint8_t* dones, const float* values,
                                      int T, int B, int A, int layout, float gamma, float clip_rho, float clip_pg,
                                      float vf_coeff, float ent_coeff, float* losses, float* d_logits, float* d_values,
                                      float* vs_out, float* pg_adv_out, void* workspace, size_t workspace_bytes,
                                      rl_stream_t stream) {
  using namespace rl;
  RL_CHECK_ARG(target_logits && behaviour_logits && actions && rewards && dones && values && losses && d_logits &&
                   d_values && workspace,
               "vtrace_loss: null pointer");
  RL_CHECK_ARG(T >= 2 && B >= 1 && A >= 1 && A <= 1024, "vtrace_loss: bad shape T=%d B=%d A=%d (need T>=2)", T, B, A);
  RL_CHECK_ARG(layout == RL_LAYOUT_TIME_MAJOR || layout == RL_LAYOUT_ENV_MAJOR, "vtrace_loss: bad layout %d", layout);
  const int grid = (B + kBW - 1) / kBW;
  if (workspace_bytes < rl_loss_workspace_bytes(B)) {
    set_error("vtrace_loss: workspace too small (%zu < %zu)", workspace_bytes, rl_loss_workspace_bytes(B));
    return RL_ERR_WORKSPACE;
  }
  // chunk of rows staged per pass: as many as fit ~30.5 KB (7 CTAs/SM) and kNT*kEPT elements
  const size_t row_bytes = (size_t)kBW * (2 * A + 2) * sizeof(float);
  int TC = (int)(30500 / row_bytes);
  TC = TC < 1 ? 1 : TC;
  if (TC > kNT * kEPT / kBW) TC = kNT * kEPT / kBW;
  if (TC >= T) TC = T; else TC &= ~3;                 // multi-chunk: keep chunk starts 16-byte aligned
  if (TC < 1) TC = 1;
  const size_t smem = (size_t)TC * row_bytes;
  RL_CHECK_ARG(smem <= 200 * 1024, "vtrace_loss: A=%d too large for the shared-memory tile", A);
  VtraceLossArgs a;
  a.tl = target_logits, a.bl = behaviour_logits, a.actions = actions, a.rewards = rewards, a.dones = dones;
  a.values = values, a.d_logits = d_logits, a.d_values = d_values, a.vs_out = vs_out, a.pg_out = pg_adv_out;
  a.losses = losses;
  a.ticket = reinterpret_cast<unsigned*>(workspace);
  a.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  a.T = T, a.B = B, a.A = A, a.TC = TC, a.act64 = actions_i64;
  a.gamma = gamma, a.clip_rho = clip_rho, a.clip_pg = clip_pg, a.vf_coeff = vf_coeff, a.ent_coeff = ent_coeff;
  const bool ptr_ok = aligned16(target_logits) && aligned16(behaviour_logits) && aligned16(d_logits);
  if (layout == RL_LAYOUT_TIME_MAJOR) {
    a.vec = ptr_ok && (((long long)B * A) % 4 == 0) && ((kBW * A) % 4 == 0);
  } else {
    a.vec = ptr_ok && (((long long)T * A) % 4 == 0) && (((long long)TC * A) % 4 == 0);
  }
  cudaStream_t st = (cudaStream_t)stream;
  switch (A) {
#define RL_CASE(N) case N: launch_vtrace_loss<N>(a, layout, grid, smem, st); break;
    RL_CASE(2) RL_CASE(3) RL_CASE(4) RL_CASE(5) RL_CASE(6) RL_CASE(7) RL_CASE(8) RL_CASE(9) RL_CASE(10) RL_CASE(12)
    RL_CASE(14) RL_CASE(16) RL_CASE(18)
#undef RL_CASE
    default: launch_vtrace_loss<0>(a, layout, grid, smem, st); break;
  }
  RL_CHECK_LAUNCH("rl_vtrace_loss_fwd_bwd");
  return RL_OK;
}
