// K1 — fused IMPALA loss: log-softmax + entropy + KL + V-trace backward scan +
// policy-gradient/value/entropy loss and its gradient, ONE launch.
//
// Reference arithmetic being replaced (PaddlePaddle/PARL):
//   parl/algorithms/paddle/impala/vtrace.py:99-139   V-trace recurrences
//   parl/algorithms/paddle/impala/impala.py:25-79    VTraceLoss (SUM reductions)
//   parl/algorithms/paddle/impala/impala.py:119-132  log-prob = sum(log_softmax * onehot)
//   parl/algorithms/paddle/impala/impala.py:148-208  row dropping / bootstrap / KL
//
// Layout in HBM: logits [T,B,A] f32, per-step scalars [T,B].  Two kernels:
//
//  * vtrace_loss_kernel (v4, the default): a CTA owns 4 adjacent env columns for ALL T rows (the scan never leaves
//    the CTA); the logits tiles [T, 4*A] move by 2-D TMA tensor maps (cp.async for env-major / unaligned shapes),
//    one (t,b) element per thread, array-free two-sweep softmax, warp-shuffle segmented suffix scan, gradient tile
//    written in place and TMA-stored, deterministic loss reduction (CTA partials -> last CTA, fp64).
//  * vtrace_loss_cta_kernel (v6, opt-in via rl_debug_set_vtrace_path(6)): same CTA shape, but every warp owns ONE
//    8-row TMA chunk (own mbarrier, own store) and the scan is composed from per-warp affine maps after a single
//    __syncthreads.  Round-2 measurements (profiles/r02_k1_*): 23.5 us vs 20.4 us for v4 at T=50, B=4096 — the
//    instruction count per element did not drop (777 vs 815 warp-instructions per warp) and 40 registers round up
//    to 6 CTAs per SM (1.15 waves); a warp-autonomous variant (v5: one warp per column block, 7 warps per SM) was
//    latency-bound at 24.8 us and was removed.  DESIGN.md section 4 has the analysis.
// Algorithmic traffic: (12A+17) bytes per kept (t,b) element.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "tma.cuh"

namespace rl {

constexpr int kBW = 4;     // env columns per CTA
constexpr int kNT = 224;   // threads per CTA: one (t,b) element per thread, TC <= kNT / kBW = 56 rows per chunk

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

// All index arithmetic is 32-bit (the host rejects T*B*A >= 2^31) and division-free:
// element i of a chunk is (t_local, b_local) = (i >> 2, i & 3) for both layouts.
template <bool EM>
__device__ __forceinline__ int gidx(int t, int b, int T, int B) {
  return EM ? b * T + t : t * B + b;
}
template <bool EM>
__device__ __forceinline__ int sidx(int tl, int bl, int TC) {
  return EM ? bl * TC + tl : tl * kBW + bl;
}

// global <-> shared tile copy, one segment per warp iteration, lanes stride the segment.
//   TM: nt row segments of nb*A floats (smem stride kBW*A);  EM: nb column segments of nt*A floats (stride TC*A).
template <bool EM, bool TO_SMEM>
__device__ __forceinline__ void copy_tile(float* s, const float* gsrc, float* gdst, int A, int T, int B, int TC,
                                          int t0, int nt, int b0, int nb, bool vec) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nseg = EM ? nb : nt;
  const int seglen = (EM ? nt : nb) * A;           // floats
  const int sstride = (EM ? TC : kBW) * A;
  const int gstride = (EM ? T : B) * A;            // floats between consecutive segments in global memory
  int goff = (EM ? (b0 * T + t0) : (t0 * B + b0)) * A + warp * gstride;
  float* sp = s + warp * sstride;
  for (int sg = warp; sg < nseg; sg += kNT / 32, goff += (kNT / 32) * gstride, sp += (kNT / 32) * sstride) {
    if (vec) {
      for (int k = lane * 4; k < seglen; k += 128) {
        if (TO_SMEM) cp_async16(sp + k, gsrc + goff + k);
        else *reinterpret_cast<float4*>(gdst + goff + k) = *reinterpret_cast<const float4*>(sp + k);
      }
    } else {
      for (int k = lane; k < seglen; k += 32) {
        if (TO_SMEM) cp_async4(sp + k, gsrc + goff + k);
        else gdst[goff + k] = sp[k];
      }
    }
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kL2E = 1.4426950408889634f;   // log2(e)
constexpr float kLN2 = 0.6931471805599453f;   // ln(2)

// Per-element softmax statistics in TWO array-free sweeps over the shared-memory rows (registers hold only
// running scalars, so 7 CTAs x 7 warps stay resident per SM):
//   sweep 1: m = max_j x_j, my = max_j y_j                                   (x target, y behaviour logits)
//   sweep 2: with xs_j = (x_j - m) log2e, e_j = 2^xs_j:  S = sum e_j, W = sum e_j xs_j, Y = sum e_j y_j,
//            Sy = sum 2^((y_j - my) log2e)
// from which  log2 p_j = xs_j - log2 S,  sum_j p_j log2 p_j = W/S - log2 S,  sum_j p_j y_j = Y/S.
// The tiles are left untouched; phase C recomputes p_j from x_j (one FFMA + MUFU.EX2 each).
struct SoftmaxStats {
  float nm;      // -m * log2e
  float l2S;     // log2 sum_j 2^xs_j
  float inv;     // 1 / S
  float la;      // log pi(a)      (natural log)
  float lma;     // log mu(a)
  float H;       // entropy of pi  (natural log)
  float KL;      // KL(pi || mu)
};

template <int A_>
__device__ __forceinline__ SoftmaxStats softmax_stats(const float* __restrict__ st, const float* __restrict__ sb, int A,
                                                      int act) {
  SoftmaxStats r;
  float m, my;
  if constexpr (A_ > 0 && (A_ & 1) == 0) {
    float2 a = *reinterpret_cast<const float2*>(st);
    float2 b = *reinterpret_cast<const float2*>(sb);
    m = fmaxf(a.x, a.y), my = fmaxf(b.x, b.y);
#pragma unroll 4
    for (int j = 2; j < A_; j += 2) {
      a = *reinterpret_cast<const float2*>(st + j);
      b = *reinterpret_cast<const float2*>(sb + j);
      m = fmaxf(m, fmaxf(a.x, a.y));
      my = fmaxf(my, fmaxf(b.x, b.y));
    }
  } else {
    m = st[0], my = sb[0];
    for (int j = 1; j < A; ++j) m = fmaxf(m, st[j]), my = fmaxf(my, sb[j]);
  }
  const float nm = -m * kL2E, nmy = -my * kL2E;
  float S0 = 0.f, S1 = 0.f, W0 = 0.f, W1 = 0.f, Y0 = 0.f, Y1 = 0.f, Sy0 = 0.f, Sy1 = 0.f;
  if constexpr (A_ > 0 && (A_ & 1) == 0) {
#pragma unroll 3
    for (int j = 0; j < A_; j += 2) {
      const float2 a = *reinterpret_cast<const float2*>(st + j);
      const float2 b = *reinterpret_cast<const float2*>(sb + j);
      const float xs0 = fmaf(a.x, kL2E, nm), xs1 = fmaf(a.y, kL2E, nm);
      const float e0 = ex2_approx(xs0), e1 = ex2_approx(xs1);
      S0 += e0, S1 += e1;
      W0 = fmaf(e0, xs0, W0), W1 = fmaf(e1, xs1, W1);
      Y0 = fmaf(e0, b.x, Y0), Y1 = fmaf(e1, b.y, Y1);
      Sy0 += ex2_approx(fmaf(b.x, kL2E, nmy));
      Sy1 += ex2_approx(fmaf(b.y, kL2E, nmy));
    }
  } else {
    for (int j = 0; j < A; ++j) {
      const float xs0 = fmaf(st[j], kL2E, nm);
      const float e0 = ex2_approx(xs0);
      S0 += e0;
      W0 = fmaf(e0, xs0, W0);
      Y0 = fmaf(e0, sb[j], Y0);
      Sy0 += ex2_approx(fmaf(sb[j], kL2E, nmy));
    }
  }
  const float S = S0 + S1, W = W0 + W1, Y = Y0 + Y1, Sy = Sy0 + Sy1;
  const float l2S = lg2_approx(S);
  const float inv = __fdividef(1.0f, S);
  const float logSy = lg2_approx(Sy) * kLN2;
  const float Hn = (W * inv - l2S) * kLN2;            // sum_j p_j log p_j
  r.nm = nm, r.l2S = l2S, r.inv = inv;
  r.H = -Hn;
  r.KL = Hn - Y * inv + my + logSy;                   // sum_j p_j (log p_j - log q_j)
  r.la = (fmaf(st[act], kL2E, nm) - l2S) * kLN2;
  r.lma = sb[act] - my - logSy;
  return r;
}

// Shared-memory carve-up (bytes): [tile_tl | pad to 128][tile_bl | pad to 128][s_acc][s_kc]
__host__ __device__ inline int tile_bytes_padded(int TC, int A) { return (TC * kBW * A * 4 + 127) & ~127; }

template <int A_, bool EM, bool TMA>
__global__ void __launch_bounds__(kNT, 7) vtrace_loss_kernel(const VtraceLossArgs p,
                                                            const __grid_constant__ CUtensorMap map_tl,
                                                            const __grid_constant__ CUtensorMap map_bl,
                                                            const __grid_constant__ CUtensorMap map_dl) {
  static_assert(!(TMA && EM), "the TMA tile path is time-major only");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int A = A_ > 0 ? A_ : p.A;
  const int T = p.T, B = p.B, TC = p.TC;
  const int tile_pad = tile_bytes_padded(TC, A);
  float* s_tl = reinterpret_cast<float*>(smem_raw);
  float* s_bl = reinterpret_cast<float*>(smem_raw + tile_pad);
  float* s_acc = reinterpret_cast<float*>(smem_raw + 2 * tile_pad);   // delta, then acc = vs - V
  float* s_kc = s_acc + TC * kBW;                                     // gamma_t * min(rho, 1)
  __shared__ float s_carry[kBW];        // acc at the first row of the chunk processed before (later in time)
  __shared__ float s_red[4][kNT / 32];
  __shared__ bool s_last;
  __shared__ __align__(8) unsigned long long s_mbar;
  uint32_t mbar_phase = 0;
  const int tid = threadIdx.x;
  if (TMA) {
    if (tid == 0) {
      tma_prefetch_desc(&map_tl);
      tma_prefetch_desc(&map_bl);
      tma_prefetch_desc(&map_dl);
      mbar_init(&s_mbar, 1);
      fence_mbar_init();
    }
  }
  if (tid < kBW) s_carry[tid] = 0.f;
  __syncthreads();

  const int b0 = blockIdx.x * kBW;
  const int nb = min(kBW, B - b0);
  // this thread's element of every chunk: (t_local, b_local) = (tid >> 2, tid & 3)
  const int tl_ = tid >> 2, bl_ = tid & 3;
  const int si = sidx<EM>(tl_, bl_, TC);
  float sum_pi = 0.f, sum_vf = 0.f, sum_ent = 0.f, sum_kl = 0.f;
  float acc_carry = 0.f;

  const int nchunks = (T + TC - 1) / TC;
  for (int c = nchunks - 1; c >= 0; --c) {
    const int t0 = c * TC;
    const int nt = min(TC, T - t0);
    const bool vec = p.vec && (((EM ? nt : nb) * A) & 3) == 0;
    if (TMA) {
      // one elected thread: two 2-D TMA tile loads (box = TC rows x kBW*A floats; rows/columns past the
      // tensor edge are zero-filled by the hardware), completion signalled on the mbarrier
      if (tid == 0) {
        mbar_arrive_expect_tx(&s_mbar, 2u * (uint32_t)(TC * kBW * A * 4));
        tma_load_2d(s_tl, &map_tl, b0 * A, t0, &s_mbar);
        tma_load_2d(s_bl, &map_bl, b0 * A, t0, &s_mbar);
      }
    } else {
      copy_tile<EM, true>(s_tl, p.tl, nullptr, A, T, B, TC, t0, nt, b0, nb, vec);
      copy_tile<EM, true>(s_bl, p.bl, nullptr, A, T, B, TC, t0, nt, b0, nb, vec);
      cp_async_commit();
    }

    // ---- per-element scalars straight from global (overlaps the tile copy) ----
    const bool valid = tl_ < nt && bl_ < nb;
    const int t = t0 + tl_;
    const bool loss_row = valid && t < T - 1;
    const int g = gidx<EM>(t, b0 + bl_, T, B);
    int act = 0;
    float e_r = 0.f, e_v = 0.f, e_vn = 0.f, e_g = 0.f;
    if (valid) {
      act = p.act64 ? (int)reinterpret_cast<const long long*>(p.actions)[g] : reinterpret_cast<const int*>(p.actions)[g];
      e_r = p.rewards[g];
      e_v = p.values[g];
      e_g = p.dones[g] ? 0.0f : p.gamma;                  // impala.py:59  (~dones) * discount
      if (t + 1 < T) e_vn = p.values[g + (EM ? 1 : B)];
    }
    if (TMA) {
      mbar_wait(&s_mbar, mbar_phase);
      mbar_phase ^= 1u;
    } else {
      cp_async_wait<0>();
      __syncthreads();
    }

    // ---- phase A: softmax / entropy / KL / rho / delta (element-private, tiles read-only) ----
    SoftmaxStats ss;
    float e_rpg = 0.f;
    ss.nm = 0.f, ss.l2S = 0.f, ss.inv = 0.f, ss.la = 0.f, ss.lma = 0.f, ss.H = 0.f, ss.KL = 0.f;
    if (valid) {
      ss = softmax_stats<A_>(s_tl + si * A, s_bl + si * A, A, act);
      sum_kl += ss.KL;                                    // impala.py:160-162: every row
      if (loss_row) {
        const float rho = expf(ss.la - ss.lma);           // vtrace.py:101-103
        const float rhoc = p.clip_rho >= 0.f ? fminf(rho, p.clip_rho) : rho;
        const float cs = fminf(rho, 1.0f);                // :109
        e_rpg = p.clip_pg >= 0.f ? fminf(rho, p.clip_pg) : rho;
        // deltas = clipped_rhos * (rewards + discounts * values_t_plus_1 - values)   :115
        const float td = __fsub_rn(__fadd_rn(e_r, __fmul_rn(e_g, e_vn)), e_v);
        s_acc[si] = __fmul_rn(rhoc, td);
        s_kc[si] = __fmul_rn(e_g, cs);
        sum_ent += ss.H;
      }
    }
    __syncthreads();

    // ---- phase B: backward-in-time scan acc_t = delta_t + k_t * acc_{t+1} as a warp-shuffle segmented prefix:
    //      warp 0 = 4 columns x 8 time segments; each lane composes the affine maps of its rows, a 3-step
    //      shuffle suffix-scan combines the segments (k_t = 0 at episode cuts makes the prefix segmented),
    //      then each lane replays its rows with the incoming accumulator.
    if (tid < 32) {
      const int col = tid & 3, seg = tid >> 2;                 // lane = seg*4 + col
      const int nrows = min(nt, T - 1 - t0);                   // loss rows of this chunk (bootstrap row skipped)
      const int per = (nrows + 7) >> 3;
      const int lo = min(seg * per, nrows), hi = min(lo + per, nrows);     // this lane's rows [lo, hi)
      const int step = EM ? 1 : kBW;
      const bool colok = col < nb;
      float D = 0.f, K = 1.f;
      if (colok) {
        int sj = sidx<EM>(hi - 1, col, TC);
        for (int q = hi - 1; q >= lo; --q, sj -= step) {
          const float k = s_kc[sj];
          D = fmaf(k, D, s_acc[sj]);
          K *= k;
        }
      }
      // inclusive suffix scan over segments (later time = higher seg): F_seg o F_{seg+off}
#pragma unroll
      for (int off = 1; off < 8; off <<= 1) {
        const float Dn = __shfl_down_sync(0xffffffffu, D, off * 4);
        const float Kn = __shfl_down_sync(0xffffffffu, K, off * 4);
        if (seg + off < 8) {
          D = fmaf(K, Dn, D);
          K *= Kn;
        }
      }
      // accumulator entering this lane's segment = composite of all later segments applied to the carry
      const float Dx = __shfl_down_sync(0xffffffffu, D, 4), Kx = __shfl_down_sync(0xffffffffu, K, 4);
      const float carry_in = __shfl_sync(0xffffffffu, acc_carry, col);       // lanes 0..3 hold the column carries
      float acc = seg < 7 ? fmaf(Kx, carry_in, Dx) : carry_in;
      if (colok) {
        int sj = sidx<EM>(hi - 1, col, TC);
        for (int q = hi - 1; q >= lo; --q, sj -= step) {
          acc = fmaf(s_kc[sj], acc, s_acc[sj]);
          s_acc[sj] = acc;
        }
      }
      // lanes 0..3 (segment 0) end on row 0 of the chunk: that is the carry for the next (earlier) chunk
      if (seg == 0) acc_carry = acc;
    }
    __syncthreads();

    // ---- phase C: advantages, losses, gradient tile (written over the target-logit tile) ----
    if (valid) {
      float* pt = s_tl + si * A;
      if (loss_row) {
        const float acc_n = (t + 1 == T - 1) ? 0.f : (tl_ + 1 < nt ? s_acc[sidx<EM>(tl_ + 1, bl_, TC)] : s_carry[bl_]);
        const float vs = __fadd_rn(s_acc[si], e_v);                    // vtrace.py:125
        const float vs_n = __fadd_rn(acc_n, e_vn);                     // :128-129 (bootstrap at the end)
        const float adv = __fmul_rn(e_rpg, __fsub_rn(__fadd_rn(e_r, __fmul_rn(e_g, vs_n)), e_v));   // :136-137
        const float dv = e_v - vs;
        sum_pi -= ss.la * adv;                                          // impala.py:67-68
        sum_vf += 0.5f * dv * dv;                                       // :71-72
        p.d_values[g] = p.vf_coeff * dv;
        if (p.vs_out) p.vs_out[t * B + b0 + bl_] = vs;
        if (p.pg_out) p.pg_out[t * B + b0 + bl_] = adv;
        // dL/dz_j = p_j (adv - c_e (H + log p_j)) - adv [j == a],  log p_j = ln2 (xs_j - log2 S)
        const float ce2 = p.ent_coeff * kLN2;
        const float c0 = fmaf(ce2, ss.l2S, adv - p.ent_coeff * ss.H) * ss.inv;   // folded with 1/S
        const float c1 = -ce2 * ss.inv;
        if constexpr (A_ > 0 && (A_ & 1) == 0) {
#pragma unroll 3
          for (int j = 0; j < A_; j += 2) {
            const float2 a = *reinterpret_cast<const float2*>(pt + j);
            const float xs0 = fmaf(a.x, kL2E, ss.nm), xs1 = fmaf(a.y, kL2E, ss.nm);
            const float d0 = ex2_approx(xs0) * fmaf(c1, xs0, c0), d1 = ex2_approx(xs1) * fmaf(c1, xs1, c0);
            *reinterpret_cast<float2*>(pt + j) = make_float2(d0, d1);
          }
        } else {
          for (int j = 0; j < A; ++j) {
            const float xs0 = fmaf(pt[j], kL2E, ss.nm);
            pt[j] = ex2_approx(xs0) * fmaf(c1, xs0, c0);
          }
        }
        pt[act] -= adv;
      } else {
        p.d_values[g] = 0.f;                                            // bootstrap row: no gradient
        for (int j = 0; j < A; ++j) pt[j] = 0.f;
      }
    }
    if (TMA) {
      fence_proxy_async_smem();          // generic-proxy writes of the gradient tile -> visible to the TMA engine
      __syncthreads();
      if (tid < nb) s_carry[tid] = acc_carry;
      if (tid == 0) {
        tma_store_2d(&map_dl, b0 * A, t0, s_tl);
        tma_store_commit();
        tma_store_wait_read();           // smem tile may be overwritten / the CTA may exit afterwards
      }
      if (c > 0) __syncthreads();
    } else {
      __syncthreads();
      if (tid < nb) s_carry[tid] = acc_carry;
      copy_tile<EM, false>(s_tl, nullptr, p.d_logits, A, T, B, TC, t0, nt, b0, nb, vec);
      if (c > 0) __syncthreads();
    }
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
    __threadfence();
  }
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

// ---------------------------------------------------------------------------
// v6: CTA per 4-column block, ONE 8-row pass per warp (the default fast path).  v5 showed that a warp walking
// all of T alone is bound by its own instruction latency (7 warps per SM cannot hide LDS -> MUFU -> FADD chains),
// so the rows of a column block are spread over ceil(T/8) warps: every warp waits only for ITS 8-row TMA chunk,
// computes the softmax statistics of its 32 elements (two register-light sweeps: 7 CTAs = 49 warps stay resident
// per SM), reduces its rows to one affine map per column (3 shuffle steps), and after the CTA's single
// __syncthreads composes the maps of the later warps (<= 6 FMAs) to get its incoming accumulator.  The gradient
// rows are recomputed in place and each warp TMA-stores its own chunk at once.
// ---------------------------------------------------------------------------
constexpr int kV6Rows = 8;                 // rows per warp / per TMA chunk
constexpr int kV6MaxWarps = 7;             // T <= 56

// packed fp32 pairs (Blackwell FFMA2 / FADD2 / FMUL2): two lanes of arithmetic per issue slot
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
// mbarrier wait that lets the hardware suspend the warp for up to ~1 ms per probe instead of spinning
__device__ __forceinline__ void mbar_wait_suspend(void* mbar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT_S:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra LAB_DONE_S;\n"
      "bra LAB_WAIT_S;\n"
      "LAB_DONE_S:\n"
      "}\n" ::"r"(smem_u32(mbar)),
      "r"(parity), "r"(0x000f4240)
      : "memory");
}

template <int A_>
__global__ void __maxnreg__(32)          // 7 CTAs x 7 warps x 32 registers: one wave of 1024 CTAs on 148 SMs
    vtrace_loss_cta_kernel(const VtraceLossArgs p, const __grid_constant__ CUtensorMap map_tl,
                           const __grid_constant__ CUtensorMap map_bl, const __grid_constant__ CUtensorMap map_dl,
                           const __grid_constant__ CUtensorMap map_tl_tail, const __grid_constant__ CUtensorMap map_bl_tail,
                           const __grid_constant__ CUtensorMap map_dl_tail) {
  static_assert(A_ >= 2 && (A_ & 1) == 0, "v6 needs an even compile-time A");
  constexpr int CW = 4, R = kV6Rows;
  constexpr uint32_t FULL = 0xffffffffu;
  constexpr int kChunkBytes = R * CW * A_ * 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long s_bar[kV6MaxWarps];
  __shared__ float2 s_comp[kV6MaxWarps][CW];
  __shared__ float s_red[4][kV6MaxWarps];
  __shared__ bool s_last;
  const int T = p.T, B = p.B;
  const int P = (T + R - 1) / R, nfull = T / R, tail_rows = T - nfull * R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * CW;
  unsigned char* s_tl = smem_raw;
  unsigned char* s_bl = smem_raw + (((size_t)T * CW * A_ * 4 + 127) & ~(size_t)127);   // TMA tiles: 128-byte aligned

  if (tid == 0) {
    for (int c = 0; c < P; ++c) mbar_init(&s_bar[c], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int c = P - 1; c >= 0; --c) {                  // latest rows first: they head the dependency chain
      const bool full = c < nfull;
      const uint32_t bytes = (uint32_t)((full ? R : tail_rows) * CW * A_ * 4);
      mbar_arrive_expect_tx(&s_bar[c], 2u * bytes);
      tma_load_2d(s_tl + c * kChunkBytes, full ? &map_tl : &map_tl_tail, b0 * A_, c * R, &s_bar[c]);
      tma_load_2d(s_bl + c * kChunkBytes, full ? &map_bl : &map_bl_tail, b0 * A_, c * R, &s_bar[c]);
    }
  }
  // ---- this lane's element: (t, b) = (warp*8 + lane/4, b0 + lane%4).  Lanes past the last row are CLAMPED onto
  //      row T-1 (they read valid memory and compute throw-away values) so that the hot path has no branches;
  //      fv / fl are the 0/1 masks "row exists" / "row carries a loss term".
  const int r = lane >> 2, c = lane & 3;
  const int t_raw = warp * R + r;
  const int t = min(t_raw, T - 1);
  const float fv = t_raw < T ? 1.f : 0.f;
  const float fl = t_raw < T - 1 ? 1.f : 0.f;
  const int g = t * B + b0 + c;
  const int act = p.act64 ? (int)reinterpret_cast<const long long*>(p.actions)[g] : reinterpret_cast<const int*>(p.actions)[g];
  const float e_r = p.rewards[g];
  const float e_v = p.values[g];
  const float e_g = p.dones[g] ? 0.0f : p.gamma;          // impala.py:59  (~dones) * discount
  const float e_vn = p.values[min(g + B, (T - 1) * B + b0 + c)];
  mbar_wait_suspend(&s_bar[warp], 0);

  // ---- phase A: softmax statistics.  Sweep 1: maxima.  Sweep 2 (packed pairs): exponentials, S, W = sum e xs,
  //      Y = sum e y, Sy; the target exponentials are parked over the (dead) behaviour logits for phase C.
  float* pt = reinterpret_cast<float*>(s_tl) + (size_t)(t * CW + c) * A_;
  float* pb = reinterpret_cast<float*>(s_bl) + (size_t)(t * CW + c) * A_;
  float m = -INFINITY, my = -INFINITY;
#pragma unroll
  for (int j = 0; j < A_; j += 2) {
    const float2 a = *reinterpret_cast<const float2*>(pt + j);
    const float2 y = *reinterpret_cast<const float2*>(pb + j);
    m = fmaxf(m, fmaxf(a.x, a.y));
    my = fmaxf(my, fmaxf(y.x, y.y));
  }
  const float nm = -m * kL2E, nmy = -my * kL2E;
  const float x_act = pt[act], y_act = pb[act];
  const float2 cL = make_float2(kL2E, kL2E), cnm = make_float2(nm, nm), cnmy = make_float2(nmy, nmy);
  float2 S2 = make_float2(0.f, 0.f), W2 = S2, Y2 = S2, Sy2 = S2;
#pragma unroll
  for (int j = 0; j < A_; j += 2) {
    const float2 a = *reinterpret_cast<const float2*>(pt + j);
    const float2 y = *reinterpret_cast<const float2*>(pb + j);
    const float2 xs = f2_fma(a, cL, cnm);
    const float2 e = make_float2(ex2_approx(xs.x), ex2_approx(xs.y));
    const float2 ys = f2_fma(y, cL, cnmy);
    const float2 ey = make_float2(ex2_approx(ys.x), ex2_approx(ys.y));
    S2 = f2_add(S2, e);
    W2 = f2_fma(e, xs, W2);
    Y2 = f2_fma(e, y, Y2);
    Sy2 = f2_add(Sy2, ey);
    *reinterpret_cast<float2*>(pb + j) = e;              // this lane's own row: no cross-lane hazard
  }
  const float S = S2.x + S2.y, Wt = W2.x + W2.y, Y = Y2.x + Y2.y, Sy = Sy2.x + Sy2.y;
  const float l2S = lg2_approx(S);
  const float inv = __fdividef(1.0f, S);
  const float logSy = lg2_approx(Sy) * kLN2;
  const float Hn = (Wt * inv - l2S) * kLN2;              // sum_j p_j log p_j
  const float H = -Hn;
  float sum_kl = fv * (Hn - Y * inv + my + logSy);       // impala.py:160-162: every row
  const float la = (fmaf(x_act, kL2E, nm) - l2S) * kLN2;
  const float lma = y_act - my - logSy;
  const float rho = ex2_approx((la - lma) * kL2E);       // vtrace.py:101-103
  const float rhoc = p.clip_rho >= 0.f ? fminf(rho, p.clip_rho) : rho;
  const float rpg = p.clip_pg >= 0.f ? fminf(rho, p.clip_pg) : rho;
  float D = fl * __fmul_rn(rhoc, __fsub_rn(__fadd_rn(e_r, __fmul_rn(e_g, e_vn)), e_v));     // :115 (0 past T-2)
  float K = fl * __fmul_rn(e_g, fminf(rho, 1.0f));                                          // :109
  float sum_ent = fl * H;
  // suffix scan of the affine maps acc -> D + K acc over the warp's 8 rows (later time = higher lane)
#pragma unroll
  for (int off = CW; off < 32; off <<= 1) {
    const float Dn = __shfl_down_sync(FULL, D, off);
    const float Kn = __shfl_down_sync(FULL, K, off);
    const bool in = lane + off < 32;
    D = in ? fmaf(K, Dn, D) : D;
    K = in ? K * Kn : K;
  }
  if (lane < CW) s_comp[warp][lane] = make_float2(D, K);   // the whole pass as one map, per column
  __syncthreads();
  // ---- phase B: accumulator entering this warp's rows = later warps' maps applied to 0, latest first
  float cin = 0.f;
  for (int w2 = P - 1; w2 > warp; --w2) {
    const float2 m2 = s_comp[w2][c];
    cin = fmaf(m2.y, cin, m2.x);
  }
  const float acc = fmaf(K, cin, D);
  float acc_n = __shfl_down_sync(FULL, acc, CW);
  acc_n = r == R - 1 ? cin : acc_n;
  // ---- phase C: advantages, losses, gradient row (exponentials read back, no second MUFU pass)
  const float vs = __fadd_rn(acc, e_v);                              // vtrace.py:125
  const float vs_n = __fadd_rn(acc_n, e_vn);                         // :128-129 (bootstrap at the end)
  const float adv = fl * __fmul_rn(rpg, __fsub_rn(__fadd_rn(e_r, __fmul_rn(e_g, vs_n)), e_v));   // :136-137
  const float dv = fl * (e_v - vs);
  float sum_pi = -la * adv;                                           // impala.py:67-68
  float sum_vf = 0.5f * dv * dv;                                      // :71-72
  if (fv != 0.f) {
    p.d_values[g] = p.vf_coeff * dv;
    if (p.vs_out) p.vs_out[g] = vs;
    if (p.pg_out) p.pg_out[g] = adv;
  }
  {
    const float ce2 = fl * p.ent_coeff * kLN2;
    const float c0 = fmaf(ce2, l2S, adv - fl * p.ent_coeff * H) * inv;   // folded with 1/S; 0 on the bootstrap row
    const float c1 = -ce2 * inv;
    const float2 c02 = make_float2(c0, c0), c12 = make_float2(c1, c1);
#pragma unroll
    for (int j = 0; j < A_; j += 2) {
      const float2 a = *reinterpret_cast<const float2*>(pt + j);
      const float2 e = *reinterpret_cast<const float2*>(pb + j);
      const float2 xs = f2_fma(a, cL, cnm);
      *reinterpret_cast<float2*>(pt + j) = f2_mul(e, f2_fma(c12, xs, c02));
    }
    pt[act] -= adv;
  }
  fence_proxy_async_smem();              // generic-proxy writes of the gradient rows -> visible to the TMA engine
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(warp < nfull ? &map_dl : &map_dl_tail, b0 * A_, warp * R, s_tl + warp * kChunkBytes);
    tma_store_commit();
  }
  // ---- loss reduction: warp -> CTA -> (last CTA) grid, fixed order, fp64 at the end
  sum_pi = warp_sum(sum_pi), sum_vf = warp_sum(sum_vf), sum_ent = warp_sum(sum_ent), sum_kl = warp_sum(sum_kl);
  if (lane == 0) {
    s_red[0][warp] = sum_pi, s_red[1][warp] = sum_vf, s_red[2][warp] = sum_ent, s_red[3][warp] = sum_kl;
    tma_store_wait_read();               // the shared-memory rows must outlive the bulk read
  }
  __syncthreads();
  if (tid < 4) {
    float a = 0.f;
    for (int x = 0; x < P; ++x) a += s_red[tid][x];
    p.partials[blockIdx.x * 4 + tid] = a;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    __threadfence();
    double accd[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = tid; i < (int)gridDim.x; i += blockDim.x) {
#pragma unroll
      for (int q = 0; q < 4; ++q) accd[q] += (double)__ldcg(p.partials + i * 4 + q);
    }
    __shared__ double s_dred[4][kV6MaxWarps];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double v = accd[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
      if (lane == 0) s_dred[q][warp] = v;
    }
    __syncthreads();
    if (tid == 0) {
      double rr[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        rr[q] = 0.0;
        for (int x = 0; x < P; ++x) rr[q] += s_dred[q][x];
      }
      const float pi = (float)rr[0], vf = (float)rr[1], ent = (float)rr[2];
      p.losses[0] = pi + vf * p.vf_coeff + ent * p.ent_coeff;      // impala.py:78-79
      p.losses[1] = pi;
      p.losses[2] = vf;
      p.losses[3] = ent;
      p.losses[4] = (float)(rr[3] / ((double)T * (double)B));
      *p.ticket = 0u;
    }
  }
}

template <int A_>
static bool try_launch_v6(const VtraceLossArgs& a, const float* tl, const float* bl, float* dl, cudaStream_t st) {
  if constexpr (A_ >= 2 && (A_ & 1) == 0 && A_ <= 64) {
    constexpr int CW = 4, R = kV6Rows;
    const int T = a.T, B = a.B;
    if (B % CW != 0 || T > R * kV6MaxWarps || ((CW * A_ * 4) % 16) != 0 || CW * A_ > 256) return false;
    if ((R * CW * A_ * 4) % 128 != 0) return false;          // every chunk starts 128-byte aligned in shared memory
    const int P = (T + R - 1) / R, nfull = T / R, tail = T - nfull * R;
    alignas(64) CUtensorMap maps[6];
    const char* err = nullptr;
    const uint64_t pitch = (uint64_t)B * A_ * sizeof(float);
    const float* bases[3] = {tl, bl, dl};
    for (int i = 0; i < 3; ++i) {
      if (make_tensor_map_2d_f32(&maps[i], bases[i], (uint64_t)B * A_, (uint64_t)T, pitch, CW * A_, R, &err)) return false;
      if (make_tensor_map_2d_f32(&maps[3 + i], bases[i], (uint64_t)B * A_, (uint64_t)T, pitch, CW * A_, tail > 0 ? tail : R,
                                 &err))
        return false;
    }
    const size_t smem = 2 * (((size_t)T * CW * A_ * 4 + 127) & ~(size_t)127);
    if (smem > 200 * 1024) return false;
    RL_SMEM_OPTIN(vtrace_loss_cta_kernel<A_>);
    static bool carve = false;
    if (!carve) {
      cudaFuncSetAttribute(vtrace_loss_cta_kernel<A_>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
      carve = true;
    }
    vtrace_loss_cta_kernel<A_><<<B / CW, P * 32, smem, st>>>(a, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5]);
    return true;
  }
  return false;
}

static int g_v6_enable = 0;      // rl_debug_set_vtrace_path(6): opt in to the v6 kernel

template <int A_>
static int launch_vtrace_loss(const VtraceLossArgs& a, int layout, bool tma, const CUtensorMap* maps, int grid,
                              size_t smem, cudaStream_t st) {
  if (layout == RL_LAYOUT_ENV_MAJOR) {
    RL_SMEM_OPTIN(vtrace_loss_kernel<A_, true, false>);
    vtrace_loss_kernel<A_, true, false><<<grid, kNT, smem, st>>>(a, maps[0], maps[1], maps[2]);
  } else if (tma) {
    RL_SMEM_OPTIN(vtrace_loss_kernel<A_, false, true>);
    cudaFuncSetAttribute(vtrace_loss_kernel<A_, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    vtrace_loss_kernel<A_, false, true><<<grid, kNT, smem, st>>>(a, maps[0], maps[1], maps[2]);
  } else {
    RL_SMEM_OPTIN(vtrace_loss_kernel<A_, false, false>);
    vtrace_loss_kernel<A_, false, false><<<grid, kNT, smem, st>>>(a, maps[0], maps[1], maps[2]);
  }
  return 0;
}

}  // namespace rl

static bool g_disable_tma = false;
// Test / triage hook: force the cp.async tile path (1) or allow the TMA path (0).
extern "C" int rl_debug_set_tma(int disable) {
  g_disable_tma = disable != 0;
  return RL_OK;
}

// Triage hook: 0 = default (the CTA-per-4-columns kernel with block-level phases, v4: the fastest measured path at
// every shape, profiles/r02_k1_matrix_b.jsonl), 6 = the v6 kernel (one 8-row TMA chunk per warp, single block sync).
extern "C" int rl_debug_set_vtrace_path(int mode) {
  if (mode != 0 && mode != 6) {
    rl::set_error("rl_debug_set_vtrace_path: mode %d not in {0, 6}", mode);
    return RL_ERR_BAD_ARG;
  }
  rl::g_v6_enable = mode == 6;
  return RL_OK;
}

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
  RL_CHECK_ARG((long long)T * B * A < (1LL << 31), "vtrace_loss: T*B*A must be < 2^31 (32-bit indexing)");
  const int grid = (B + kBW - 1) / kBW;
  if (workspace_bytes < rl_loss_workspace_bytes(B)) {
    set_error("vtrace_loss: workspace too small (%zu < %zu)", workspace_bytes, rl_loss_workspace_bytes(B));
    return RL_ERR_WORKSPACE;
  }
  // chunk of rows staged per pass: as many as fit ~30.5 KB (7 CTAs/SM), one element per thread
  const size_t row_bytes = (size_t)kBW * (2 * A + 2) * sizeof(float);
  int TC = (int)(30500 / row_bytes);
  TC = TC < 1 ? 1 : TC;
  if (TC > kNT / kBW) TC = kNT / kBW;               // one (t,b) element per thread
  if (TC >= T) TC = T; else TC &= ~3;                 // multi-chunk: keep chunk starts 16-byte aligned
  if (TC < 1) TC = 1;
  const size_t smem = 2 * (size_t)tile_bytes_padded(TC, A) + (size_t)TC * kBW * 2 * sizeof(float);
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
  // TMA tile path: time-major, 16-byte aligned bases, row pitch a multiple of 16 bytes, box <= 256 elements
  alignas(64) CUtensorMap maps[3];
  memset(maps, 0, sizeof(maps));
  bool tma = layout == RL_LAYOUT_TIME_MAJOR && ptr_ok && (((long long)B * A) % 4 == 0) && kBW * A <= 256 && TC <= 256 &&
             !g_disable_tma;
  if (tma) {
    const char* err = nullptr;
    const uint64_t pitch = (uint64_t)B * A * sizeof(float);
    if (make_tensor_map_2d_f32(&maps[0], target_logits, (uint64_t)B * A, (uint64_t)T, pitch, kBW * A, TC, &err) ||
        make_tensor_map_2d_f32(&maps[1], behaviour_logits, (uint64_t)B * A, (uint64_t)T, pitch, kBW * A, TC, &err) ||
        make_tensor_map_2d_f32(&maps[2], d_logits, (uint64_t)B * A, (uint64_t)T, pitch, kBW * A, TC, &err)) {
      tma = false;                       // fall back to the cp.async tile path
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool v6_ok = tma && g_v6_enable;
  switch (A) {
#define RL_CASE(N)                                                                                     \
  case N:                                                                                              \
    if (!(v6_ok && try_launch_v6<N>(a, target_logits, behaviour_logits, d_logits, st)))                 \
      launch_vtrace_loss<N>(a, layout, tma, maps, grid, smem, st);                                      \
    break;
    RL_CASE(2) RL_CASE(3) RL_CASE(4) RL_CASE(5) RL_CASE(6) RL_CASE(7) RL_CASE(8) RL_CASE(9) RL_CASE(10) RL_CASE(12)
    RL_CASE(14) RL_CASE(16) RL_CASE(18)
#undef RL_CASE
    default: launch_vtrace_loss<0>(a, layout, tma, maps, grid, smem, st); break;
  }
  RL_CHECK_LAUNCH("rl_vtrace_loss_fwd_bwd");
  return RL_OK;
}
