// K6 for the MLP model family — whole-network forward and backward in ONE launch each, fp32 on the CUDA cores.
//
// Models being replaced (PaddlePaddle/PARL; executed there by torch/paddle eager, two trunk passes per learn):
//   benchmark/torch/ppo/mujoco_model.py:27-53       17 -> 64 -> 64 (tanh) -> {policy mean 6, value 1}
//   benchmark/torch/QuickStart/cartpole_model.py:21-38   4 -> 20 (tanh) -> 2 (softmax outside)
//   examples/DQN/cartpole_model.py:21-41            4 -> 128 -> 128 (relu) -> 2
// The networks are a few thousand parameters: every weight lives in shared memory for the whole kernel, a CTA
// walks over tiles of 64 samples, activations of a tile never leave shared memory, and the backward kernel
// RECOMPUTES the forward activations from the observations instead of reading saved ones (68 B of input per
// sample instead of 0.6 KB of activations).  fp32 throughout: the reference is fp32 and the parity bar is 1e-4.
//
// Thread mapping (256 threads, tile = 64 samples, activations stored [feature][sample] with row stride 68):
//   forward layer      thread (ty = tid/16, tx = tid%16) -> outputs j = ty*JT .. +JT-1 (JT = out_p/16) of samples
//                      4tx .. 4tx+3; weights [out_p][in_p] read as broadcast float4 along k, activations as float4
//                      along the sample axis: 16*JT FMA per (JT + 4) LDS.128
//   delta propagation  same register tile with the roles of j and k swapped (W read along its rows)
//   weight gradient    thread owns the 16 entries j in {jq + a*Gj}, k in {kq + b*G}; both operands are float4 along
//                      the sample axis (conflict-free: consecutive lanes = consecutive rows), 64 FMA per 8 LDS.128;
//                      per-CTA partial sums live in an L2-resident workspace, a second kernel adds them up in a
//                      fixed order (deterministic)
#include <string.h>

#include "common.cuh"
#include "env_common.cuh"

namespace rl {

constexpr int kMlpThreads = 256;
constexpr int kTN = 64;          // samples per tile
constexpr int kTS = 68;          // shared-memory row stride (floats) of one feature row of a tile
constexpr int kMaxLayers = 4;
constexpr int kMaxSeg = 8;
constexpr int kMaxWidth = 128;   // padded layer width limit

struct MlpSeg {
  const float* w;   // [rows, in]  row-major (torch nn.Linear.weight)
  const float* b;   // [rows] or NULL
  float* dw;
  float* db;
  int layer, row0, rows;
};

struct MlpArgs {
  const float* x;        // [n, dims[0]]
  float* out;            // [n, dims[L]]
  const float* d_out;    // [n, dims[L]]
  float* out2;           // optional second output: columns [split, dims[L]) go to out2 [n, dims[L]-split],
  const float* d_out2;   //   columns [0, split) to out [n, split]  (heads of an actor-critic as separate tensors)
  int split;
  float* partial;        // [grid, np_pad]  backward: per-CTA partial gradients (padded layout)
  int n, L, n_seg, act, accumulate;
  int dims[kMaxLayers + 1];
  int pd[kMaxLayers + 1];        // padded widths: pd[0] multiple of 4; pd[l>=1] in {16,32,64,128}
  int w_off[kMaxLayers];         // float offsets of layer l's padded [pd[l+1]][pd[l]] weights / [pd[l+1]] bias
  int b_off[kMaxLayers];
  int np_pad;                    // padded parameter count (weights + biases)
  int a_off[kMaxLayers + 1];     // float offsets (inside the activation arena) of layer l's INPUT rows
  MlpSeg seg[kMaxSeg];
};

__device__ __forceinline__ float act_fwd(float v, int act) {
  return act == 0 ? fmaxf(v, 0.f) : (act == 1 ? tanhf(v) : v);
}
// derivative expressed through the activation OUTPUT h
__device__ __forceinline__ float act_bwd(float h, int act) {
  return act == 0 ? (h > 0.f ? 1.f : 0.f) : (act == 1 ? 1.f - h * h : 1.f);
}

// params (global, per segment) -> padded shared-memory copy [pd[l+1]][pd[l]] + bias
__device__ void load_params(const MlpArgs& p, float* __restrict__ s_par) {
  for (int i = threadIdx.x; i < p.np_pad; i += kMlpThreads) s_par[i] = 0.f;
  __syncthreads();
  for (int s = 0; s < p.n_seg; ++s) {
    const MlpSeg& g = p.seg[s];
    const int in = p.dims[g.layer], inp = p.pd[g.layer];
    float* w = s_par + p.w_off[g.layer] + g.row0 * inp;
    for (int i = threadIdx.x; i < g.rows * in; i += kMlpThreads) w[(i / in) * inp + (i % in)] = g.w[i];
    if (g.b) {
      float* b = s_par + p.b_off[g.layer] + g.row0;
      for (int i = threadIdx.x; i < g.rows; i += kMlpThreads) b[i] = g.b[i];
    }
  }
  __syncthreads();
}

// x tile [n0 .. n0+64) x dims[0]  ->  s_in[k][n]   (zero beyond n / beyond dims[0])
__device__ void load_input_tile(const MlpArgs& p, int n0, float* __restrict__ s_in) {
  const int D = p.dims[0], Dp = p.pd[0];
  for (int i = threadIdx.x; i < Dp * kTN; i += kMlpThreads) {
    const int nn = i / Dp, k = i - nn * Dp;          // consecutive threads walk k: coalesced rows of x
    float v = 0.f;
    if (k < D && n0 + nn < p.n) v = p.x[(size_t)(n0 + nn) * D + k];
    s_in[k * kTS + nn] = v;
  }
}

// out[j][n] = act(b[j] + sum_k W[j][k] in[k][n]);  W padded [out_p][in_p], in_p % 4 == 0, out_p == 16 * JT
template <int JT>
__device__ __forceinline__ void layer_fwd(const float* __restrict__ W, const float* __restrict__ bias,
                                          const float* __restrict__ in, float* __restrict__ outp, int in_p, int act) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[JT][4];
#pragma unroll
  for (int jj = 0; jj < JT; ++jj) {
    const float b = bias[ty * JT + jj];
    acc[jj][0] = b, acc[jj][1] = b, acc[jj][2] = b, acc[jj][3] = b;
  }
  const float* wrow = W + (size_t)ty * JT * in_p;
  for (int k = 0; k < in_p; k += 4) {
    float4 a[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) a[kk] = *reinterpret_cast<const float4*>(in + (k + kk) * kTS + tx * 4);
#pragma unroll
    for (int jj = 0; jj < JT; ++jj) {
      const float4 w = *reinterpret_cast<const float4*>(wrow + jj * in_p + k);
      acc[jj][0] = fmaf(w.x, a[0].x, acc[jj][0]), acc[jj][1] = fmaf(w.x, a[0].y, acc[jj][1]);
      acc[jj][2] = fmaf(w.x, a[0].z, acc[jj][2]), acc[jj][3] = fmaf(w.x, a[0].w, acc[jj][3]);
      acc[jj][0] = fmaf(w.y, a[1].x, acc[jj][0]), acc[jj][1] = fmaf(w.y, a[1].y, acc[jj][1]);
      acc[jj][2] = fmaf(w.y, a[1].z, acc[jj][2]), acc[jj][3] = fmaf(w.y, a[1].w, acc[jj][3]);
      acc[jj][0] = fmaf(w.z, a[2].x, acc[jj][0]), acc[jj][1] = fmaf(w.z, a[2].y, acc[jj][1]);
      acc[jj][2] = fmaf(w.z, a[2].z, acc[jj][2]), acc[jj][3] = fmaf(w.z, a[2].w, acc[jj][3]);
      acc[jj][0] = fmaf(w.w, a[3].x, acc[jj][0]), acc[jj][1] = fmaf(w.w, a[3].y, acc[jj][1]);
      acc[jj][2] = fmaf(w.w, a[3].z, acc[jj][2]), acc[jj][3] = fmaf(w.w, a[3].w, acc[jj][3]);
    }
  }
#pragma unroll
  for (int jj = 0; jj < JT; ++jj) {
    float4 o;
    o.x = act_fwd(acc[jj][0], act), o.y = act_fwd(acc[jj][1], act);
    o.z = act_fwd(acc[jj][2], act), o.w = act_fwd(acc[jj][3], act);
    *reinterpret_cast<float4*>(outp + (ty * JT + jj) * kTS + tx * 4) = o;
  }
}

// Same layer for a tile of 16 samples, ONE sample per thread (tx = tid % 16 is the sample, ty = tid / 16 the output
// group): the per-step dependency chain of the fused rollout is 4x shorter and a pool of B envs spreads over B/16 CTAs.
// The k loop adds the products in exactly the order of layer_fwd, so both produce bit-identical outputs.
template <int JT>
__device__ __forceinline__ void layer_fwd_s1(const float* __restrict__ W, const float* __restrict__ bias,
                                             const float* __restrict__ in, float* __restrict__ outp, int in_p, int act) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[JT];
#pragma unroll
  for (int jj = 0; jj < JT; ++jj) acc[jj] = bias[ty * JT + jj];
  const float* wrow = W + (size_t)ty * JT * in_p;
  for (int k = 0; k < in_p; k += 4) {
    const float a0 = in[(k + 0) * kTS + tx], a1 = in[(k + 1) * kTS + tx];
    const float a2 = in[(k + 2) * kTS + tx], a3 = in[(k + 3) * kTS + tx];
#pragma unroll
    for (int jj = 0; jj < JT; ++jj) {
      const float4 w = *reinterpret_cast<const float4*>(wrow + jj * in_p + k);
      acc[jj] = fmaf(w.x, a0, acc[jj]);
      acc[jj] = fmaf(w.y, a1, acc[jj]);
      acc[jj] = fmaf(w.z, a2, acc[jj]);
      acc[jj] = fmaf(w.w, a3, acc[jj]);
    }
  }
#pragma unroll
  for (int jj = 0; jj < JT; ++jj) outp[(ty * JT + jj) * kTS + tx] = act_fwd(acc[jj], act);
}

template <int SPT>
__device__ __forceinline__ void layer_fwd_spt(const float* W, const float* bias, const float* in, float* outp, int in_p,
                                              int out_p, int act) {
  if (SPT == 4) {
    switch (out_p) {
      case 16: layer_fwd<1>(W, bias, in, outp, in_p, act); break;
      case 32: layer_fwd<2>(W, bias, in, outp, in_p, act); break;
      case 64: layer_fwd<4>(W, bias, in, outp, in_p, act); break;
      default: layer_fwd<8>(W, bias, in, outp, in_p, act); break;
    }
  } else {
    switch (out_p) {
      case 16: layer_fwd_s1<1>(W, bias, in, outp, in_p, act); break;
      case 32: layer_fwd_s1<2>(W, bias, in, outp, in_p, act); break;
      case 64: layer_fwd_s1<4>(W, bias, in, outp, in_p, act); break;
      default: layer_fwd_s1<8>(W, bias, in, outp, in_p, act); break;
    }
  }
}

__device__ __forceinline__ void layer_fwd_any(const float* W, const float* bias, const float* in, float* outp, int in_p,
                                              int out_p, int act) {
  switch (out_p) {
    case 16: layer_fwd<1>(W, bias, in, outp, in_p, act); break;
    case 32: layer_fwd<2>(W, bias, in, outp, in_p, act); break;
    case 64: layer_fwd<4>(W, bias, in, outp, in_p, act); break;
    default: layer_fwd<8>(W, bias, in, outp, in_p, act); break;
  }
}

// dprev[k][n] = act'(h[k][n]) * sum_j W[j][k] d[j][n];  in_p == 16 * KT (the previous layer's padded width)
template <int KT>
__device__ __forceinline__ void layer_bwd_delta(const float* __restrict__ W, const float* __restrict__ d,
                                                const float* __restrict__ h, float* __restrict__ dprev, int in_p,
                                                int out_p, int act) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[KT][4];
#pragma unroll
  for (int kk = 0; kk < KT; ++kk) acc[kk][0] = acc[kk][1] = acc[kk][2] = acc[kk][3] = 0.f;
  for (int j = 0; j < out_p; ++j) {
    const float4 dj = *reinterpret_cast<const float4*>(d + j * kTS + tx * 4);
    const float* wr = W + (size_t)j * in_p + ty * KT;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      const float w = wr[kk];
      acc[kk][0] = fmaf(w, dj.x, acc[kk][0]), acc[kk][1] = fmaf(w, dj.y, acc[kk][1]);
      acc[kk][2] = fmaf(w, dj.z, acc[kk][2]), acc[kk][3] = fmaf(w, dj.w, acc[kk][3]);
    }
  }
#pragma unroll
  for (int kk = 0; kk < KT; ++kk) {
    const int k = ty * KT + kk;
    const float4 hv = *reinterpret_cast<const float4*>(h + k * kTS + tx * 4);
    float4 o;
    o.x = acc[kk][0] * act_bwd(hv.x, act), o.y = acc[kk][1] * act_bwd(hv.y, act);
    o.z = acc[kk][2] * act_bwd(hv.z, act), o.w = acc[kk][3] * act_bwd(hv.w, act);
    *reinterpret_cast<float4*>(dprev + k * kTS + tx * 4) = o;
  }
}

__device__ __forceinline__ void layer_bwd_delta_any(const float* W, const float* d, const float* h, float* dprev,
                                                    int in_p, int out_p, int act) {
  switch (in_p) {
    case 16: layer_bwd_delta<1>(W, d, h, dprev, in_p, out_p, act); break;
    case 32: layer_bwd_delta<2>(W, d, h, dprev, in_p, out_p, act); break;
    case 64: layer_bwd_delta<4>(W, d, h, dprev, in_p, out_p, act); break;
    default: layer_bwd_delta<8>(W, d, h, dprev, in_p, out_p, act); break;
  }
}

// partial[w_off + j*in_p + k] += sum_n d[j][n] h[k][n];  partial[b_off + j] += sum_n d[j][n]
__device__ __forceinline__ void layer_wgrad(const float* __restrict__ d, const float* __restrict__ h,
                                            float* __restrict__ pw, float* __restrict__ pb, int in_p, int out_p,
                                            bool first_tile) {
  const int G = in_p >> 2, Gj = out_p >> 2;
  for (int bi = threadIdx.x; bi < G * Gj; bi += kMlpThreads) {
    const int kq = bi % G, jq = bi / G;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.f;
    for (int n = 0; n < kTN; n += 4) {
      float4 dv[4], hv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) dv[a] = *reinterpret_cast<const float4*>(d + (jq + a * Gj) * kTS + n);
#pragma unroll
      for (int b = 0; b < 4; ++b) hv[b] = *reinterpret_cast<const float4*>(h + (kq + b * G) * kTS + n);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          acc[a][b] = fmaf(dv[a].x, hv[b].x, acc[a][b]);
          acc[a][b] = fmaf(dv[a].y, hv[b].y, acc[a][b]);
          acc[a][b] = fmaf(dv[a].z, hv[b].z, acc[a][b]);
          acc[a][b] = fmaf(dv[a].w, hv[b].w, acc[a][b]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float* q = pw + (jq + a * Gj) * in_p + kq + b * G;
        *q = first_tile ? acc[a][b] : *q + acc[a][b];
      }
    }
  }
  for (int j = threadIdx.x; j < out_p; j += kMlpThreads) {
    float s = 0.f;
    for (int n = 0; n < kTN; n += 4) {
      const float4 v = *reinterpret_cast<const float4*>(d + j * kTS + n);
      s += (v.x + v.y) + (v.z + v.w);
    }
    pb[j] = first_tile ? s : pb[j] + s;
  }
}

// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kMlpThreads) mlp_fwd_kernel(const MlpArgs p) {
  extern __shared__ __align__(16) float smem_f[];
  float* s_par = smem_f;
  float* s_a = smem_f + ((p.np_pad + 3) & ~3);           // ping
  float* s_b = s_a + kMaxWidth * kTS;                     // pong
  load_params(p, s_par);
  const int ntiles = (p.n + kTN - 1) / kTN;
  const int O = p.dims[p.L];
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n0 = tile * kTN;
    load_input_tile(p, n0, s_a);
    __syncthreads();
    float* cur = s_a;
    float* nxt = s_b;
    for (int l = 0; l < p.L; ++l) {
      layer_fwd_any(s_par + p.w_off[l], s_par + p.b_off[l], cur, nxt, p.pd[l], p.pd[l + 1], l + 1 < p.L ? p.act : 2);
      __syncthreads();
      float* t = cur;
      cur = nxt, nxt = t;
    }
    const int S = p.out2 ? p.split : O;
    for (int i = threadIdx.x; i < kTN * O; i += kMlpThreads) {
      const int nn = i / O, o = i - nn * O;
      if (n0 + nn < p.n) {
        const float v = cur[o * kTS + nn];
        if (o < S) p.out[(size_t)(n0 + nn) * S + o] = v;
        else p.out2[(size_t)(n0 + nn) * (O - S) + (o - S)] = v;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kMlpThreads) mlp_bwd_kernel(const MlpArgs p) {
  extern __shared__ __align__(16) float smem_f[];
  float* s_par = smem_f;
  float* s_act = smem_f + ((p.np_pad + 3) & ~3);          // inputs of every layer: a_off[l]
  float* s_d0 = s_act + p.a_off[p.L];                     // delta ping
  float* s_d1 = s_d0 + kMaxWidth * kTS;                   // delta pong
  load_params(p, s_par);
  const int ntiles = (p.n + kTN - 1) / kTN;
  const int O = p.dims[p.L], Op = p.pd[p.L];
  float* part = p.partial + (size_t)blockIdx.x * p.np_pad;
  bool first = true;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n0 = tile * kTN;
    load_input_tile(p, n0, s_act + p.a_off[0]);
    // d_out tile -> s_d0[o][n]   (zero beyond n / beyond O)
    const int S = p.d_out2 ? p.split : O;
    for (int i = threadIdx.x; i < Op * kTN; i += kMlpThreads) {
      const int nn = i / Op, o = i - nn * Op;
      float v = 0.f;
      if (o < O && n0 + nn < p.n)
        v = o < S ? p.d_out[(size_t)(n0 + nn) * S + o] : p.d_out2[(size_t)(n0 + nn) * (O - S) + (o - S)];
      s_d0[o * kTS + nn] = v;
    }
    __syncthreads();
    // recompute the hidden activations (the output layer itself is not needed)
    for (int l = 0; l + 1 < p.L; ++l) {
      layer_fwd_any(s_par + p.w_off[l], s_par + p.b_off[l], s_act + p.a_off[l], s_act + p.a_off[l + 1], p.pd[l],
                    p.pd[l + 1], p.act);
      __syncthreads();
    }
    float* dcur = s_d0;
    float* dnxt = s_d1;
    for (int l = p.L - 1; l >= 0; --l) {
      layer_wgrad(dcur, s_act + p.a_off[l], part + p.w_off[l], part + p.b_off[l], p.pd[l], p.pd[l + 1], first);
      if (l > 0) {
        layer_bwd_delta_any(s_par + p.w_off[l], dcur, s_act + p.a_off[l], dnxt, p.pd[l], p.pd[l + 1], p.act);
        __syncthreads();
        float* t = dcur;
        dcur = dnxt, dnxt = t;
      }
    }
    first = false;
    __syncthreads();
  }
  if (first) {            // a CTA without tiles still owns a partial row: zero it
    for (int i = threadIdx.x; i < p.np_pad; i += kMlpThreads) part[i] = 0.f;
  }
}

// grads (true layout, per segment) (+)= sum over CTAs of the padded partials.  One WARP per parameter: lane l adds
// parts l, l+32, ... (coalescing does not matter at 4 B per part; latency does), then a fixed xor-shuffle tree — the
// order never changes between runs (deterministic).
__global__ void __launch_bounds__(256) mlp_grad_reduce_kernel(const MlpArgs p, int nparts) {
  const int lane = threadIdx.x & 31;
  const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int s = 0; s < p.n_seg; ++s) {
    const MlpSeg& g = p.seg[s];
    const int in = p.dims[g.layer], inp = p.pd[g.layer];
    const int nw = g.rows * in;
    for (int i = wglobal; i < nw + g.rows; i += nwarps) {
      int off;
      float* dst;
      if (i < nw) {
        off = p.w_off[g.layer] + (g.row0 + i / in) * inp + (i % in);
        dst = g.dw + i;
      } else {
        off = p.b_off[g.layer] + g.row0 + (i - nw);
        dst = g.db ? g.db + (i - nw) : nullptr;
      }
      if (!dst) continue;
      float a = 0.f;
      for (int c = lane; c < nparts; c += 32) a += p.partial[(size_t)c * p.np_pad + off];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) *dst = p.accumulate ? *dst + a : a;
    }
  }
}

// ---------------------------------------------------------------------------
// Fused on-device actor pool for MLP policies: ONE launch runs all T lock-step steps of a tile of 64 envs —
// policy/value forward (weights resident in shared memory), action sampling, env step, episode bookkeeping — and
// writes the trajectory straight into the time-major (T,B) rollout buffers.  Replaces, per step and per env, the
// reference's agent.sample -> env.step -> rollout.append round trip (benchmark/torch/ppo/train.py:91-101,
// benchmark/torch/a2c/actor.py:56-80) and its xparl RPC (parl/remote/remote_wrapper.py:178-227).
// Arithmetic is the same device code as the stand-alone kernels (rl_mlp_fwd, rl_sample_*, rl_env_*_step), so a
// fused rollout is bit-identical to stepping those kernels one at a time (tests/test_gpu_rollout.py).
// ---------------------------------------------------------------------------
struct RolloutArgs {
  int env_kind;        // 0 MuJoCo-shaped synthetic (obs ~ N(0,1)^D), 1 CartPole physics
  int policy_kind;     // 0 categorical over AD actions, 1 diagonal Gaussian with AD dimensions
  int T, B, AD, has_value, max_steps;
  const float* logstd; // [AD] (Gaussian)
  float* obs_cur;      // [B, D]  observation each env is in (in: before step 0, out: after step T-1)
  EpisodeStats st;
  uint32_t k0, k1, step0, env_offset, done_thr;
  float* obs_out;      // [T, B, D]
  void* act_out;       // [T, B] int32 (categorical) or [T, B, AD] float32 (Gaussian)
  float* logp_out;     // [T, B] or NULL
  float* val_out;      // [T+1, B] (row T: value of the observation after the last step) or NULL
  float* logits_out;   // [T, B, AD] or NULL (categorical: behaviour logits for off-policy corrections)
  float* rew_out;      // [T, B]
  uint8_t* done_out;   // [T, B]
  int use_vn;          // VecNormalizeEnv between the env and the policy (per-env running statistics)
  VecNormState vn;
};

// sample_categorical_exact (philox.cuh) over a strided row — same operations in the same order
__device__ __forceinline__ int sample_categorical_exact_strided(const float* __restrict__ lg, int stride, int A, float u) {
  float m = lg[0];
  for (int j = 1; j < A; ++j) m = fmaxf(m, lg[j * stride]);
  float total = 0.f;
  for (int j = 0; j < A; ++j) total = __fadd_rn(total, exp_exact(__fsub_rn(lg[j * stride], m)));
  const float thr = __fmul_rn(u, total);
  float acc = 0.f;
  int a = 0;
  for (int j = 0; j < A; ++j) {
    acc = __fadd_rn(acc, exp_exact(__fsub_rn(lg[j * stride], m)));
    a += (acc <= thr) ? 1 : 0;
  }
  return min(a, A - 1);
}

template <int SPT>
__global__ void __launch_bounds__(kMlpThreads) rollout_mlp_kernel(const MlpArgs p, const RolloutArgs r) {
  constexpr int TNR = 16 * SPT;                            // envs per CTA (64: throughput tile, 16: latency tile)
  extern __shared__ __align__(16) float smem_f[];
  float* s_par = smem_f;
  float* s_x = smem_f + ((p.np_pad + 3) & ~3);            // observation tile [pd[0]][kTS]
  float* s_a = s_x + kMaxWidth * kTS;                      // ping
  float* s_b = s_a + kMaxWidth * kTS;                      // pong
  __shared__ double s_cnt[TNR];                            // VecNormalize: observation count before this step
  load_params(p, s_par);
  const int n0 = blockIdx.x * TNR;
  const int D = p.dims[0], Dp = p.pd[0], B = r.B;
  const int tid = threadIdx.x;
  // the envs' current observations -> s_x[k][n]
  for (int i = tid; i < Dp * TNR; i += kMlpThreads) {
    const int nn = i / Dp, k = i - nn * Dp;
    float v = 0.f;
    if (k < D && n0 + nn < B) v = r.obs_cur[(size_t)(n0 + nn) * D + k];
    s_x[k * kTS + nn] = v;
  }
  __syncthreads();
  for (int t = 0; t <= r.T; ++t) {
    if (t == r.T && !(r.has_value && r.val_out)) break;
    if (t < r.T) {
      for (int i = tid; i < D * TNR; i += kMlpThreads) {     // trajectory: observation of step t
        const int nn = i / D, k = i - nn * D;
        if (n0 + nn < B) r.obs_out[((size_t)t * B + n0 + nn) * D + k] = s_x[k * kTS + nn];
      }
    }
    const float* cur = s_x;
    float* nxt = s_a;
    for (int l = 0; l < p.L; ++l) {
      layer_fwd_spt<SPT>(s_par + p.w_off[l], s_par + p.b_off[l], cur, nxt, p.pd[l], p.pd[l + 1], l + 1 < p.L ? p.act : 2);
      __syncthreads();
      cur = nxt;
      nxt = (nxt == s_a) ? s_b : s_a;
    }
    // one thread per env: value, action, env step (whole warps: the episode bookkeeping is warp-synchronous)
    if (tid < (TNR < 32 ? 32 : TNR)) {
      const int b = n0 + tid;
      const bool valid = tid < TNR && b < B;
      const uint32_t env = r.env_offset + (uint32_t)b;
      const uint32_t step = r.step0 + (uint32_t)t;
      const float* o = cur + tid;                            // output j of this env: o[j * kTS]
      if (valid && r.has_value && r.val_out) r.val_out[(size_t)t * B + b] = o[r.AD * kTS];
      float reward = 0.f;
      bool done = false;
      if (t < r.T) {
        int a_cat = 0;
        if (valid) {
          const size_t tb = (size_t)t * B + b;
          if (r.policy_kind == 0) {
            const uint4 ua = philox4x32_10(env, step, 0u, STREAM_ACTION, r.k0, r.k1);
            a_cat = sample_categorical_exact_strided(o, kTS, r.AD, u01_24(ua.x));
            reinterpret_cast<int*>(r.act_out)[tb] = a_cat;
            if (r.logp_out) {
              float m = o[0];
              for (int j = 1; j < r.AD; ++j) m = fmaxf(m, o[j * kTS]);
              float S = 0.f;
              for (int j = 0; j < r.AD; ++j) S += expf(o[j * kTS] - m);
              r.logp_out[tb] = o[a_cat * kTS] - m - logf(S);
            }
            if (r.logits_out)
              for (int j = 0; j < r.AD; ++j) r.logits_out[tb * r.AD + j] = o[j * kTS];
          } else {
            float lp = 0.f;
            float* ao = reinterpret_cast<float*>(r.act_out) + tb * r.AD;
            for (int blk = 0; blk * 4 < r.AD; ++blk) {
              float z[4];
              gauss_block(env, step, (uint32_t)blk, STREAM_GAUSS, r.k0, r.k1, z);
              for (int k = 0; k < 4 && blk * 4 + k < r.AD; ++k) {
                const int d = blk * 4 + k;
                const float ls = r.logstd[d], sd = expf(ls);
                ao[d] = fmaf(sd, z[k], o[d * kTS]);
                lp += -0.5f * z[k] * z[k] - ls - 0.9189385332046727f;
              }
            }
            if (r.logp_out) r.logp_out[tb] = lp;
          }
          // ---- env step (same draws / physics as rl_env_mujoco_synth_step / rl_env_cartpole_step); the MuJoCo-shaped
          //      env's next observation is produced by ALL threads of the CTA after this block (stage 2)
          if (r.env_kind == 0) {
            const uint4 x = philox4x32_10(env, step, 0u, STREAM_REWDONE, r.k0, r.k1);
            reward = (float)(x.x & 1u);
            done = x.y < r.done_thr;
            if (r.max_steps > 0 && r.st.ep_len[b] + 1 >= r.max_steps) done = true;
          } else {
            float4 s = make_float4(s_x[tid], s_x[kTS + tid], s_x[2 * kTS + tid], s_x[3 * kTS + tid]);
            done = cartpole_physics(s, a_cat);
            reward = 1.0f;
            if (r.max_steps > 0 && r.st.ep_len[b] + 1 >= r.max_steps) done = true;
            if (done) s = cartpole_reset_state(env, step + 1u, r.k0, r.k1);
            s_x[tid] = s.x, s_x[kTS + tid] = s.y, s_x[2 * kTS + tid] = s.z, s_x[3 * kTS + tid] = s.w;
          }
          float rew_seen = reward;           // what the agent sees; the episode statistics keep the raw reward
          if (r.use_vn) {                    // VecNormalizeEnv.step (reward side) of env b
            rew_seen = vecnorm_reward(r.vn, b, reward, done);
            if (r.env_kind == 0) s_cnt[tid] = r.vn.ob_count[b];
            else vecnorm_obs_filter(r.vn, b, D, s_x + tid, kTS);
          }
          r.rew_out[tb] = rew_seen;
          r.done_out[tb] = done ? 1 : 0;
        }
        episode_update(r.st, b, valid, reward, done);
      }
    }
    __syncthreads();
    // ---- stage 2 (MuJoCo-shaped env): next observation of every env, one Box-Muller block (4 components) per
    //      task over all 256 threads; with VecNormalize each component is filtered where it is produced
    if (t < r.T && r.env_kind == 0) {
      const int nblk = (D + 3) >> 2;
      const uint32_t step = r.step0 + (uint32_t)t;
      const bool filt = r.use_vn && r.vn.norm_ob;
      for (int task = tid; task < TNR * nblk; task += kMlpThreads) {
        const int blk = task / TNR, n = task - blk * TNR;           // consecutive threads = consecutive envs
        const int b = n0 + n;
        if (b >= B) continue;
        float z[4];
        gauss_block(r.env_offset + (uint32_t)b, step + 1u, (uint32_t)blk, STREAM_OBS, r.k0, r.k1, z);
        const double cnt = filt ? s_cnt[n] : 0.0;
        for (int k = 0; k < 4 && blk * 4 + k < D; ++k) {
          const int d = blk * 4 + k;
          s_x[d * kTS + n] = filt ? vecnorm_obs_dim(r.vn, b, D, d, cnt, z[k]) : z[k];
        }
      }
      if (filt && r.vn.update && tid < TNR && n0 + tid < B) r.vn.ob_count[n0 + tid] = s_cnt[tid] + 1.0;
      __syncthreads();
    }
  }
  // carry the envs' observations to the next rollout
  for (int i = tid; i < D * TNR; i += kMlpThreads) {
    const int nn = i / D, k = i - nn * D;
    if (n0 + nn < B) r.obs_cur[(size_t)(n0 + nn) * D + k] = s_x[k * kTS + nn];
  }
}

static int pad_width(int d) { return d <= 16 ? 16 : d <= 32 ? 32 : d <= 64 ? 64 : 128; }

// Fills the shape part of MlpArgs; returns RL_OK or an error code with the message set.
static int mlp_shape(MlpArgs& a, int n, int n_layers, const int* dims, int n_seg, const int* seg_layer,
                     const int* seg_rows, const float* const* seg_w, const float* const* seg_b, int act) {
  RL_CHECK_ARG(n >= 1 && n_layers >= 1 && n_layers <= kMaxLayers, "mlp: n=%d layers=%d (1..%d)", n, n_layers, kMaxLayers);
  RL_CHECK_ARG(n_seg >= n_layers && n_seg <= kMaxSeg, "mlp: %d parameter segments (need %d..%d)", n_seg, n_layers, kMaxSeg);
  RL_CHECK_ARG(act >= 0 && act <= 2, "mlp: act %d not in {0 relu, 1 tanh, 2 none}", act);
  a.n = n, a.L = n_layers, a.n_seg = n_seg, a.act = act;
  for (int l = 0; l <= n_layers; ++l) {
    RL_CHECK_ARG(dims[l] >= 1 && dims[l] <= kMaxWidth, "mlp: width %d of layer %d outside 1..%d", dims[l], l, kMaxWidth);
    a.dims[l] = dims[l];
    a.pd[l] = l == 0 ? (dims[0] + 3) & ~3 : pad_width(dims[l]);
  }
  int off = 0, aoff = 0;
  for (int l = 0; l < n_layers; ++l) {
    a.w_off[l] = off;
    off += a.pd[l + 1] * a.pd[l];
    a.b_off[l] = off;
    off += a.pd[l + 1];
    a.a_off[l] = aoff;
    aoff += a.pd[l] * kTS;
  }
  a.a_off[n_layers] = aoff;
  a.np_pad = off;
  int rows_seen[kMaxLayers] = {0, 0, 0, 0};
  for (int s = 0; s < n_seg; ++s) {
    const int l = seg_layer[s];
    RL_CHECK_ARG(l >= 0 && l < n_layers && seg_rows[s] >= 1 && seg_w[s], "mlp: bad segment %d", s);
    a.seg[s].w = seg_w[s], a.seg[s].b = seg_b ? seg_b[s] : nullptr;
    a.seg[s].dw = nullptr, a.seg[s].db = nullptr;
    a.seg[s].layer = l, a.seg[s].row0 = rows_seen[l], a.seg[s].rows = seg_rows[s];
    rows_seen[l] += seg_rows[s];
  }
  for (int l = 0; l < n_layers; ++l)
    RL_CHECK_ARG(rows_seen[l] == dims[l + 1], "mlp: segments of layer %d cover %d of %d rows", l, rows_seen[l], dims[l + 1]);
  return RL_OK;
}

}  // namespace rl

extern "C" size_t rl_mlp_workspace_bytes(int n_layers, const int* dims) {
  using namespace rl;
  if (n_layers < 1 || n_layers > kMaxLayers) return 0;
  size_t np = 0;
  for (int l = 0; l < n_layers; ++l) {
    const int inp = l == 0 ? (dims[0] + 3) & ~3 : pad_width(dims[l]);
    np += (size_t)pad_width(dims[l + 1]) * (inp + 1);
  }
  return np * sizeof(float) * 304 + 256;          // up to 2 CTAs on each of 152 SMs
}

extern "C" int rl_mlp_fwd(const float* x, int n, int n_layers, const int* dims, int n_seg, const int* seg_layer,
                          const int* seg_rows, const float* const* seg_w, const float* const* seg_b, int act, float* out,
                          float* out2, int split, rl_stream_t stream) {
  using namespace rl;
  RL_CHECK_ARG(x && out && dims && seg_layer && seg_rows && seg_w, "mlp_fwd: null pointer");
  MlpArgs a;
  const int rc = mlp_shape(a, n, n_layers, dims, n_seg, seg_layer, seg_rows, seg_w, seg_b, act);
  if (rc != RL_OK) return rc;
  RL_CHECK_ARG(!out2 || (split >= 1 && split < dims[n_layers]), "mlp_fwd: split %d outside 1..%d", split, dims[n_layers] - 1);
  a.x = x, a.out = out, a.d_out = nullptr, a.partial = nullptr, a.accumulate = 0;
  a.out2 = out2, a.d_out2 = nullptr, a.split = split;
  const size_t smem = (((size_t)a.np_pad + 3) & ~(size_t)3) * 4 + 2 * (size_t)kMaxWidth * kTS * 4;
  RL_CHECK_ARG(smem <= 220 * 1024, "mlp_fwd: network too large for shared memory (%zu B)", smem);
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = (n + kTN - 1) / kTN;
  const int grid = ntiles < 2 * nsm ? ntiles : 2 * nsm;
  RL_SMEM_OPTIN(mlp_fwd_kernel);
  mlp_fwd_kernel<<<grid, kMlpThreads, smem, (cudaStream_t)stream>>>(a);
  RL_CHECK_LAUNCH("rl_mlp_fwd");
  return RL_OK;
}

extern "C" int rl_mlp_bwd(const float* x, int n, int n_layers, const int* dims, int n_seg, const int* seg_layer,
                          const int* seg_rows, const float* const* seg_w, const float* const* seg_b, int act,
                          const float* d_out, const float* d_out2, int split, float* const* seg_dw,
                          float* const* seg_db, int accumulate, void* workspace, size_t workspace_bytes,
                          rl_stream_t stream) {
  using namespace rl;
  RL_CHECK_ARG(x && d_out && dims && seg_layer && seg_rows && seg_w && seg_dw && workspace, "mlp_bwd: null pointer");
  MlpArgs a;
  const int rc = mlp_shape(a, n, n_layers, dims, n_seg, seg_layer, seg_rows, seg_w, seg_b, act);
  if (rc != RL_OK) return rc;
  for (int s = 0; s < n_seg; ++s) {
    RL_CHECK_ARG(seg_dw[s], "mlp_bwd: null weight-gradient pointer for segment %d", s);
    a.seg[s].dw = seg_dw[s];
    a.seg[s].db = seg_db ? seg_db[s] : nullptr;
  }
  RL_CHECK_ARG(!d_out2 || (split >= 1 && split < dims[n_layers]), "mlp_bwd: split %d outside 1..%d", split, dims[n_layers] - 1);
  a.x = x, a.out = nullptr, a.d_out = d_out, a.accumulate = accumulate;
  a.out2 = nullptr, a.d_out2 = d_out2, a.split = split;
  const size_t smem = (((size_t)a.np_pad + 3) & ~(size_t)3) * 4 + (size_t)a.a_off[a.L] * 4 + 2 * (size_t)kMaxWidth * kTS * 4;
  RL_CHECK_ARG(smem <= 220 * 1024, "mlp_bwd: network too large for shared memory (%zu B)", smem);
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = (n + kTN - 1) / kTN;
  const int per_sm = smem <= 110 * 1024 ? 2 : 1;
  int grid = ntiles < per_sm * nsm ? ntiles : per_sm * nsm;
  const size_t need = (size_t)grid * a.np_pad * sizeof(float);
  if (workspace_bytes < need) {
    set_error("mlp_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return RL_ERR_WORKSPACE;
  }
  a.partial = reinterpret_cast<float*>(workspace);
  RL_SMEM_OPTIN(mlp_bwd_kernel);
  mlp_bwd_kernel<<<grid, kMlpThreads, smem, (cudaStream_t)stream>>>(a);
  RL_CHECK_LAUNCH("rl_mlp_bwd");
  mlp_grad_reduce_kernel<<<(a.np_pad + 7) / 8, 256, 0, (cudaStream_t)stream>>>(a, grid);
  RL_CHECK_LAUNCH("rl_mlp_bwd(reduce)");
  return RL_OK;
}

extern "C" int rl_rollout_mlp(int n_layers, const int* dims, int n_seg, const int* seg_layer, const int* seg_rows,
                              const float* const* seg_w, const float* const* seg_b, int act, int env_kind,
                              int policy_kind, int T, int B, int action_dim, int has_value, const float* logstd,
                              float* obs_cur, float* ep_ret, int32_t* ep_len, float* totals, float* ring_ret,
                              int32_t* ring_len, uint32_t* ring_head, int ring_cap, uint64_t seed, uint32_t step0,
                              uint32_t env_offset, float p_done, int max_episode_steps, float* obs_out, void* act_out,
                              float* logp_out, float* val_out, float* logits_out, float* rew_out, uint8_t* done_out,
                              double* const* vecnorm_state, const double* vecnorm_cfg, int vecnorm_flags,
                              rl_stream_t stream) {
  using namespace rl;
  RL_CHECK_ARG(dims && seg_layer && seg_rows && seg_w && obs_cur && ep_ret && ep_len && totals && obs_out && act_out &&
                   rew_out && done_out,
               "rollout_mlp: null pointer");
  RL_CHECK_ARG(env_kind == 0 || env_kind == 1, "rollout_mlp: env_kind %d not in {0 mujoco-synth, 1 cartpole}", env_kind);
  RL_CHECK_ARG(policy_kind == 0 || policy_kind == 1, "rollout_mlp: policy_kind %d not in {0 categorical, 1 gaussian}", policy_kind);
  RL_CHECK_ARG(T >= 1 && B >= 1 && action_dim >= 1, "rollout_mlp: T=%d B=%d action_dim=%d", T, B, action_dim);
  MlpArgs a;
  const int rc = mlp_shape(a, B, n_layers, dims, n_seg, seg_layer, seg_rows, seg_w, seg_b, act);
  if (rc != RL_OK) return rc;
  RL_CHECK_ARG(dims[n_layers] == action_dim + (has_value ? 1 : 0),
               "rollout_mlp: network has %d outputs, expected action_dim %d + value %d", dims[n_layers], action_dim,
               has_value ? 1 : 0);
  RL_CHECK_ARG(env_kind != 1 || (dims[0] == 4 && policy_kind == 0 && action_dim == 2),
               "rollout_mlp: CartPole needs obs dim 4 and a 2-way categorical policy");
  RL_CHECK_ARG(policy_kind != 1 || logstd, "rollout_mlp: Gaussian policy without logstd");
  a.x = nullptr, a.out = nullptr, a.d_out = nullptr, a.out2 = nullptr, a.d_out2 = nullptr, a.split = 0;
  a.partial = nullptr, a.accumulate = 0;
  RolloutArgs r;
  r.env_kind = env_kind, r.policy_kind = policy_kind, r.T = T, r.B = B, r.AD = action_dim, r.has_value = has_value;
  r.max_steps = max_episode_steps, r.logstd = logstd, r.obs_cur = obs_cur;
  r.st = make_episode_stats(ep_ret, ep_len, totals, ring_ret, ring_len, ring_head, ring_cap);
  r.k0 = (uint32_t)seed, r.k1 = (uint32_t)(seed >> 32), r.step0 = step0, r.env_offset = env_offset;
  r.done_thr = prob_threshold(p_done);
  r.obs_out = obs_out, r.act_out = act_out, r.logp_out = logp_out, r.val_out = val_out, r.logits_out = logits_out;
  r.rew_out = rew_out, r.done_out = done_out;
  r.use_vn = vecnorm_state != nullptr;
  memset(&r.vn, 0, sizeof(r.vn));
  if (r.use_vn) {
    RL_CHECK_ARG(vecnorm_cfg, "rollout_mlp: vecnorm_state without vecnorm_cfg");
    for (int i = 0; i < 7; ++i) RL_CHECK_ARG(vecnorm_state[i], "rollout_mlp: vecnorm_state[%d] is NULL", i);
    r.vn.ob_mean = vecnorm_state[0], r.vn.ob_var = vecnorm_state[1], r.vn.ob_count = vecnorm_state[2];
    r.vn.ret = vecnorm_state[3], r.vn.ret_mean = vecnorm_state[4], r.vn.ret_var = vecnorm_state[5];
    r.vn.ret_count = vecnorm_state[6];
    r.vn.clipob = vecnorm_cfg[0], r.vn.cliprew = vecnorm_cfg[1], r.vn.gamma = vecnorm_cfg[2], r.vn.eps = vecnorm_cfg[3];
    r.vn.update = vecnorm_flags & 1, r.vn.norm_ob = (vecnorm_flags >> 1) & 1, r.vn.norm_ret = (vecnorm_flags >> 2) & 1;
  }
  const size_t smem = (((size_t)a.np_pad + 3) & ~(size_t)3) * 4 + 3 * (size_t)kMaxWidth * kTS * 4;
  RL_CHECK_ARG(smem <= 220 * 1024, "rollout_mlp: network too large for shared memory (%zu B)", smem);
  // tile choice: 64 envs per CTA when that alone fills the GPU, else 16 (4x shorter per-step chain, 4x the CTAs)
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  if ((B + kTN - 1) / kTN >= nsm) {
    RL_SMEM_OPTIN(rollout_mlp_kernel<4>);
    rollout_mlp_kernel<4><<<(B + kTN - 1) / kTN, kMlpThreads, smem, (cudaStream_t)stream>>>(a, r);
  } else {
    RL_SMEM_OPTIN(rollout_mlp_kernel<1>);
    rollout_mlp_kernel<1><<<(B + 15) / 16, kMlpThreads, smem, (cudaStream_t)stream>>>(a, r);
  }
  RL_CHECK_LAUNCH("rl_rollout_mlp");
  return RL_OK;
}
