// Learner update: global-norm gradient clipping + Adam on ONE flat parameter buffer, 2 launches,
// no host synchronisation (the clip scale stays on the device).
//
// Reference behaviour being replaced (PaddlePaddle/PARL learner steps):
//   parl/algorithms/paddle/impala/impala.py:113-117,210-213  Adam + ClipGradByGlobalNorm(40):
//        scale = clip / max(||g||, clip)
//   parl/algorithms/torch/a2c.py:65-69, ppo.py:144-147        clip_grad_norm_(max_norm) then Adam:
//        scale = min(1, max_norm / (||g|| + 1e-6))
//   parl/algorithms/torch/dqn.py:68-71                        plain Adam
// Adam (torch.optim.Adam / paddle.optimizer.Adam, identical algebra):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// Bytes per parameter: read 16 (p,g,m,v) + write 12 (p,m,v); norm pass reads 4.
#include <cuda_bf16.h>

#include "common.cuh"
#include "reduce.cuh"

namespace rl {

constexpr int kOT = 256;

__global__ void __launch_bounds__(kOT) grad_sq_norm_kernel(const float* __restrict__ g, long long n,
                                                           float* __restrict__ out_norm, float* partials,
                                                           unsigned* ticket) {
  float acc = 0.f;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * kOT;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = (long long)blockIdx.x * kOT + threadIdx.x; i < n4; i += stride) {
    const float4 x = g4[i];
    acc += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float x = g[(n4 << 2) + threadIdx.x];
    acc += x * x;
  }
  const float sums[1] = {acc};
  double tot[1];
  if (grid_reduce<1, kOT>(sums, partials, ticket, tot)) out_norm[0] = (float)sqrt(tot[0]);
}

// clip_mode: 0 none, 1 torch clip_grad_norm_ (max_norm/(norm+1e-6), clamped to 1), 2 paddle ClipGradByGlobalNorm
__global__ void __launch_bounds__(kOT) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, long long n,
                                                        const float* __restrict__ lr_ptr, float lr, float b1, float b2,
                                                        float eps, float bc1, float bc2_sqrt, float grad_div,
                                                        const float* __restrict__ norm, float max_norm, int clip_mode,
                                                        int zero_grad, float* __restrict__ g_mut,
                                                        const int* __restrict__ step_ptr) {
  if (step_ptr) {                      // update count resident on the device (CUDA-graph replay): same fp64 formula
    const double st = (double)step_ptr[0];
    bc1 = (float)(1.0 - pow((double)b1, st));
    bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, st));
  }
  float scale = 1.0f / grad_div;
  if (clip_mode != 0) {
    const float nrm = norm[0] / grad_div;
    if (clip_mode == 1) scale *= fminf(1.0f, max_norm / (nrm + 1e-6f));
    else scale *= max_norm / fmaxf(nrm, max_norm);
  }
  if (lr_ptr) lr = lr_ptr[0];
  const float step = lr / bc1;
  const long long stride = (long long)gridDim.x * kOT;
  for (long long i = (long long)blockIdx.x * kOT + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * scale;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step * mi / (sqrtf(vi) / bc2_sqrt + eps);
    if (zero_grad) g_mut[i] = 0.f;
  }
}

}  // namespace rl

using namespace rl;

// Operand refresh after a learner update: every bf16 / fp32 operand copy the network kernels keep (KRSC conv
// filters, their space-to-depth and transposed forms, (H,W,C)-ordered fc columns, bias vectors) is an index
// permutation of the flat fp32 master buffer, so ONE gather launch rebuilds them all: out[i] = cast(src[idx[i]]),
// idx < 0 -> 0 (padding, unused head columns).  Replaces ~26 permute+copy launches per update.
__global__ void __launch_bounds__(256) gather_cast_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                                          long long n, void* __restrict__ out, int out_bf16) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (out_bf16) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i * 8 < n; i += stride) {
      const long long e = i * 8;
      float v[8];
      if (e + 8 <= n) {
        const int4 a = __ldg(reinterpret_cast<const int4*>(idx + e)), b = __ldg(reinterpret_cast<const int4*>(idx + e) + 1);
        const int j[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = j[k] >= 0 ? __ldg(src + j[k]) : 0.f;
        uint32_t pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
          pk[k] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + e) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      } else {
        for (long long t = e; t < n; ++t) {
          const int jj = idx[t];
          reinterpret_cast<__nv_bfloat16*>(out)[t] = __float2bfloat16(jj >= 0 ? src[jj] : 0.f);
        }
      }
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const int jj = idx[i];
      reinterpret_cast<float*>(out)[i] = jj >= 0 ? __ldg(src + jj) : 0.f;
    }
  }
}

extern "C" int rl_gather_cast(const float* src, const int32_t* idx, long long n, void* out, int out_bf16,
                              rl_stream_t stream) {
  RL_CHECK_ARG(src && idx && out && n > 0, "gather_cast: bad argument");
  RL_CHECK_ARG(aligned16(idx) && aligned16(out), "gather_cast: idx and out must be 16-byte aligned");
  const long long work = out_bf16 ? (n + 7) / 8 : n;
  long long blocks = (work + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gather_cast_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, idx, n, out, out_bf16);
  RL_CHECK_LAUNCH("rl_gather_cast");
  return RL_OK;
}

extern "C" int rl_grad_global_norm(const float* grad, long long n, float* out_norm, void* workspace,
                                   size_t workspace_bytes, rl_stream_t stream) {
  RL_CHECK_ARG(grad && out_norm && workspace && n > 0 && aligned16(grad), "grad_global_norm: bad argument");
  long long blocks = ((n >> 2) + kOT - 1) / kOT;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  if (workspace_bytes < 256 + (size_t)blocks * sizeof(float)) {
    set_error("grad_global_norm: workspace too small");
    return RL_ERR_WORKSPACE;
  }
  grad_sq_norm_kernel<<<(unsigned)blocks, kOT, 0, (cudaStream_t)stream>>>(
      grad, n, out_norm, reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256),
      reinterpret_cast<unsigned*>(workspace));
  RL_CHECK_LAUNCH("rl_grad_global_norm");
  return RL_OK;
}

extern "C" int rl_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                            const float* lr_device, float lr, float beta1, float beta2, float eps, int step,
                            float grad_div, const float* grad_norm, float max_norm, int clip_mode, int zero_grad,
                            const int32_t* step_device, rl_stream_t stream) {
  RL_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && (step >= 1 || step_device), "adam_step: bad argument");
  if (step < 1) step = 1;
  RL_CHECK_ARG(clip_mode == 0 || grad_norm, "adam_step: clip_mode %d needs grad_norm", clip_mode);
  RL_CHECK_ARG(grad_div > 0.f, "adam_step: grad_div must be positive");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  long long blocks = (n + kOT - 1) / kOT;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_step_kernel<<<(unsigned)blocks, kOT, 0, (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, n, lr_device, lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), grad_div,
      grad_norm, max_norm, clip_mode, zero_grad, grad, step_device);
  RL_CHECK_LAUNCH("rl_adam_step");
  return RL_OK;
}
