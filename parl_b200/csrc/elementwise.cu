// Small bandwidth-bound glue kernels of the learner network (bf16, 16-byte vector accesses).
#include <cuda_bf16.h>

#include "common.cuh"

namespace rl {

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// x[m, n] = act(x[m, n] + bias[n]) in place; N % 8 == 0
__global__ void __launch_bounds__(256) bias_act_bf16_kernel(uint4* __restrict__ x, const float* __restrict__ bias,
                                                            long long total_vec, int nvec_per_row, int relu) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += stride) {
    const int c = (int)(i % nvec_per_row) * 8;
    uint4 q = x[i];
    uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a = bf_lo(w[k]) + __ldg(bias + c + 2 * k), b = bf_hi(w[k]) + __ldg(bias + c + 2 * k + 1);
      if (relu) a = fmaxf(a, 0.f), b = fmaxf(b, 0.f);
      w[k] = pack_bf2(a, b);
    }
    x[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ReLU backward + re-gridding: dst[n, y, x, :] = src[n, y, x, :] * (act[n, y, x, :] > 0) for y < PH, x < PW,
// src/act compact [N, PH, PW, C], dst on a larger grid [N, GH, GW, C] (cells outside PHxPW are left untouched).
__global__ void __launch_bounds__(256) mask_scatter_grid_bf16_kernel(const uint4* __restrict__ src,
                                                                     const uint4* __restrict__ act,
                                                                     uint4* __restrict__ dst, long long total_vec,
                                                                     int PH, int PW, int GH, int GW, int cvec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += stride) {
    const int cv = (int)(i % cvec);
    long long p = i / cvec;
    const int xx = (int)(p % PW);
    p /= PW;
    const int yy = (int)(p % PH);
    const long long n = p / PH;
    const uint4 g = __ldcs(src + i), a = __ldcs(act + i);
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, aw[4] = {a.x, a.y, a.z, a.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float lo = bf_lo(aw[k]) > 0.f ? bf_lo(gw[k]) : 0.f, hi = bf_hi(aw[k]) > 0.f ? bf_hi(gw[k]) : 0.f;
      o[k] = pack_bf2(lo, hi);
    }
    dst[((n * GH + yy) * GW + xx) * cvec + cv] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace rl

using namespace rl;

extern "C" int rl_bias_act_bf16(void* x, const float* bias, long long M, int N, int relu, rl_stream_t stream) {
  RL_CHECK_ARG(x && bias && M > 0 && N > 0 && N % 8 == 0 && aligned16(x), "bias_act_bf16: bad argument");
  const long long total = M * (N / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  bias_act_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((uint4*)x, bias, total, N / 8, relu);
  RL_CHECK_LAUNCH("rl_bias_act_bf16");
  return RL_OK;
}

extern "C" int rl_mask_scatter_grid_bf16(const void* src, const void* act, void* dst, long long N, int PH, int PW, int GH,
                                         int GW, int C, rl_stream_t stream) {
  RL_CHECK_ARG(src && act && dst && N > 0 && C % 8 == 0 && PH <= GH && PW <= GW, "mask_scatter_grid_bf16: bad argument");
  RL_CHECK_ARG(aligned16(src) && aligned16(act) && aligned16(dst), "mask_scatter_grid_bf16: alignment");
  const long long total = N * PH * PW * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  mask_scatter_grid_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)src, (const uint4*)act, (uint4*)dst, total, PH, PW, GH, GW, C / 8);
  RL_CHECK_LAUNCH("rl_mask_scatter_grid_bf16");
  return RL_OK;
}
