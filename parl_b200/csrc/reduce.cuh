// Deterministic single-launch grid reduction: warp shuffle -> CTA partial ->
// (last CTA by ticket) fixed-order fp64 sum over all partials.
#pragma once
#include "common.cuh"

namespace rl {

// Every thread of every CTA calls this with its private sums v[NV].
// Returns true only in thread 0 of the LAST CTA to arrive, with `out` = grid totals.
// `ticket` must be 0 on entry and is reset to 0 before returning true.
// partials: [gridDim.x * NV] floats.
template <int NV, int NT>
__device__ __forceinline__ bool grid_reduce(const float (&v)[NV], float* __restrict__ partials,
                                            unsigned* __restrict__ ticket, double (&out)[NV]) {
  __shared__ float s_red[NV][NT / 32];
  __shared__ double s_dred[NV][NT / 32];
  __shared__ bool s_last;
  const int tid = threadIdx.x;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const float w = warp_sum(v[q]);
    if ((tid & 31) == 0) s_red[q][tid >> 5] = w;
  }
  __syncthreads();
  if (tid < NV) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) a += s_red[tid][w];
    partials[blockIdx.x * NV + tid] = a;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  double acc[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) acc[q] = 0.0;
  for (int i = tid; i < (int)gridDim.x; i += NT) {
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] += (double)__ldcg(partials + i * NV + q);
  }
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double x = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((tid & 31) == 0) s_dred[q][tid >> 5] = x;
  }
  __syncthreads();
  if (tid != 0) return false;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    out[q] = 0.0;
    for (int w = 0; w < NT / 32; ++w) out[q] += s_dred[q][w];
  }
  *ticket = 0u;
  return true;
}

}  // namespace rl
