// Philox4x32-10 and the exact-arithmetic sampling helpers of the RNG contract.
// Device twin of oracle/philox.py: integer ops and single correctly-rounded
// binary32 ops only (__fmul_rn/__fadd_rn, never contracted to FMA), so the CPU
// oracle reproduces every draw bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rl {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

enum : uint32_t {
  STREAM_FRAME = 0,
  STREAM_REWDONE = 1,
  STREAM_ACTION = 2,
  STREAM_OBS = 3,
  STREAM_GAUSS = 4,
  STREAM_REPLAY = 5,
};

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
    const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0, c1 = lo1, c2 = n2, c3 = lo0;
    k0 += kPhiloxW0, k1 += kPhiloxW1;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float u01_24(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }  // 2^-24

// exp(x), x <= 0: t = x*log2e; n = rint(t); 2^(t-n) by degree-7 Horner with
// separately rounded multiply and add; times 2^n.  == oracle.philox.exp_exact.
__device__ __forceinline__ float exp_exact(float x) {
  float t = __fmul_rn(x, 1.4426950408889634f);
  t = fmaxf(t, -126.0f);
  const float n = rintf(t);
  const float f = __fsub_rn(t, n);
  float p = 1.5252733646775596e-05f;
  p = __fadd_rn(__fmul_rn(p, f), 1.540352968731895e-04f);
  p = __fadd_rn(__fmul_rn(p, f), 1.3333557872101665e-03f);
  p = __fadd_rn(__fmul_rn(p, f), 9.618128649890423e-03f);
  p = __fadd_rn(__fmul_rn(p, f), 5.550410971045494e-02f);
  p = __fadd_rn(__fmul_rn(p, f), 2.4022650718688965e-01f);
  p = __fadd_rn(__fmul_rn(p, f), 6.931471824645996e-01f);
  p = __fadd_rn(__fmul_rn(p, f), 1.0f);
  const float scale = __int_as_float(((int)n + 127) << 23);
  const float out = __fmul_rn(p, scale);
  return x < -87.0f ? 0.0f : out;
}

// Inverse-CDF categorical sample with exact arithmetic (== sample_categorical_exact).
__device__ __forceinline__ int sample_categorical_exact(const float* __restrict__ logits, int A, float u) {
  float m = logits[0];
  for (int j = 1; j < A; ++j) m = fmaxf(m, logits[j]);
  float total = 0.f;
  for (int j = 0; j < A; ++j) total = __fadd_rn(total, exp_exact(__fsub_rn(logits[j], m)));
  const float thr = __fmul_rn(u, total);
  float acc = 0.f;
  int a = 0;
  for (int j = 0; j < A; ++j) {
    acc = __fadd_rn(acc, exp_exact(__fsub_rn(logits[j], m)));
    a += (acc <= thr) ? 1 : 0;
  }
  return min(a, A - 1);
}

}  // namespace rl
