// uint8 input windows for the conv1 kernels (K6): the observation stays uint8 in HBM (the frames a PARL actor
// sends are uint8, examples/IMPALA/atari_agent.py:35-42 divides by 255 in the model) and becomes the bf16
// SWIZZLE_128B operand only in shared memory, inside the kernels that consume it.
//
// Per position tile the producer's TMA brings the DENSE uint8 window [wrows][64 B] into a staging ring; 128
// converter threads (4 per row, 16 input bytes each) write the 128-byte bf16 rows of the operand window exactly
// where a SWIZZLE_128B tensor-map load would have put them — 16-byte chunk j of row r lands at chunk j ^ (r & 7)
// (the window base is 1024-byte aligned) — then fence.proxy.async + mbarrier.arrive hand the window to the
// tcgen05 issuers.  byte -> bf16(byte * scale) uses the same 2^23 magic-number FMA and cvt.rn.bf16x2 as the
// bf16 gather (rl_obs_stack_gather out_dtype 3), so both input forms give bit-identical operands.
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "tma.cuh"

namespace rl {

constexpr int kU8Stages = 3;       // uint8 staging ring depth
constexpr int kU8Threads = 128;    // converter threads (4 warps)

#ifdef __CUDACC__
__device__ __forceinline__ int u8_stage_bytes(int wrows) { return (wrows * 64 + 1023) & ~1023; }

__device__ __forceinline__ void u8_window_to_bf16_sw128(const unsigned char* __restrict__ stage,
                                                        unsigned char* __restrict__ win, int wrows, int ct, float scale,
                                                        float bias) {
  const int q = ct & 3;
  for (int row = ct >> 2; row < wrows; row += kU8Threads / 4) {
    const uint4 in = *reinterpret_cast<const uint4*>(stage + row * 64 + q * 16);
    const uint32_t w[4] = {in.x, in.y, in.z, in.w};
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float f0 = fmaf(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7540u)), scale, bias);
      const float f1 = fmaf(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7541u)), scale, bias);
      const float f2 = fmaf(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7542u)), scale, bias);
      const float f3 = fmaf(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7543u)), scale, bias);
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(pk[2 * i]) : "f"(f1), "f"(f0));
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(pk[2 * i + 1]) : "f"(f3), "f"(f2));
    }
    unsigned char* drow = win + row * 128;
    const int sw = row & 7;
    *reinterpret_cast<uint4*>(drow + (((2 * q) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    *reinterpret_cast<uint4*>(drow + (((2 * q + 1) ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  }
}
#endif

// Host: tensor map of a uint8 [rows][64] matrix, box {64, box_rows}, dense (no swizzle) rows in shared memory.
inline int make_tensor_map_u8_rows64(CUtensorMap* map, const void* base, uint64_t rows, uint32_t box_rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
    fn = reinterpret_cast<EncodeFn>(p);
  }
  const cuuint64_t gdim[2] = {64, rows};
  const cuuint64_t gstride[1] = {64};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0
             : -2;
}

}  // namespace rl
