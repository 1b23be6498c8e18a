// K6 (convolution weight gradient, TMA-window form) — for a stride-1 NHWC conv written as shifted GEMMs
//     out[q, co] = sum_t sum_ci in[q + off_t, ci] W_t[co, ci]
// the weight gradient is a sum over ALL positions of outer products
//     dW_t[co, ci] = sum_q dout_grid[q, co] * in[q + off_t, ci]
// i.e. per tap a GEMM whose reduction (K) dimension is the position index q.  Both operands are read exactly
// as they lie in HBM — rows = positions, 128 contiguous bytes = 64 channels — through the same TMA windows as
// the forward kernel and consumed by tcgen05.mma as MN-MAJOR operands (a_major = b_major = 1): A = dout tile
// (M = 64 output channels), B = input window starting off_t rows down (N = 64 input channels), K = 16
// positions per instruction.  Accumulators for all (tap, channel-block) pairs stay in TMEM across the CTA's
// whole persistent loop over position tiles; taps are split over CTA groups when they exceed 512 columns.
// Partials are dumped once per CTA and reduced in fixed order by a second kernel (deterministic).
//
// Issue rate.  With M = 64 tiles of K = 16 the tensor core needs only 16-32 cycles per instruction, so ONE
// issuing thread (about 50 cycles per tcgen05.mma even with the descriptors reduced to "add to the low word")
// was the bound (ncu round 1: tensor pipe 18-40 % active, issuer never waiting on data).  Two warps issue now,
// each owning a disjoint half of the CTA's filter taps (= disjoint TMEM columns); both wait on the same "full"
// barrier and both commit to the stage's "empty" barrier (count 2).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tma.cuh"
#include "u8win.cuh"

namespace rl {

__device__ __forceinline__ void w_tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void w_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void w_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void w_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void w_umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void w_commit(void* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(mbar))
               : "memory");
}
__device__ __forceinline__ void w_tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// Shared-memory descriptors.  Only the 14-bit start-address field (address >> 4, shared memory < 256 KB) varies,
// so the issuing threads keep "lo" words = (address >> 4) + kWgLoLbo1 and ADD 16-byte offsets to them.
//   MN-major SWIZZLE_128B (cute canonical ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): 64 MN-elements = one
//     128-byte row per K index, 8 rows per 1024-byte swizzle atom (SBO = 1024 B between K-groups of 8); LBO unused.
//   MN-major SWIZZLE_64B  (((4,n),(8,k)):((1,LBO),(4,SBO)), layout_type 4): 32 MN-elements = one 64-byte row per
//     K index, 8 rows per 512-byte atom.
__device__ __forceinline__ void w_mbar_arrive(void* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(mbar)) : "memory");
}
constexpr uint32_t kWgLoLbo1 = 1u << 16;
constexpr uint32_t kWgHiSw128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t kWgHiSw64 = (uint32_t)(512 >> 4) | (1u << 14) | (4u << 29);
__device__ __forceinline__ uint64_t w_desc(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | (uint64_t)lo; }
// kind::f16, BF16 x BF16 -> F32, A and B MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t w_idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

constexpr int kWgBM = 128;            // positions (K of the GEMM) per tile
constexpr int kWgMaxStages = 6;
constexpr int kWgThreads = 224;       // warp 0 TMA, warps 1 and 6 MMA issuers (disjoint taps), warps 2-5 TMEM dump
constexpr int kWgTapsPerIssuer = 5;   // <= 10 filter taps per CTA group

struct WgradArgs {
  float* partials;                    // [gridDim.x][128 lanes][ncols_max] raw TMEM dumps
  int W, KH, KW;
  int Q, wrows, num_tiles;
  int ngroups, taps_per_group, ncols_max;
  int stages;
};

template <int CBLK>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_window_kernel(const __grid_constant__ CUtensorMap map_dout,
                                                                     const __grid_constant__ CUtensorMap map_x,
                                                                     const WgradArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int win_bytes = (g.wrows * 128 + 1023) & ~1023;
  const int stage_bytes = kWgBM * 128 + CBLK * win_bytes;           // dout tile + input window blocks
  __shared__ __align__(8) unsigned long long full_bar[kWgMaxStages], empty_bar[kWgMaxStages], done_bar;
  const uint32_t nstages = (uint32_t)g.stages;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntaps = g.KH * g.KW;
  const int gid = blockIdx.x % g.ngroups;                           // which slice of the filter taps
  const int tb = gid * g.taps_per_group, te = min(ntaps, tb + g.taps_per_group);
  const int cta_in_group = blockIdx.x / g.ngroups, ctas_per_group = (gridDim.x + g.ngroups - 1 - gid) / g.ngroups;
  const int ncols = (te - tb) * CBLK * 64;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)g.ncols_max) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_dout);
    tma_prefetch_desc(&map_x);
    for (int s = 0; s < kWgMaxStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 2);          // both issuers commit
    }
    mbar_init(&done_bar, 2);
    fence_mbar_init();
  }
  if (warp == 1) w_tmem_alloc(&tmem_base_smem, tmem_cols);
  w_fence_before();
  __syncthreads();
  w_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = cta_in_group; tile < g.num_tiles; tile += ctas_per_group, ++it) {
        const uint32_t s = it % nstages;
        mbar_wait(&empty_bar[s], ((it / nstages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(kWgBM * 128 + CBLK * g.wrows * 128));
        unsigned char* st = smem + s * stage_bytes;
        tma_load_2d(st, &map_dout, 0, tile * kWgBM, &full_bar[s]);
        for (int cb = 0; cb < CBLK; ++cb)
          tma_load_2d(st + kWgBM * 128 + cb * win_bytes, &map_x, cb * 64, tile * kWgBM, &full_bar[s]);
      }
    }
  } else if (warp == 1 || warp == 6) {
    if (lane == 0) {
      constexpr uint32_t idesc = w_idesc_bf16_mn(64, 64);
      // this issuer's taps of the group: local indices [j0, j0 + nmine)
      const int ntg = te - tb, half = (ntg + 1) >> 1;
      const int j0 = warp == 1 ? 0 : half, nmine = warp == 1 ? half : ntg - half;
      uint32_t tap_off[kWgTapsPerIssuer];                    // window shift of each tap in 16-byte units
#pragma unroll
      for (int j = 0; j < kWgTapsPerIssuer; ++j) {
        const int tap = tb + j0 + j, r = tap / g.KW;
        tap_off[j] = (uint32_t)(r * g.W + tap - r * g.KW) * 8u;
      }
      const uint32_t lo0 = (smem_u32(smem) >> 4) + kWgLoLbo1, stage16 = (uint32_t)stage_bytes >> 4;
      const uint32_t win16 = (uint32_t)win_bytes >> 4;
      uint32_t s = 0, par = 0, acc0 = 0;
      for (int tile = cta_in_group; tile < g.num_tiles; tile += ctas_per_group) {
        mbar_wait(&full_bar[s], par);
        w_fence_after();
        const uint32_t a_lo = lo0 + s * stage16, x_lo = a_lo + (uint32_t)(kWgBM * 128 >> 4);
#pragma unroll
        for (int j = 0; j < kWgTapsPerIssuer; ++j) {
          if (j < nmine) {
#pragma unroll
            for (int cb = 0; cb < CBLK; ++cb) {
              const uint32_t d_tmem = tmem_base + (uint32_t)(((j0 + j) * CBLK + cb) * 64);
              const uint32_t b_lo = x_lo + cb * win16 + tap_off[j];
#pragma unroll
              for (int kk = 0; kk < kWgBM / 16; ++kk)   // K advances by 16 positions = 16 rows of 128 bytes (128 units)
                w_umma(d_tmem, w_desc(kWgHiSw128, a_lo + kk * 128u), w_desc(kWgHiSw128, b_lo + kk * 128u), idesc,
                       kk == 0 ? acc0 : 1u);
            }
          }
        }
        w_commit(&empty_bar[s]);
        acc0 = 1u;
        if (++s == nstages) s = 0, par ^= 1u;
      }
      w_commit(&done_bar);
    }
  } else if (warp < 6) {
    // ===== dump the TMEM accumulators once: [128 lanes][ncols] raw (the reduce kernel maps lanes -> rows) =====
    const int qd = warp & 3;
    mbar_wait(&done_bar, 0);
    w_fence_after();
    float* dst = g.partials + ((size_t)blockIdx.x * 128 + qd * 32 + lane) * g.ncols_max;
    const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16);
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      float v[16];
      w_tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
      for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  }
  w_fence_before();
  __syncthreads();
  if (warp == 1) w_tmem_dealloc(tmem_base, tmem_cols);
}

// dW[co][(tap, ci)] = sum over the CTAs of the tap's group, in CTA order (deterministic).
// lane_map 0: accumulator row i of an M=64 tile lives in TMEM lane 32*(i/16) + i%16 ; 1: lane i.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partials, int nctas, int ngroups,
                                                           int taps_per_group, int ntaps, int cblk, int ncols_max,
                                                           int lane_map, float* __restrict__ dw, int accumulate) {
  const int K = ntaps * cblk * 64;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 64 * K) return;
  const int co = idx / K, k = idx - co * K;
  const int tap = k / (cblk * 64), within = k - tap * cblk * 64;
  const int gid = tap / taps_per_group;
  const int col = (tap - gid * taps_per_group) * cblk * 64 + within;
  const int ln = lane_map == 0 ? 32 * (co >> 4) + (co & 15) : co;
  float acc = 0.f;
  for (int c = gid; c < nctas; c += ngroups) acc += partials[((size_t)c * 128 + ln) * ncols_max + col];
  dw[idx] = accumulate ? dw[idx] + acc : acc;
}

// Cout = 32 variant (conv1 of the Atari net): the roles are swapped so that the 64-byte dout rows become the
// N = 32 operand:  D_t[ci, co] = sum_q in[q + off_t, ci] * dout[q, co];  A = input window (MN-major, SW128, M = 64),
// B = dout tile (MN-major, SWIZZLE_64B, N = 32).  One group: KH*KW taps x 32 columns of TMEM.
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_window_n32_kernel(const __grid_constant__ CUtensorMap map_dout,
                                                                         const __grid_constant__ CUtensorMap map_x,
                                                                         const WgradArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int win_bytes = (g.wrows * 128 + 1023) & ~1023;
  const int stage_bytes = kWgBM * 64 + win_bytes;                    // 8 KB dout tile + input window
  __shared__ __align__(8) unsigned long long full_bar[kWgMaxStages], empty_bar[kWgMaxStages], done_bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t nstages = (uint32_t)g.stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntaps = g.KH * g.KW;
  const int ncols = ntaps * 32;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)ncols) tmem_cols <<= 1;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_dout);
    tma_prefetch_desc(&map_x);
    for (int s = 0; s < kWgMaxStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 2);          // both issuers commit
    }
    mbar_init(&done_bar, 2);
    fence_mbar_init();
  }
  if (warp == 1) w_tmem_alloc(&tmem_base_smem, tmem_cols);
  w_fence_before();
  __syncthreads();
  w_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t s = it % nstages;
        mbar_wait(&empty_bar[s], ((it / nstages) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(kWgBM * 64 + g.wrows * 128));
        unsigned char* st = smem + s * stage_bytes;
        tma_load_2d(st + kWgBM * 64, &map_x, 0, tile * kWgBM, &full_bar[s]);     // window first: 1024-byte aligned
        tma_load_2d(st, &map_dout, 0, tile * kWgBM, &full_bar[s]);
      }
    }
  } else if (warp == 1 || warp == 6) {
    if (lane == 0) {
      constexpr uint32_t idesc = w_idesc_bf16_mn(64, 32);
      const int half = (ntaps + 1) >> 1;
      const int j0 = warp == 1 ? 0 : half, nmine = warp == 1 ? half : ntaps - half;
      uint32_t tap_off[kWgTapsPerIssuer];
#pragma unroll
      for (int j = 0; j < kWgTapsPerIssuer; ++j) {
        const int tap = j0 + j, r = tap / g.KW;
        tap_off[j] = (uint32_t)(r * g.W + tap - r * g.KW) * 8u;
      }
      const uint32_t lo0 = (smem_u32(smem) >> 4) + kWgLoLbo1, stage16 = (uint32_t)stage_bytes >> 4;
      uint32_t s = 0, par = 0, acc0 = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        mbar_wait(&full_bar[s], par);
        w_fence_after();
        const uint32_t b_lo = lo0 + s * stage16;                              // dout tile, 64-byte rows
        const uint32_t x_lo = b_lo + (uint32_t)(kWgBM * 64 >> 4);
#pragma unroll
        for (int j = 0; j < kWgTapsPerIssuer; ++j) {
          if (j < nmine) {
            const uint32_t a_lo = x_lo + tap_off[j];
            const uint32_t d_tmem = tmem_base + (uint32_t)((j0 + j) * 32);
#pragma unroll
            for (int kk = 0; kk < kWgBM / 16; ++kk)     // 16 positions = 2048 B of window rows, 1024 B of dout rows
              w_umma(d_tmem, w_desc(kWgHiSw128, a_lo + kk * 128u), w_desc(kWgHiSw64, b_lo + kk * 64u), idesc,
                     kk == 0 ? acc0 : 1u);
          }
        }
        w_commit(&empty_bar[s]);
        acc0 = 1u;
        if (++s == nstages) s = 0, par ^= 1u;
      }
      w_commit(&done_bar);
    }
  } else if (warp < 6) {
    const int qd = warp & 3;
    mbar_wait(&done_bar, 0);
    w_fence_after();
    float* dst = g.partials + ((size_t)blockIdx.x * 128 + qd * 32 + lane) * g.ncols_max;
    const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16);
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      float v[16];
      w_tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
      for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  }
  w_fence_before();
  __syncthreads();
  if (warp == 1) w_tmem_dealloc(tmem_base, tmem_cols);
}

// dW[co][(tap, ci)] (Cout = 32, Cin = 64) from the [ci lanes][tap*32 + co] partials, CTA order (deterministic)
__global__ void __launch_bounds__(256) wgrad_reduce_n32_kernel(const float* __restrict__ partials, int nctas, int ntaps,
                                                               int ncols_max, float* __restrict__ dw, int accumulate) {
  const int K = ntaps * 64;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 32 * K) return;
  const int co = idx / K, k = idx - co * K;
  const int tap = k >> 6, ci = k & 63;
  const int ln = 32 * (ci >> 4) + (ci & 15);
  float acc = 0.f;
  for (int c = 0; c < nctas; ++c) acc += partials[((size_t)c * 128 + ln) * ncols_max + tap * 32 + co];
  dw[idx] = accumulate ? dw[idx] + acc : acc;
}


// ---------------------------------------------------------------------------------------------------------------
// Paired-tap form (default).  Roles swapped — A = input window (MN-major, SWIZZLE_128B), B = dout tile (MN-major;
// SWIZZLE_128B for 64 output channels, SWIZZLE_64B for 32) — and TWO filter taps share one instruction: an
// MN-major operand's second 64-element block lies LBO bytes after the first, and the windows of two taps are the
// same rows shifted by (off_{t+1} - off_t) * 128 bytes, so LBO = that shift gives an M = 128 tile
//     D[(tap_lo | tap_hi, ci), co] += in[q + off, ci]^T dout[q, co]
// at the tensor-core cost of one M = 64 tile.  Half the instructions, all 128 TMEM lanes used, so every tap of a
// 3x3 / 64-channel layer fits the 512 columns at once (5 pairs x 64) and no position tile is read by two CTA groups
// (the M = 64 form read conv3's operands twice: ncu round 1, 11.9 GB for 6.3 GB of unique data).  An odd last tap
// is paired with the window one row further down; its upper 64 lanes are never read back.
// The bias gradient rides along: one more M = 64 instruction per K step whose A operand is a constant tile of ones,
//     D_b[m, co] += sum_q 1 * dout[q, co]        (every row m holds the column sums; lane 0 is read back)
// so the gradient grid is not streamed from HBM a second time by a column-sum kernel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWgSlotsPerIssuer = 3;      // (pair, channel-block) accumulators per issuing warp

struct WgradPairArgs {
  float* partials;                    // [gridDim.x][128 lanes][ncols] raw TMEM dumps
  int W, KW, ntaps;
  int wrows, num_tiles, stages;
  int npairs, ncols;                  // ncols = npairs * CBLK * COUT (+ 2 * COUT bias-gradient columns)
  int bias;                           // 1: also accumulate the column sums of dout (bias gradient) in TMEM
  float in_scale;                     // U8X: input operand = bf16(byte * in_scale)
};

// U8X (conv1 on the uint8 observation): map_x is the uint8 [Q][64] matrix; the producer fills a dense staging ring
// and eight warps — the four dump warps (idle until the last tile) plus four extra ones (7-10) — convert each window
// into the bf16 SWIZZLE_128B slot (u8win.cuh).
constexpr int kWgU8Threads = kWgThreads + kU8Threads - 128;
template <int COUT, int CBLK, bool U8X = false>
__global__ void __launch_bounds__(U8X ? kWgU8Threads : kWgThreads, 1) wgrad_pair_kernel(const __grid_constant__ CUtensorMap map_dout,
                                                                   const __grid_constant__ CUtensorMap map_x,
                                                                   const WgradPairArgs g) {
  static_assert(!U8X || CBLK == 1, "the uint8 window is one 64-channel block");
  constexpr int DOUT_BYTES = kWgBM * COUT * 2;                        // 16 KB (SW128) or 8 KB (SW64)
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int win_bytes = (g.wrows * 128 + 1023) & ~1023;
  const int stage_bytes = CBLK * win_bytes + DOUT_BYTES;              // windows first (1024-byte aligned), then dout
  __shared__ __align__(8) unsigned long long full_bar[kWgMaxStages], empty_bar[kWgMaxStages], done_bar;
  __shared__ __align__(8) unsigned long long u8_full[kU8Stages], u8_empty[kU8Stages];
  __shared__ uint32_t tmem_base_smem;
  const uint32_t nstages = (uint32_t)g.stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)g.ncols) tmem_cols <<= 1;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_dout);
    tma_prefetch_desc(&map_x);
    for (int s = 0; s < kWgMaxStages; ++s) {
      mbar_init(&full_bar[s], U8X ? 1 + kU8Threads : 1);   // U8X: the dout TMA + every converter thread
      mbar_init(&empty_bar[s], 2);          // both issuers commit
    }
    for (int s = 0; s < kU8Stages; ++s) {
      mbar_init(&u8_full[s], 1);
      mbar_init(&u8_empty[s], kU8Threads);
    }
    mbar_init(&done_bar, 2);
    fence_mbar_init();
  }
  if (warp == 1) w_tmem_alloc(&tmem_base_smem, tmem_cols);
  // constant A operand of the bias-gradient instruction: 16 K-rows x 64 bf16 ones, after the stage ring
  unsigned char* s_ones = smem + nstages * stage_bytes;
  unsigned char* sStage = s_ones + 2048;                              // U8X: [kU8Stages][wrows][64 B]
  if (g.bias) {
    for (int i = threadIdx.x; i < 512; i += kWgThreads) reinterpret_cast<uint32_t*>(s_ones)[i] = 0x3F803F80u;
    fence_proxy_async_smem();               // generic-proxy stores -> visible to the tensor core's async proxy
  }
  w_fence_before();
  __syncthreads();
  w_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t s = 0, par = 1, ss = 0, spar = 1;
      const int sbytes = u8_stage_bytes(g.wrows);
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        if (U8X) {
          mbar_wait(&u8_empty[ss], spar);
          mbar_arrive_expect_tx(&u8_full[ss], (uint32_t)(g.wrows * 64));
          tma_load_2d(sStage + ss * sbytes, &map_x, 0, tile * kWgBM, &u8_full[ss]);
          if (++ss == kU8Stages) ss = 0, spar ^= 1u;
        }
        mbar_wait(&empty_bar[s], par);
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(DOUT_BYTES + (U8X ? 0 : CBLK * g.wrows * 128)));
        unsigned char* st = smem + s * stage_bytes;
        if (!U8X) {
#pragma unroll
          for (int cb = 0; cb < CBLK; ++cb) tma_load_2d(st + cb * win_bytes, &map_x, cb * 64, tile * kWgBM, &full_bar[s]);
        }
        tma_load_2d(st + CBLK * win_bytes, &map_dout, 0, tile * kWgBM, &full_bar[s]);
        if (++s == nstages) s = 0, par ^= 1u;
      }
    }
  } else if (warp == 1 || warp == 6) {
    if (lane == 0) {
      constexpr uint32_t idesc = w_idesc_bf16_mn(128, COUT);
      constexpr uint32_t hiB = COUT == 64 ? kWgHiSw128 : kWgHiSw64;
      constexpr uint32_t kstepB = COUT == 64 ? 128u : 64u;             // 16 positions of dout rows, 16-byte units
      // this issuer's accumulator slots: slot = pair * CBLK + cb, local indices [j0, j0 + nmine)
      const int nslots = g.npairs * CBLK, half = (nslots + 1) >> 1;
      const int j0 = warp == 1 ? 0 : half, nmine = warp == 1 ? half : nslots - half;
      uint32_t a_off[kWgSlotsPerIssuer];      // (window shift of the pair's first tap + channel block) | LBO << 16
#pragma unroll
      for (int j = 0; j < kWgSlotsPerIssuer; ++j) {
        const int slot = j0 + j, pair = slot / CBLK, cb = slot - pair * CBLK;
        const int t0 = 2 * pair, t1 = min(2 * pair + 1, g.ntaps - 1);
        const int r0 = t0 / g.KW, r1 = t1 / g.KW;
        const int o0 = r0 * g.W + t0 - r0 * g.KW, o1 = r1 * g.W + t1 - r1 * g.KW;
        const int delta = o1 > o0 ? o1 - o0 : 1;                       // odd last tap: dummy partner one row down
        a_off[j] = (uint32_t)(o0 * 8 + cb * (win_bytes >> 4)) + ((uint32_t)(delta * 8) << 16);
      }
      const uint32_t lo0 = smem_u32(smem) >> 4, stage16 = (uint32_t)stage_bytes >> 4;
      constexpr uint32_t idesc_bias = w_idesc_bf16_mn(64, COUT);
      // bias gradient: each issuer takes every other K step into its OWN accumulator (no ordering between issuers)
      const bool do_bias = g.bias != 0;
      const uint32_t issuer = warp == 1 ? 0u : 1u;
      const uint32_t ones_lo = (smem_u32(s_ones) >> 4) + kWgLoLbo1;
      const uint32_t d_bias = tmem_base + (uint32_t)(nslots * COUT) + issuer * COUT;
      uint32_t s = 0, par = 0, acc0 = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        mbar_wait(&full_bar[s], par);
        w_fence_after();
        const uint32_t x_lo = lo0 + s * stage16;
        const uint32_t b_lo = x_lo + (uint32_t)((CBLK * win_bytes) >> 4) + kWgLoLbo1;
#pragma unroll
        for (int j = 0; j < kWgSlotsPerIssuer; ++j) {
          if (j < nmine) {
            const uint32_t d_tmem = tmem_base + (uint32_t)((j0 + j) * COUT);
            const uint32_t a_lo = x_lo + a_off[j];
#pragma unroll
            for (int kk = 0; kk < kWgBM / 16; ++kk)       // 16 positions = 16 window rows (128 units) per K step
              w_umma(d_tmem, w_desc(kWgHiSw128, a_lo + kk * 128u), w_desc(hiB, b_lo + kk * kstepB), idesc,
                     kk == 0 ? acc0 : 1u);
          }
        }
        if (do_bias) {
#pragma unroll
          for (int k2 = 0; k2 < kWgBM / 32; ++k2)
            w_umma(d_bias, w_desc(kWgHiSw128, ones_lo), w_desc(hiB, b_lo + (2u * k2 + issuer) * kstepB), idesc_bias,
                   k2 == 0 ? acc0 : 1u);
        }
        w_commit(&empty_bar[s]);
        acc0 = 1u;
        if (++s == nstages) s = 0, par ^= 1u;
      }
      w_commit(&done_bar);
    }
  } else {
    // ===== dump the TMEM accumulators once: [128 lanes][ncols] =====
    const int qd = warp & 3;
    if (U8X) {
      // uint8 -> bf16 window converters (these warps have nothing else to do until the accumulators are final)
      const int ct = warp < 6 ? threadIdx.x - 64 : threadIdx.x - kWgThreads + 128;
      const int sbytes = u8_stage_bytes(g.wrows);
      const float cbias = -8388608.0f * g.in_scale;
      uint32_t s = 0, epar = 1, ss = 0, fpar = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        mbar_wait(&u8_full[ss], fpar);
        mbar_wait(&empty_bar[s], epar);            // the MMAs that read this slot have completed
        w_fence_after();
        u8_window_to_bf16_sw128(sStage + ss * sbytes, smem + s * stage_bytes, g.wrows, ct, g.in_scale, cbias);
        fence_proxy_async_smem();
        w_mbar_arrive(&full_bar[s]);
        w_mbar_arrive(&u8_empty[ss]);
        if (++s == nstages) s = 0, epar ^= 1u;
        if (++ss == kU8Stages) ss = 0, fpar ^= 1u;
      }
    }
    if (warp < 6) {
      mbar_wait(&done_bar, 0);
      w_fence_after();
      float* dst = g.partials + ((size_t)blockIdx.x * 128 + qd * 32 + lane) * g.ncols;
      const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16);
      for (int c0 = 0; c0 < g.ncols; c0 += 16) {
        float v[16];
        w_tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
  }
  w_fence_before();
  __syncthreads();
  if (warp == 1) w_tmem_dealloc(tmem_base, tmem_cols);
}

// dW[co][(tap, cb, ci)] from the [lane = (tap & 1) * 64 + ci][col = ((tap >> 1) * cblk + cb) * cout + co] partials,
// summed over the CTAs in index order (deterministic); db[co] from lane 0 of the bias-gradient columns
__global__ void __launch_bounds__(256) wgrad_pair_reduce_kernel(const float* __restrict__ partials, int nctas, int ntaps,
                                                                int cblk, int cout, int ncols, float* __restrict__ dw,
                                                                float* __restrict__ db, int accumulate) {
  const int K = ntaps * cblk * 64;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cout * K) {
    const int co = idx - cout * K;
    if (db && co < cout) {
      const int col = ((ntaps + 1) / 2) * cblk * cout + co;          // two accumulators (one per issuer): col, col + cout
      float acc = 0.f;
      for (int c = 0; c < nctas; ++c) acc += partials[(size_t)c * 128 * ncols + col] + partials[(size_t)c * 128 * ncols + col + cout];
      db[co] = accumulate ? db[co] + acc : acc;
    }
    return;
  }
  const int co = idx / K, k = idx - co * K;
  const int tap = k / (cblk * 64), within = k - tap * cblk * 64, cb = within >> 6, ci = within & 63;
  const int ln = (tap & 1) * 64 + ci, col = ((tap >> 1) * cblk + cb) * cout + co;
  float acc = 0.f;
  for (int c = 0; c < nctas; ++c) acc += partials[((size_t)c * 128 + ln) * ncols + col];
  dw[idx] = accumulate ? dw[idx] + acc : acc;
}

// column sums of a [rows, C] bf16 matrix (bias gradients): deterministic two-stage, 16-byte loads.
// A thread owns 8 adjacent columns (one uint4) and walks the rows of its block's chunk with stride
// (256 / (C/8)) so that a warp reads whole 128-byte lines.
__global__ void __launch_bounds__(256) colsum_bf16_stage1(const __nv_bfloat16* __restrict__ x, long long rows, int C,
                                                          float* __restrict__ part) {
  const int vpr = C >> 3;                              // uint4 vectors per row
  const long long chunk = (rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * chunk, r1 = min(rows, r0 + chunk);
  __shared__ float sm[256][8];
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (vpr <= 256) {
    const int nph = 256 / vpr;                         // row phases handled concurrently by the block
    const int v = threadIdx.x % vpr, ph = threadIdx.x / vpr;
    if (ph < nph) {
      const uint4* base = reinterpret_cast<const uint4*>(x) + v;
      for (long long r = r0 + ph; r < r1; r += nph) {
        const uint4 q = __ldcs(base + r * vpr);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[2 * i] += __uint_as_float(w[i] << 16);
          acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[threadIdx.x][i] = acc[i];
    __syncthreads();
    if (threadIdx.x < vpr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a = 0.f;
        for (int p = 0; p < nph; ++p) a += sm[p * vpr + threadIdx.x][i];
        part[(size_t)blockIdx.x * C + threadIdx.x * 8 + i] = a;
      }
    }
  } else {
    for (int c = threadIdx.x; c < C; c += 256) {       // very wide matrices: scalar fallback
      float a = 0.f;
      for (long long r = r0; r < r1; ++r) a += __bfloat162float(x[r * C + c]);
      part[(size_t)blockIdx.x * C + c] = a;
    }
  }
}
__global__ void colsum_stage2(const float* __restrict__ part, int nblocks, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f;
  for (int b = 0; b < nblocks; ++b) a += part[(size_t)b * C + c];
  out[c] = a;
}

static int wg_make_map(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint32_t box_rows,
                       uint32_t box_cols = 64, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
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
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {cols * 2};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0
             : -2;
}

}  // namespace rl

using namespace rl;

// bit 0: legacy M = 64 kernels read accumulator row i from TMEM lane i (wrong on B200; the measured map is
// 32*(i/16) + i%16) — triage only.  bit 1: use the legacy one-tap-per-instruction (M = 64) kernels instead of the
// paired-tap form.
static int g_wg_lane_map = 0;
extern "C" int rl_debug_set_wgrad_lane_map(int mode) {
  g_wg_lane_map = mode;
  return RL_OK;
}

extern "C" size_t rl_conv_wgrad_workspace_bytes(int KH, int KW, int Cin) {
  const int cblk = Cin / 64, ntaps = KH * KW;
  int per = 512 / (cblk * 64);
  if (per > ntaps) per = ntaps;
  return (size_t)160 * 128 * (size_t)(per * cblk * 64) * sizeof(float) + 4096;
}

static int wgrad_legacy_bias(const void* dout_grid, long long Q, int Cout, float* db, int accumulate, void* workspace,
                             size_t workspace_bytes, rl_stream_t stream) {
  if (!db) return RL_OK;
  RL_CHECK_ARG(!accumulate, "conv2d_s1_wgrad: accumulate with a bias gradient needs the paired-tap form");
  return rl_colsum_bf16(dout_grid, Q, Cout, db, workspace, workspace_bytes, stream);
}

static int wgrad_dispatch(const void* dout_grid, const void* in, float* dw_krsc, float* db, int N, int H, int W, int Cin,
                          int Cout, int KH, int KW, int accumulate, void* workspace, size_t workspace_bytes,
                          rl_stream_t stream, int u8in, float in_scale) {
  RL_CHECK_ARG(dout_grid && in && dw_krsc && workspace && N > 0, "conv2d_s1_wgrad: bad argument");
  RL_CHECK_ARG(aligned16(dout_grid) && aligned16(in) && aligned16(workspace), "conv2d_s1_wgrad: alignment");
  RL_CHECK_ARG((Cout == 64 && (Cin == 64 || Cin == 128)) || (Cout == 32 && Cin == 64),
               "conv2d_s1_wgrad: (Cout, Cin) must be (64, 64|128) or (32, 64)");
  const long long Q = (long long)N * H * W;
  RL_CHECK_ARG(Q < (1LL << 31), "conv2d_s1_wgrad: too many positions");
  const int cblk = Cin / 64, ntaps = KH * KW;
  WgradArgs g;
  g.partials = reinterpret_cast<float*>(workspace);
  g.W = W, g.KH = KH, g.KW = KW, g.Q = (int)Q;
  g.wrows = kWgBM + (KH - 1) * W + (KW - 1);
  RL_CHECK_ARG(g.wrows <= 256, "conv2d_s1_wgrad: window too tall");
  g.num_tiles = (int)((Q + kWgBM - 1) / kWgBM);
  const int npairs = (ntaps + 1) / 2;
  if (u8in || (!(g_wg_lane_map & 2) && npairs * cblk <= 2 * kWgSlotsPerIssuer && npairs * cblk * Cout <= 512)) {
    // ---- paired-tap form (default) ----
    WgradPairArgs a;
    a.partials = reinterpret_cast<float*>(workspace);
    a.W = W, a.KW = KW, a.ntaps = ntaps, a.wrows = g.wrows + 1;       // + the dummy partner row of an odd last tap
    a.num_tiles = g.num_tiles, a.npairs = npairs, a.bias = db ? 1 : 0, a.in_scale = in_scale;
    a.ncols = npairs * cblk * Cout + (db ? 2 * Cout : 0);            // + one bias accumulator per issuer
    RL_CHECK_ARG(a.ncols <= 512, "conv2d_s1_wgrad: accumulators exceed the 512 TMEM columns");
    RL_CHECK_ARG(a.wrows <= 256, "conv2d_s1_wgrad: window too tall");
    int devp = 0, smsp = 148;
    cudaGetDevice(&devp);
    cudaDeviceGetAttribute(&smsp, cudaDevAttrMultiProcessorCount, devp);
    smsp = effective_sms(smsp);
    const int gridp = smsp < a.num_tiles ? smsp : a.num_tiles;
    if (workspace_bytes < (size_t)gridp * 128 * a.ncols * sizeof(float)) {
      set_error("conv2d_s1_wgrad: workspace too small");
      return RL_ERR_WORKSPACE;
    }
    alignas(64) CUtensorMap mdp, mxp;
    if (wg_make_map(&mdp, dout_grid, (uint64_t)Cout, (uint64_t)Q, kWgBM, (uint32_t)Cout,
                    Cout == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B) ||
        (u8in ? make_tensor_map_u8_rows64(&mxp, in, (uint64_t)Q, (uint32_t)a.wrows)
              : wg_make_map(&mxp, in, (uint64_t)Cin, (uint64_t)Q, (uint32_t)a.wrows))) {
      set_error("conv2d_s1_wgrad: cuTensorMapEncodeTiled failed");
      return RL_ERR_CUDA;
    }
    const size_t winp = (size_t)((a.wrows * 128 + 1023) & ~1023);
    const size_t stagep = (size_t)cblk * winp + (size_t)kWgBM * Cout * 2;
    const size_t u8ring = u8in ? (size_t)kU8Stages * (size_t)((a.wrows * 64 + 1023) & ~1023) : 0;
    long long nstp = (long long)((216 * 1024 - u8ring) / stagep);
    a.stages = (int)(nstp > kWgMaxStages ? kWgMaxStages : (nstp < 2 ? 2 : nstp));
    const size_t smemp = (size_t)a.stages * stagep + 2048 + 1024 + u8ring;   // + the 2 KB tile of ones (+ uint8 staging)
    cudaStream_t stp = (cudaStream_t)stream;
    if (u8in) {
      auto kern = wgrad_pair_kernel<32, 1, true>;
      RL_SMEM_OPTIN(kern);
      kern<<<gridp, kWgU8Threads, smemp, stp>>>(mdp, mxp, a);
    } else if (Cout == 64 && cblk == 1) {
      RL_SMEM_OPTIN(wgrad_pair_kernel<64, 1>);
      wgrad_pair_kernel<64, 1><<<gridp, kWgThreads, smemp, stp>>>(mdp, mxp, a);
    } else if (Cout == 64) {
      RL_SMEM_OPTIN(wgrad_pair_kernel<64, 2>);
      wgrad_pair_kernel<64, 2><<<gridp, kWgThreads, smemp, stp>>>(mdp, mxp, a);
    } else {
      RL_SMEM_OPTIN(wgrad_pair_kernel<32, 1>);
      wgrad_pair_kernel<32, 1><<<gridp, kWgThreads, smemp, stp>>>(mdp, mxp, a);
    }
    const int Kp = ntaps * cblk * 64;
    wgrad_pair_reduce_kernel<<<(Cout * Kp + Cout + 255) / 256, 256, 0, stp>>>(a.partials, gridp, ntaps, cblk, Cout,
                                                                             a.ncols, dw_krsc, db, accumulate);
    RL_CHECK_LAUNCH("rl_conv2d_s1_nhwc_bf16_wgrad");
    return RL_OK;
  }
  if (Cout == 32) {
    // swapped-role variant: D[ci, co], dout rows are 64 bytes (SWIZZLE_64B operand)
    RL_CHECK_ARG(ntaps <= 2 * kWgTapsPerIssuer, "conv2d_s1_wgrad: too many taps for Cout = 32");
    g.ngroups = 1, g.taps_per_group = ntaps, g.ncols_max = ntaps * 32;
    int dev32 = 0, sms32 = 148;
    cudaGetDevice(&dev32);
    cudaDeviceGetAttribute(&sms32, cudaDevAttrMultiProcessorCount, dev32);
    sms32 = effective_sms(sms32);
    int grid32 = sms32 < g.num_tiles ? sms32 : g.num_tiles;
    if (workspace_bytes < (size_t)grid32 * 128 * g.ncols_max * sizeof(float)) {
      set_error("conv2d_s1_wgrad: workspace too small");
      return RL_ERR_WORKSPACE;
    }
    alignas(64) CUtensorMap md32, mx32;
    if (wg_make_map(&md32, dout_grid, 32, (uint64_t)Q, kWgBM, 32, CU_TENSOR_MAP_SWIZZLE_64B) ||
        wg_make_map(&mx32, in, 64, (uint64_t)Q, (uint32_t)g.wrows)) {
      set_error("conv2d_s1_wgrad: cuTensorMapEncodeTiled failed");
      return RL_ERR_CUDA;
    }
    const size_t win32 = (size_t)((g.wrows * 128 + 1023) & ~1023);
    long long nst32 = (long long)((218 * 1024) / (kWgBM * 64 + win32));
    g.stages = (int)(nst32 > kWgMaxStages ? kWgMaxStages : (nst32 < 2 ? 2 : nst32));
    const size_t smem32 = (size_t)g.stages * (kWgBM * 64 + win32) + 1024;
    RL_SMEM_OPTIN(wgrad_window_n32_kernel);
    wgrad_window_n32_kernel<<<grid32, kWgThreads, smem32, (cudaStream_t)stream>>>(md32, mx32, g);
    wgrad_reduce_n32_kernel<<<(32 * ntaps * 64 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        g.partials, grid32, ntaps, g.ncols_max, dw_krsc, accumulate);
    RL_CHECK_LAUNCH("rl_conv2d_s1_nhwc_bf16_wgrad");
    return wgrad_legacy_bias(dout_grid, Q, Cout, db, accumulate, workspace, workspace_bytes, stream);
  }
  int cap = 512 / (cblk * 64);                       // taps whose accumulators fit the 512 TMEM columns
  if (cap > ntaps) cap = ntaps;
  g.ngroups = (ntaps + cap - 1) / cap;
  g.taps_per_group = (ntaps + g.ngroups - 1) / g.ngroups;   // balanced split (e.g. 9 taps -> 5 + 4)
  g.ncols_max = g.taps_per_group * cblk * 64;
  RL_CHECK_ARG(g.taps_per_group <= 2 * kWgTapsPerIssuer, "conv2d_s1_wgrad: too many taps per CTA group");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  sms = effective_sms(sms);
  int grid = sms;
  if (grid > g.num_tiles * g.ngroups) grid = g.num_tiles * g.ngroups;
  if (grid < g.ngroups) grid = g.ngroups;
  if (workspace_bytes < (size_t)grid * 128 * g.ncols_max * sizeof(float)) {
    set_error("conv2d_s1_wgrad: workspace too small");
    return RL_ERR_WORKSPACE;
  }
  alignas(64) CUtensorMap md, mx;
  if (wg_make_map(&md, dout_grid, (uint64_t)Cout, (uint64_t)Q, kWgBM) ||
      wg_make_map(&mx, in, (uint64_t)Cin, (uint64_t)Q, (uint32_t)g.wrows)) {
    set_error("conv2d_s1_wgrad: cuTensorMapEncodeTiled failed");
    return RL_ERR_CUDA;
  }
  const size_t win = (size_t)((g.wrows * 128 + 1023) & ~1023);
  long long nst = (long long)((218 * 1024) / (kWgBM * 128 + cblk * win));
  g.stages = (int)(nst > kWgMaxStages ? kWgMaxStages : (nst < 2 ? 2 : nst));
  const size_t smem = (size_t)g.stages * (kWgBM * 128 + cblk * win) + 1024;
  cudaStream_t st = (cudaStream_t)stream;
  if (cblk == 1) {
    RL_SMEM_OPTIN(wgrad_window_kernel<1>);
    wgrad_window_kernel<1><<<grid, kWgThreads, smem, st>>>(md, mx, g);
  } else {
    RL_SMEM_OPTIN(wgrad_window_kernel<2>);
    wgrad_window_kernel<2><<<grid, kWgThreads, smem, st>>>(md, mx, g);
  }
  const int K = ntaps * cblk * 64;
  wgrad_reduce_kernel<<<(64 * K + 255) / 256, 256, 0, st>>>(g.partials, grid, g.ngroups, g.taps_per_group, ntaps, cblk,
                                                           g.ncols_max, g_wg_lane_map & 1, dw_krsc, accumulate);
  RL_CHECK_LAUNCH("rl_conv2d_s1_nhwc_bf16_wgrad");
  return wgrad_legacy_bias(dout_grid, Q, Cout, db, accumulate, workspace, workspace_bytes, stream);
}

extern "C" int rl_conv2d_s1_nhwc_bf16_wgrad(const void* dout_grid, const void* in, float* dw_krsc, float* db, int N,
                                            int H, int W, int Cin, int Cout, int KH, int KW, int accumulate,
                                            void* workspace, size_t workspace_bytes, rl_stream_t stream) {
  return wgrad_dispatch(dout_grid, in, dw_krsc, db, N, H, W, Cin, Cout, KH, KW, accumulate, workspace, workspace_bytes,
                        stream, 0, 1.f);
}

extern "C" int rl_conv2d_s1_u8in_bf16_wgrad(const void* dout_grid, const void* in_u8, float in_scale, float* dw_krsc,
                                            float* db, int N, int H, int W, int Cout, int KH, int KW, int accumulate,
                                            void* workspace, size_t workspace_bytes, rl_stream_t stream) {
  RL_CHECK_ARG(Cout == 32 && KH == 2 && KW == 2, "conv2d_s1_u8in_wgrad: built for the 2x2, 64 -> 32 layer (conv1, space-to-depth)");
  return wgrad_dispatch(dout_grid, in_u8, dw_krsc, db, N, H, W, 64, Cout, KH, KW, accumulate, workspace, workspace_bytes,
                        stream, 1, in_scale);
}

extern "C" int rl_colsum_bf16(const void* x, long long rows, int C, float* out, void* workspace, size_t workspace_bytes,
                              rl_stream_t stream) {
  RL_CHECK_ARG(x && out && workspace && rows > 0 && C >= 8 && C % 8 == 0 && aligned16(x) &&
                   ((C / 8) > 256 || 256 % (C / 8) == 0),
               "colsum_bf16: C must be a multiple of 8 with C/8 dividing 256 (or C > 2048)");
  const int nblocks = 1184;
  if (workspace_bytes < (size_t)nblocks * C * sizeof(float)) {
    set_error("colsum_bf16: workspace too small (need %d*C*4 bytes)", nblocks);
    return RL_ERR_WORKSPACE;
  }
  colsum_bf16_stage1<<<nblocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, rows, C, (float*)workspace);
  colsum_stage2<<<(C + 63) / 64, 64, 0, (cudaStream_t)stream>>>((const float*)workspace, nblocks, C, out);
  RL_CHECK_LAUNCH("rl_colsum_bf16");
  return RL_OK;
}
