// K5 / K7 — on-device vectorised actor pool: synthetic env steppers + action sampling.
//
// Replaces, for thousands of envs in lock-step, the reference's serial Python loop
//   parl/env/vector_env.py:41-63      VectorEnv.step with auto-reset on done
//   parl/tests/gym.py:117-207         mock CartPole / Pong / HalfCheetah distributions
//   parl/env/atari_wrappers.py:270-307 FrameStack (reset fills all k slots)
//   examples/IMPALA/atari_agent.py:39-40  per-row np.random.choice action sampling
//   parl/algorithms/torch/ppo.py:164-177  Normal / Categorical sampling + log-prob
// State is SoA in HBM ([B] arrays), frames are written once into a ring of
// planes [P, B, H*W] uint8 (the 4-frame stack is virtual: age[t,b] says how many
// older planes belong to the same episode), done masks are aggregated with
// __ballot_sync for the episode statistics.
// RNG contract: Philox4x32-10, counter (env_id, step, block, stream) — philox.cuh.
#include <cuda_bf16.h>

#include "common.cuh"
#include "philox.cuh"
#include "env_common.cuh"

namespace rl {

// 16 pixels of frame n of env e: bytes of one Philox block mapped to U{0..254}.
__device__ __forceinline__ uint4 frame_block(uint32_t env, uint32_t n, uint32_t blk, uint32_t k0, uint32_t k1) {
  uint4 x = philox4x32_10(env, n, blk, STREAM_FRAME, k0, k1);
  // per byte: max(b,1)-1  == (b*255)>>8  (parl/tests/gym.py:165 randint(0,255): high exclusive)
  x.x = __vsub4(__vmaxu4(x.x, 0x01010101u), 0x01010101u);
  x.y = __vsub4(__vmaxu4(x.y, 0x01010101u), 0x01010101u);
  x.z = __vsub4(__vmaxu4(x.z, 0x01010101u), 0x01010101u);
  x.w = __vsub4(__vmaxu4(x.w, 0x01010101u), 0x01010101u);
  return x;
}

struct AtariStepArgs {
  uint8_t* frame_out;       // [B, HW] plane receiving frame (step+1)
  float* reward_out;        // [B]
  uint8_t* done_out;        // [B]
  const uint8_t* age_in;    // [B] (may be NULL on reset)
  uint8_t* age_out;         // [B]
  const float* logits;      // [B, A] or NULL
  int* actions_out;         // [B] (written when logits != NULL)
  EpisodeStats st;
  int B, HW, A;
  uint32_t k0, k1, step, env_offset, done_thr;
  const uint32_t* step_dev; // if non-NULL the step index is read from device memory (CUDA-graph replay)
  int reset;                // 1: only emit frame `step` and zero the state
};

__global__ void __launch_bounds__(256) atari_synth_step_kernel(AtariStepArgs p) {
  pdl_wait();            // chain kernel (launch_chain): the logits / step counter come from the previous kernels
  pdl_trigger();
  if (p.step_dev) p.step = *p.step_dev;
  const int nblk = p.HW >> 4;
  const long long total = (long long)p.B * nblk;
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  // ---- per-env scalars: reward, done, age, episode stats, (optional) action sampling ----
  const int nwarp_env = (p.B + 31) >> 5;
  if ((gtid >> 5) < nwarp_env) {
    const int b = (int)gtid;
    const bool valid = b < p.B;
    if (p.reset) {
      if (valid) {
        p.age_out[b] = 0;
        p.st.ep_ret[b] = 0.f;
        p.st.ep_len[b] = 0;
      }
    } else {
      float reward = 0.f;
      bool done = false;
      if (valid) {
        const uint32_t env = p.env_offset + (uint32_t)b;
        const uint4 x = philox4x32_10(env, p.step, 0u, STREAM_REWDONE, p.k0, p.k1);
        reward = (float)(x.x & 1u);
        done = x.y < p.done_thr;
        p.reward_out[b] = reward;
        p.done_out[b] = done ? 1 : 0;
        const int age = p.age_in[b];
        p.age_out[b] = done ? 0 : (uint8_t)min(age + 1, 3);
        if (p.logits) {
          const uint4 ua = philox4x32_10(env, p.step, 0u, STREAM_ACTION, p.k0, p.k1);
          p.actions_out[b] = sample_categorical_exact(p.logits + (long long)b * p.A, p.A, u01_24(ua.x));
        }
      }
      episode_update(p.st, b, valid, reward, done);
    }
  }
  // ---- frame pixels: one 16-byte Philox block per thread, coalesced uint4 stores ----
  // 32-bit, division-free walk: (env, blk) advance by (gstride / nblk, gstride % nblk) per grid stride
  const uint32_t n = p.reset ? p.step : p.step + 1u;
  uint4* out = reinterpret_cast<uint4*>(p.frame_out);
  const uint32_t unblk = (uint32_t)nblk;
  const uint32_t ustride = (uint32_t)gstride;
  const uint32_t db = ustride / unblk, dk = ustride - db * unblk;
  uint32_t b = (uint32_t)gtid / unblk, blk = (uint32_t)gtid - b * unblk;
  for (uint32_t i = (uint32_t)gtid; i < (uint32_t)total; i += ustride) {
    out[i] = frame_block(p.env_offset + b, n, blk, p.k0, p.k1);
    b += db, blk += dk;
    if (blk >= unblk) blk -= unblk, ++b;
  }
}

// ---------------------------------------------------------------------------
// Virtual frame stack -> materialised observations.
// obs(t,b) channel j (0 = oldest) = plane[t + 3 - min(3 - j, age[t,b])].
// ---------------------------------------------------------------------------
template <typename OutT>
__device__ __forceinline__ void store16(OutT* dst, uint4 v, float scale);
template <>
__device__ __forceinline__ void store16<uint8_t>(uint8_t* dst, uint4 v, float) {
  *reinterpret_cast<uint4*>(dst) = v;
}
template <>
__device__ __forceinline__ void store16<float>(float* dst, uint4 v, float scale) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float4 f;
    f.x = (float)(w[k] & 0xffu) * scale;
    f.y = (float)((w[k] >> 8) & 0xffu) * scale;
    f.z = (float)((w[k] >> 16) & 0xffu) * scale;
    f.w = (float)(w[k] >> 24) * scale;
    reinterpret_cast<float4*>(dst)[k] = f;
  }
}

template <typename OutT>
__global__ void __launch_bounds__(256) obs_stack_gather_kernel(const uint8_t* __restrict__ planes,
                                                               const uint8_t* __restrict__ ages, int B, int HW,
                                                               int t_begin, int t_count, int env_major, float scale,
                                                               OutT* __restrict__ out) {
  const int nblk = HW >> 4;
  const long long total = (long long)t_count * B * 4 * nblk;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gstride) {
    const int blk = (int)(i % nblk);
    long long r = i / nblk;
    const int j = (int)(r & 3);
    r >>= 2;                                   // sample index in output order
    int t, b;
    if (env_major) {
      b = (int)(r / t_count), t = (int)(r - (long long)b * t_count);
    } else {
      t = (int)(r / B), b = (int)(r - (long long)t * B);
    }
    t += t_begin;
    const int age = ages[(long long)t * B + b];
    const int plane = t + 3 - min(3 - j, age);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(planes + ((long long)plane * B + b) * HW) + blk);
    store16<OutT>(out + (r * 4 + j) * HW + (blk << 4), v, scale);
  }
}


// Same gather, fused with the network's input transform: out[sample, pixel, channel] (NHWC) in bf16,
// value * scale (scale = 1/255 reproduces ``obs / 255.0`` of the reference models).  One thread = 16 pixels
// x 4 channels: four 16-byte plane reads, one 128-byte contiguous write.
__global__ void __launch_bounds__(256) obs_stack_gather_nhwc_bf16_kernel(const uint8_t* __restrict__ planes,
                                                                         const uint8_t* __restrict__ ages, int B,
                                                                         int HW, int t_begin, int t_count,
                                                                         int env_major, float scale,
                                                                         __nv_bfloat16* __restrict__ out) {
  const int nblk = HW >> 4;
  const long long total = (long long)t_count * B * nblk;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gstride) {
    const int blk = (int)(i % nblk);
    const long long r = i / nblk;
    int t, b;
    if (env_major) {
      b = (int)(r / t_count), t = (int)(r - (long long)b * t_count);
    } else {
      t = (int)(r / B), b = (int)(r - (long long)t * B);
    }
    t += t_begin;
    const int age = ages[(long long)t * B + b];
    uint32_t w[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int plane = t + 3 - min(3 - j, age);
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(planes + ((long long)plane * B + b) * HW) + blk);
      w[j][0] = v.x, w[j][1] = v.y, w[j][2] = v.z, w[j][3] = v.w;
    }
    uint4* dst = reinterpret_cast<uint4*>(out + (r * HW + (blk << 4)) * 4);
#pragma unroll
    for (int px = 0; px < 16; px += 2) {     // two pixels (8 bf16 = 16 bytes) per store
      uint32_t pk[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = px + h;
        float c[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) c[j] = (float)((w[j][p >> 2] >> ((p & 3) * 8)) & 0xffu) * scale;
        __nv_bfloat162 lo = __floats2bfloat162_rn(c[0], c[1]), hi = __floats2bfloat162_rn(c[2], c[3]);
        pk[h * 2] = *reinterpret_cast<uint32_t*>(&lo), pk[h * 2 + 1] = *reinterpret_cast<uint32_t*>(&hi);
      }
      dst[px >> 1] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// Gather fused with conv1's space-to-depth transform.  The first conv of the Atari models is 8x8 / stride 4 /
// pad 1 on 84x84x4 (benchmark/torch/a2c/atari_model.py:26-27); it never reads the last image row/column, so it
// equals a 2x2 / stride 1 conv on a 21x21 grid of 4x4 pixel blocks with 64 channels:
//   out[sample, Y, X, (dy*4+dx)*4 + c] = scale * frame_c[4Y+dy-1][4X+dx-1]      (zero outside the image)
// bf16, 128 contiguous bytes per block -> tensor-core friendly NHWC input.  One CTA walks whole samples: the four
// source frames (4 x 7056 B) are staged in shared memory with coalesced 16-byte cp.async (the first version issued
// eight 4-byte global loads per thread and sat in long-scoreboard stalls, ncu round 1), then one thread = one
// (Y,X,dy) item = 32 output bytes.  u8 -> float goes through the 2^23 magic number on the FMA pipe:
// fma(as_float(0x4B000000 | byte), scale, -2^23 * scale) == float(byte) * scale bit for bit (one rounding).
// U8OUT (out_dtype 4): the same layout with the bytes left as they are, 64 B per block — half the traffic; the
// conv1 kernels widen it to bf16 in shared memory (rl_conv2d_s1_u8in_bf16_{fwd,wgrad}, u8win.cuh).
template <bool U8OUT>
__global__ void __launch_bounds__(256) obs_stack_gather_s2d_kernel(const uint8_t* __restrict__ planes,
                                                                   const uint8_t* __restrict__ ages, int B, int t_begin,
                                                                   int t_count, int env_major, float scale,
                                                                   void* __restrict__ out_v) {
  constexpr int W = 84, G = 21, ITEMS = G * G * 4, FRAME = W * W, CHUNKS = FRAME / 16;     // 7056 B = 441 x 16
  __shared__ __align__(16) uint8_t sfr[4][FRAME];
  pdl_wait();            // chain kernel (launch_chain): the frames / ages come from the env-step kernel before it
  pdl_trigger();
  const long long nsamples = (long long)t_count * B;
  const float bias = -8388608.0f * scale;
  for (long long r = blockIdx.x; r < nsamples; r += gridDim.x) {      // sample index in output order
    int t, b;
    if (env_major) {
      b = (int)(r / t_count), t = (int)(r - (long long)b * t_count);
    } else {
      t = (int)(r / B), b = (int)(r - (long long)t * B);
    }
    t += t_begin;
    const int age = ages ? ages[(long long)t * B + b] : 0;
    for (int i = threadIdx.x; i < 4 * CHUNKS; i += 256) {
      const int c = i / CHUNKS, k = i - c * CHUNKS;
      // ages == NULL: the source is an already stacked observation tensor [n, 4, 84, 84] (host contract path)
      const long long img = ages ? ((long long)(t + 3 - min(3 - c, age)) * B + b) : (r * 4 + c);
      cp_async16(&sfr[c][k * 16], planes + img * FRAME + k * 16);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    __nv_bfloat16* dst_sample = reinterpret_cast<__nv_bfloat16*>(out_v) + r * (long long)(G * G * 64);
    uint8_t* dst_sample_u8 = reinterpret_cast<uint8_t*>(out_v) + r * (long long)(G * G * 64);
    for (int j = threadIdx.x; j < ITEMS; j += 256) {
      const int dy = j & 3, p = j >> 2, Y = p / G, X = p - Y * G;
      const int y = 4 * Y + dy - 1;
      uint32_t px[4];                           // per channel: the 4 bytes at x = 4X-1 .. 4X+2
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        px[c] = 0u;
        if (y >= 0) {
          const uint32_t* row = reinterpret_cast<const uint32_t*>(&sfr[c][0]) + y * (W / 4);
          const uint32_t w0 = X > 0 ? row[X - 1] : 0u;             // bytes 4X-4 .. 4X-1
          const uint32_t w1 = row[X];                              // bytes 4X   .. 4X+3
          px[c] = (w0 >> 24) | (w1 << 8);
        }
      }
      if (U8OUT) {
        // 4x4 byte transpose: word dx = {px[0].b[dx], px[1].b[dx], px[2].b[dx], px[3].b[dx]}
        uint32_t wd[4];
#pragma unroll
        for (int dx = 0; dx < 4; ++dx) {
          const uint32_t sel = (uint32_t)dx | ((uint32_t)(4 + dx) << 4);
          wd[dx] = __byte_perm(__byte_perm(px[0], px[1], sel), __byte_perm(px[2], px[3], sel), 0x5410u);
        }
        *reinterpret_cast<uint4*>(dst_sample_u8 + p * 64 + dy * 16) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        continue;
      }
      uint32_t pk[8];
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          v[c] = fmaf(__uint_as_float(__byte_perm(px[c], 0x4B000000u, 0x7540u | dx)), scale, bias);
        __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
        pk[dx * 2] = *reinterpret_cast<uint32_t*>(&lo), pk[dx * 2 + 1] = *reinterpret_cast<uint32_t*>(&hi);
      }
      uint4* dst = reinterpret_cast<uint4*>(dst_sample + p * 64 + dy * 16);
      dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    __syncthreads();                            // the next sample overwrites the staged frames
  }
}

// ---------------------------------------------------------------------------
// Env step t fused with the observation gather of step t+1 (uint8 space-to-depth form): the per-env-step chain of the
// actor loses one launch, and the new frame is used from shared memory instead of being read back from HBM.
//   CTAs [0, nscalar)   : the per-env scalar part of atari_synth_step_kernel, unchanged (reward, done, age, action
//                         sampling from the logits, episode statistics) — 256 envs per CTA;
//   CTAs [nscalar, ...) : one env at a time (grid-stride): generate frame (step+1) into shared memory AND the frame
//                         ring (the ring stays the system of record: later steps' stacks, the learner-side and host
//                         paths read it), recompute this env's done flag (a pure function of (env, step)) to get
//                         age(t+1), fetch the up to three older frames of the stack from the ring with cp.async, and
//                         write obs(t+1) as [21,21,64] uint8 blocks — bit-identical to rl_env_atari_synth_step
//                         followed by rl_obs_stack_gather(out_dtype 4).
// ---------------------------------------------------------------------------
struct AtariStepGatherArgs {
  AtariStepArgs s;
  const uint8_t* planes;     // ring base [P, B, 7056]; s.frame_out = planes[t + 4]
  int t;                     // row of this step: the older frames of obs(t+1) are planes[t + 4 - k], k = 1..3
  uint8_t* obs_next;         // [B, 21, 21, 64] uint8
  int nscalar;
};

__global__ void __launch_bounds__(256) atari_step_gather_kernel(AtariStepGatherArgs a) {
  constexpr int W = 84, G = 21, ITEMS = G * G * 4, FRAME = W * W, CHUNKS = FRAME / 16;
  __shared__ __align__(16) uint8_t sfr[4][FRAME];
  pdl_wait();            // chain kernel (launch_chain)
  pdl_trigger();
  AtariStepArgs& p = a.s;
  if (p.step_dev) p.step = *p.step_dev;
  if ((int)blockIdx.x < a.nscalar) {
    const int b = blockIdx.x * 256 + threadIdx.x;
    const bool valid = b < p.B;
    float reward = 0.f;
    bool done = false;
    if (valid) {
      const uint32_t env = p.env_offset + (uint32_t)b;
      const uint4 x = philox4x32_10(env, p.step, 0u, STREAM_REWDONE, p.k0, p.k1);
      reward = (float)(x.x & 1u);
      done = x.y < p.done_thr;
      p.reward_out[b] = reward;
      p.done_out[b] = done ? 1 : 0;
      const int age = p.age_in[b];
      p.age_out[b] = done ? 0 : (uint8_t)min(age + 1, 3);
      if (p.logits) {
        const uint4 ua = philox4x32_10(env, p.step, 0u, STREAM_ACTION, p.k0, p.k1);
        p.actions_out[b] = sample_categorical_exact(p.logits + (long long)b * p.A, p.A, u01_24(ua.x));
      }
    }
    episode_update(p.st, b, valid, reward, done);
    return;
  }
  const uint32_t n = p.step + 1u;
  for (int b = blockIdx.x - a.nscalar; b < p.B; b += gridDim.x - a.nscalar) {
    const uint32_t env = p.env_offset + (uint32_t)b;
    // age of obs(t+1): same arithmetic as the scalar part (done is a pure function of (env, step))
    const uint4 x = philox4x32_10(env, p.step, 0u, STREAM_REWDONE, p.k0, p.k1);
    const int age = (x.y < p.done_thr) ? 0 : min((int)p.age_in[b] + 1, 3);
    // newest frame: generate once, keep in shared memory (channel 3), write to the ring
    uint4* ring = reinterpret_cast<uint4*>(p.frame_out + (long long)b * FRAME);
    for (int k = threadIdx.x; k < CHUNKS; k += 256) {
      const uint4 v = frame_block(env, n, (uint32_t)k, p.k0, p.k1);
      *reinterpret_cast<uint4*>(&sfr[3][k * 16]) = v;
      ring[k] = v;
    }
    // older channels c = 0..2: plane[t + 4 - min(3 - c, age)]; min(..) == 0 is the new frame itself
    for (int i = threadIdx.x; i < 3 * CHUNKS; i += 256) {
      const int c = i / CHUNKS, k = i - c * CHUNKS;
      const int back = min(3 - c, age);
      if (back > 0)
        cp_async16(&sfr[c][k * 16], a.planes + ((long long)(a.t + 4 - back) * p.B + b) * FRAME + k * 16);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (age < 3) {                                   // channels that repeat the new frame (after a reset)
      for (int i = threadIdx.x; i < 3 * CHUNKS; i += 256) {
        const int c = i / CHUNKS, k = i - c * CHUNKS;
        if (min(3 - c, age) == 0) *reinterpret_cast<uint4*>(&sfr[c][k * 16]) = *reinterpret_cast<const uint4*>(&sfr[3][k * 16]);
      }
      __syncthreads();
    }
    uint8_t* dst_sample = a.obs_next + (long long)b * (G * G * 64);
    for (int j = threadIdx.x; j < ITEMS; j += 256) {
      const int dy = j & 3, pq = j >> 2, Y = pq / G, X = pq - Y * G;
      const int y = 4 * Y + dy - 1;
      uint32_t px[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        px[c] = 0u;
        if (y >= 0) {
          const uint32_t* row = reinterpret_cast<const uint32_t*>(&sfr[c][0]) + y * (W / 4);
          const uint32_t w0 = X > 0 ? row[X - 1] : 0u;
          const uint32_t w1 = row[X];
          px[c] = (w0 >> 24) | (w1 << 8);
        }
      }
      uint32_t wd[4];
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        const uint32_t sel = (uint32_t)dx | ((uint32_t)(4 + dx) << 4);
        wd[dx] = __byte_perm(__byte_perm(px[0], px[1], sel), __byte_perm(px[2], px[3], sel), 0x5410u);
      }
      *reinterpret_cast<uint4*>(dst_sample + pq * 64 + dy * 16) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
    __syncthreads();                                 // the next env overwrites the staged frames
  }
}

// ---------------------------------------------------------------------------
// MuJoCo-shaped synthetic env (obs N(0,1)^D) and CartPole physics: one lane per env.
// ---------------------------------------------------------------------------
struct VecStepArgs {
  float* obs_out;           // [B, D]
  float* reward_out;
  uint8_t* done_out;
  float* state;             // CartPole: [B,4] in/out
  const void* actions;      // CartPole: [B] int32
  EpisodeStats st;
  int B, D, max_steps;
  uint32_t k0, k1, step, env_offset, done_thr;
  int reset;
};

__global__ void __launch_bounds__(128) mujoco_synth_step_kernel(const VecStepArgs p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = b < p.B;
  const uint32_t env = p.env_offset + (uint32_t)b;
  float reward = 0.f;
  bool done = false;
  if (valid && !p.reset) {
    const uint4 x = philox4x32_10(env, p.step, 0u, STREAM_REWDONE, p.k0, p.k1);
    reward = (float)(x.x & 1u);
    done = x.y < p.done_thr;
    if (p.max_steps > 0 && p.st.ep_len[b] + 1 >= p.max_steps) done = true;
    p.reward_out[b] = reward;
    p.done_out[b] = done ? 1 : 0;
  }
  if (p.reset) {
    if (valid) p.st.ep_ret[b] = 0.f, p.st.ep_len[b] = 0;
  } else {
    episode_update(p.st, b, valid, reward, done);
  }
  if (valid) {
    const uint32_t n = p.reset ? p.step : p.step + 1u;
    for (int blk = 0; blk * 4 < p.D; ++blk) {
      float z[4];
      gauss_block(env, n, (uint32_t)blk, STREAM_OBS, p.k0, p.k1, z);
      for (int k = 0; k < 4 && blk * 4 + k < p.D; ++k) p.obs_out[(long long)b * p.D + blk * 4 + k] = z[k];
    }
  }
}

__global__ void __launch_bounds__(128) cartpole_step_kernel(const VecStepArgs p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = b < p.B;
  const uint32_t env = p.env_offset + (uint32_t)b;
  float reward = 0.f;
  bool done = false;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid && !p.reset) {
    s = reinterpret_cast<const float4*>(p.state)[b];
    const int a = reinterpret_cast<const int*>(p.actions)[b];
    done = cartpole_physics(s, a);
    reward = 1.0f;
    if (p.max_steps > 0 && p.st.ep_len[b] + 1 >= p.max_steps) done = true;
    p.reward_out[b] = reward;
    p.done_out[b] = done ? 1 : 0;
  }
  if (p.reset) {
    if (valid) p.st.ep_ret[b] = 0.f, p.st.ep_len[b] = 0;
  } else {
    episode_update(p.st, b, valid, reward, done);
  }
  if (valid) {
    if (p.reset || done) {
      s = cartpole_reset_state(env, p.reset ? p.step : p.step + 1u, p.k0, p.k1);
    }
    reinterpret_cast<float4*>(p.state)[b] = s;
    reinterpret_cast<float4*>(p.obs_out)[b] = s;
  }
}

// ---------------------------------------------------------------------------
// VecNormalizeEnv as a stand-alone kernel (one lane per env); the fused rollout applies the same device functions.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) vecnormalize_step_kernel(const VecNormState v, const float* __restrict__ obs_in,
                                                                const float* __restrict__ term_obs,
                                                                float* __restrict__ obs_out, float* __restrict__ reward,
                                                                const uint8_t* __restrict__ done, int B, int D,
                                                                int reward_step) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (reward_step) {
    const bool dn = done[b] != 0;
    if (term_obs && dn) vecnorm_obs_absorb(v, b, D, term_obs + (size_t)b * D, 1);
    reward[b] = vecnorm_reward(v, b, reward[b], dn);
  }
  float* o = obs_out + (size_t)b * D;
  if (obs_out != obs_in)
    for (int d = 0; d < D; ++d) o[d] = obs_in[(size_t)b * D + d];
  vecnorm_obs_filter(v, b, D, o, 1);
}

// ---------------------------------------------------------------------------
// K7 standalone samplers.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) sample_categorical_kernel(const float* __restrict__ logits, int N, int A,
                                                                 uint32_t k0, uint32_t k1, uint32_t step,
                                                                 uint32_t env_offset, int* __restrict__ actions,
                                                                 float* __restrict__ logp_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= N) return;
  const uint4 ua = philox4x32_10(env_offset + (uint32_t)b, step, 0u, STREAM_ACTION, k0, k1);
  const float* row = logits + (long long)b * A;
  const int a = sample_categorical_exact(row, A, u01_24(ua.x));
  actions[b] = a;
  if (logp_out) {
    float m = row[0];
    for (int j = 1; j < A; ++j) m = fmaxf(m, row[j]);
    float S = 0.f;
    for (int j = 0; j < A; ++j) S += expf(row[j] - m);
    logp_out[b] = row[a] - m - logf(S);
  }
}

// action = mean + exp(logstd) * z ; logp = sum_d [-(a-mu)^2/(2 s^2) - logstd - 0.5 log 2pi]  (ppo.py:164-169)
__global__ void __launch_bounds__(128) sample_gaussian_kernel(const float* __restrict__ mean,
                                                              const float* __restrict__ logstd, int N, int D,
                                                              uint32_t k0, uint32_t k1, uint32_t step,
                                                              uint32_t env_offset, float* __restrict__ action,
                                                              float* __restrict__ logp_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= N) return;
  float lp = 0.f;
  for (int blk = 0; blk * 4 < D; ++blk) {
    float z[4];
    gauss_block(env_offset + (uint32_t)b, step, (uint32_t)blk, STREAM_GAUSS, k0, k1, z);
    for (int k = 0; k < 4 && blk * 4 + k < D; ++k) {
      const int d = blk * 4 + k;
      const float ls = logstd[d], sd = expf(ls);
      const float a = fmaf(sd, z[k], mean[(long long)b * D + d]);
      action[(long long)b * D + d] = a;
      lp += -0.5f * z[k] * z[k] - ls - 0.9189385332046727f;
    }
  }
  if (logp_out) logp_out[b] = lp;
}

static EpisodeStats make_stats(float* ep_ret, int* ep_len, float* totals, float* ring_ret, int* ring_len,
                               unsigned* ring_head, int ring_cap) {
  return make_episode_stats(ep_ret, ep_len, totals, ring_ret, ring_len, ring_head, ring_cap);
}

static uint32_t prob_thr(float p) { return prob_threshold(p); }

}  // namespace rl

using namespace rl;

extern "C" int rl_env_atari_synth_step(uint8_t* frame_out, float* reward_out, uint8_t* done_out,
                                       const uint8_t* age_in, uint8_t* age_out, const float* logits, int A,
                                       int32_t* actions_out, float* ep_ret, int32_t* ep_len, float* totals,
                                       float* ring_ret, int32_t* ring_len, uint32_t* ring_head, int ring_cap, int B,
                                       int HW, uint64_t seed, uint32_t step, const uint32_t* step_dev,
                                       uint32_t env_offset, float p_done, int reset, rl_stream_t stream) {
  RL_CHECK_ARG(frame_out && age_out && ep_ret && ep_len && totals, "atari_synth_step: null pointer");
  RL_CHECK_ARG(reset || (reward_out && done_out && age_in), "atari_synth_step: null pointer");
  RL_CHECK_ARG(B > 0 && HW > 0 && HW % 16 == 0, "atari_synth_step: B=%d HW=%d (HW must be a multiple of 16)", B, HW);
  RL_CHECK_ARG((long long)B * (HW / 16) < (1LL << 31), "atari_synth_step: B*HW/16 must be < 2^31");
  RL_CHECK_ARG(aligned16(frame_out), "atari_synth_step: frame plane must be 16-byte aligned");
  RL_CHECK_ARG(!logits || (A >= 1 && actions_out), "atari_synth_step: logits given but A=%d / actions_out null", A);
  AtariStepArgs a;
  a.frame_out = frame_out, a.reward_out = reward_out, a.done_out = done_out, a.age_in = age_in, a.age_out = age_out;
  a.logits = logits, a.actions_out = actions_out, a.A = A;
  a.st = make_stats(ep_ret, ep_len, totals, ring_ret, ring_len, ring_head, ring_cap);
  a.B = B, a.HW = HW, a.k0 = (uint32_t)seed, a.k1 = (uint32_t)(seed >> 32), a.step = step, a.env_offset = env_offset;
  a.done_thr = prob_thr(p_done), a.reset = reset, a.step_dev = step_dev;
  const long long total = (long long)B * (HW / 16);
  long long blocks = (total + 255) / 256;
  const long long cap = 148LL * 8 * 4;      // 4 waves of 8 CTAs/SM, grid-stride beyond
  if (blocks > cap) blocks = cap;
  if (blocks < (B + 255) / 256) blocks = (B + 255) / 256;
  launch_chain(atari_synth_step_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, a);
  RL_CHECK_LAUNCH("rl_env_atari_synth_step");
  return RL_OK;
}

extern "C" int rl_env_atari_synth_step_gather(uint8_t* planes, int t, float* reward_out, uint8_t* done_out,
                                              const uint8_t* age_in, uint8_t* age_out, const float* logits, int A,
                                              int32_t* actions_out, float* ep_ret, int32_t* ep_len, float* totals,
                                              float* ring_ret, int32_t* ring_len, uint32_t* ring_head, int ring_cap,
                                              int B, uint64_t seed, uint32_t step, const uint32_t* step_dev,
                                              uint32_t env_offset, float p_done, uint8_t* obs_next, rl_stream_t stream) {
  RL_CHECK_ARG(planes && reward_out && done_out && age_in && age_out && ep_ret && ep_len && totals && obs_next && t >= 0,
               "atari_synth_step_gather: null pointer");
  RL_CHECK_ARG(B > 0 && aligned16(planes) && aligned16(obs_next), "atari_synth_step_gather: bad shape / alignment");
  RL_CHECK_ARG(!logits || (A >= 1 && actions_out), "atari_synth_step_gather: logits given but A=%d / actions_out null", A);
  AtariStepGatherArgs g;
  AtariStepArgs& a = g.s;
  a.frame_out = planes + (size_t)(t + 4) * B * 7056, a.reward_out = reward_out, a.done_out = done_out;
  a.age_in = age_in, a.age_out = age_out, a.logits = logits, a.actions_out = actions_out, a.A = A;
  a.st = make_stats(ep_ret, ep_len, totals, ring_ret, ring_len, ring_head, ring_cap);
  a.B = B, a.HW = 7056, a.k0 = (uint32_t)seed, a.k1 = (uint32_t)(seed >> 32), a.step = step, a.env_offset = env_offset;
  a.done_thr = prob_thr(p_done), a.reset = 0, a.step_dev = step_dev;
  g.planes = planes, g.t = t, g.obs_next = obs_next, g.nscalar = (B + 255) / 256;
  long long env_ctas = B;
  if (env_ctas > 148LL * 7) env_ctas = 148LL * 7;          // 28 KB of staged frames per CTA: 7 CTAs per SM
  launch_chain(atari_step_gather_kernel, dim3((unsigned)(g.nscalar + env_ctas)), dim3(256), 0, (cudaStream_t)stream, g);
  RL_CHECK_LAUNCH("rl_env_atari_synth_step_gather");
  return RL_OK;
}

extern "C" int rl_obs_stack_gather(const uint8_t* planes, const uint8_t* ages, int B, int HW, int t_begin, int t_count,
                                   int out_layout, int out_dtype, float scale, void* out, rl_stream_t stream) {
  RL_CHECK_ARG(planes && out && (ages || out_dtype == 3 || out_dtype == 4), "obs_stack_gather: null pointer");
  RL_CHECK_ARG(B > 0 && HW > 0 && HW % 16 == 0 && t_count > 0 && t_begin >= 0, "obs_stack_gather: bad shape");
  RL_CHECK_ARG(aligned16(planes) && aligned16(out), "obs_stack_gather: 16-byte alignment required");
  const long long total = (long long)t_count * B * 4 * (HW / 16);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  const int em = out_layout == RL_LAYOUT_ENV_MAJOR;
  if (out_dtype == 0) {
    obs_stack_gather_kernel<uint8_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        planes, ages, B, HW, t_begin, t_count, em, scale, (uint8_t*)out);
  } else if (out_dtype == 1) {
    obs_stack_gather_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        planes, ages, B, HW, t_begin, t_count, em, scale, (float*)out);
  } else if (out_dtype == 2) {
    const long long tot2 = (long long)t_count * B * (HW / 16);
    long long b2 = (tot2 + 255) / 256;
    if (b2 > 148LL * 32) b2 = 148LL * 32;
    obs_stack_gather_nhwc_bf16_kernel<<<(unsigned)b2, 256, 0, (cudaStream_t)stream>>>(
        planes, ages, B, HW, t_begin, t_count, em, scale, (__nv_bfloat16*)out);
  } else if (out_dtype == 3 || out_dtype == 4) {
    RL_CHECK_ARG(HW == 84 * 84, "obs_stack_gather: the space-to-depth layout is defined for 84x84 frames");
    RL_CHECK_ARG(aligned16(planes) && aligned16(out), "obs_stack_gather: 16-byte alignment required (cp.async staging)");
    long long b3 = (long long)t_count * B;              // one CTA per sample, grid-stride beyond 8 CTAs per SM
    if (b3 > 148LL * 7) b3 = 148LL * 7;                 // 28 KB of staged frames per CTA: 7 CTAs per SM
    if (out_dtype == 3)
      launch_chain(obs_stack_gather_s2d_kernel<false>, dim3((unsigned)b3), dim3(256), 0, (cudaStream_t)stream, planes, ages,
                   B, t_begin, t_count, em, scale, out);
    else
      launch_chain(obs_stack_gather_s2d_kernel<true>, dim3((unsigned)b3), dim3(256), 0, (cudaStream_t)stream, planes, ages,
                   B, t_begin, t_count, em, scale, out);
  } else {
    set_error("obs_stack_gather: out_dtype %d unsupported (0=u8, 1=f32, 2=bf16 NHWC, 3=bf16 space-to-depth, 4=u8 space-to-depth)",
              out_dtype);
    return RL_ERR_UNSUPPORTED;
  }
  RL_CHECK_LAUNCH("rl_obs_stack_gather");
  return RL_OK;
}

static int vec_step(bool cartpole, float* obs_out, float* reward_out, uint8_t* done_out, float* state,
                    const void* actions, float* ep_ret, int32_t* ep_len, float* totals, float* ring_ret,
                    int32_t* ring_len, uint32_t* ring_head, int ring_cap, int B, int D, int max_steps, uint64_t seed,
                    uint32_t step, uint32_t env_offset, float p_done, int reset, rl_stream_t stream) {
  VecStepArgs a;
  a.obs_out = obs_out, a.reward_out = reward_out, a.done_out = done_out, a.state = state, a.actions = actions;
  a.st = make_stats(ep_ret, ep_len, totals, ring_ret, ring_len, ring_head, ring_cap);
  a.B = B, a.D = D, a.max_steps = max_steps, a.k0 = (uint32_t)seed, a.k1 = (uint32_t)(seed >> 32);
  a.step = step, a.env_offset = env_offset, a.done_thr = prob_thr(p_done), a.reset = reset;
  const int blocks = (B + 127) / 128;
  if (cartpole)
    cartpole_step_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(a);
  else
    mujoco_synth_step_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(a);
  return 0;
}

extern "C" int rl_env_mujoco_synth_step(float* obs_out, float* reward_out, uint8_t* done_out, float* ep_ret,
                                        int32_t* ep_len, float* totals, float* ring_ret, int32_t* ring_len,
                                        uint32_t* ring_head, int ring_cap, int B, int obs_dim, int max_episode_steps,
                                        uint64_t seed, uint32_t step, uint32_t env_offset, float p_done, int reset,
                                        rl_stream_t stream) {
  RL_CHECK_ARG(obs_out && ep_ret && ep_len && totals && (reset || (reward_out && done_out)),
               "mujoco_synth_step: null pointer");
  RL_CHECK_ARG(B > 0 && obs_dim > 0, "mujoco_synth_step: bad shape");
  vec_step(false, obs_out, reward_out, done_out, nullptr, nullptr, ep_ret, ep_len, totals, ring_ret, ring_len,
           ring_head, ring_cap, B, obs_dim, max_episode_steps, seed, step, env_offset, p_done, reset, stream);
  RL_CHECK_LAUNCH("rl_env_mujoco_synth_step");
  return RL_OK;
}

extern "C" int rl_env_cartpole_step(float* state, float* obs_out, float* reward_out, uint8_t* done_out,
                                    const int32_t* actions, float* ep_ret, int32_t* ep_len, float* totals,
                                    float* ring_ret, int32_t* ring_len, uint32_t* ring_head, int ring_cap, int B,
                                    int max_episode_steps, uint64_t seed, uint32_t step, uint32_t env_offset, int reset,
                                    rl_stream_t stream) {
  RL_CHECK_ARG(state && obs_out && ep_ret && ep_len && totals && (reset || (reward_out && done_out && actions)),
               "cartpole_step: null pointer");
  RL_CHECK_ARG(B > 0 && aligned16(state) && aligned16(obs_out), "cartpole_step: bad shape / alignment");
  vec_step(true, obs_out, reward_out, done_out, state, actions, ep_ret, ep_len, totals, ring_ret, ring_len, ring_head,
           ring_cap, B, 4, max_episode_steps, seed, step, env_offset, 0.f, reset, stream);
  RL_CHECK_LAUNCH("rl_env_cartpole_step");
  return RL_OK;
}

extern "C" int rl_vecnormalize_step(const float* obs_in, const float* term_obs, float* obs_out, float* reward,
                                    const uint8_t* done, double* ob_mean, double* ob_var, double* ob_count, double* ret,
                                    double* ret_mean, double* ret_var, double* ret_count, int B, int D, int update,
                                    int norm_ob, int norm_ret, int reward_step, double clipob, double cliprew,
                                    double gamma, double eps, rl_stream_t stream) {
  RL_CHECK_ARG(obs_in && obs_out && ob_mean && ob_var && ob_count && ret && ret_mean && ret_var && ret_count,
               "vecnormalize_step: null pointer");
  RL_CHECK_ARG(!reward_step || (reward && done), "vecnormalize_step: a reward step needs reward and done");
  RL_CHECK_ARG(B > 0 && D > 0, "vecnormalize_step: bad shape");
  VecNormState v;
  v.ob_mean = ob_mean, v.ob_var = ob_var, v.ob_count = ob_count, v.ret = ret, v.ret_mean = ret_mean;
  v.ret_var = ret_var, v.ret_count = ret_count, v.clipob = clipob, v.cliprew = cliprew, v.gamma = gamma, v.eps = eps;
  v.update = update, v.norm_ob = norm_ob, v.norm_ret = norm_ret;
  vecnormalize_step_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(v, obs_in, term_obs, obs_out, reward, done,
                                                                             B, D, reward_step);
  RL_CHECK_LAUNCH("rl_vecnormalize_step");
  return RL_OK;
}

extern "C" int rl_sample_categorical(const float* logits, int N, int A, uint64_t seed, uint32_t step,
                                     uint32_t env_offset, int32_t* actions, float* logp_out, rl_stream_t stream) {
  RL_CHECK_ARG(logits && actions && N > 0 && A > 0, "sample_categorical: bad argument");
  sample_categorical_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      logits, N, A, (uint32_t)seed, (uint32_t)(seed >> 32), step, env_offset, actions, logp_out);
  RL_CHECK_LAUNCH("rl_sample_categorical");
  return RL_OK;
}

extern "C" int rl_sample_gaussian(const float* mean, const float* logstd, int N, int D, uint64_t seed, uint32_t step,
                                  uint32_t env_offset, float* action, float* logp_out, rl_stream_t stream) {
  RL_CHECK_ARG(mean && logstd && action && N > 0 && D > 0, "sample_gaussian: bad argument");
  sample_gaussian_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      mean, logstd, N, D, (uint32_t)seed, (uint32_t)(seed >> 32), step, env_offset, action, logp_out);
  RL_CHECK_LAUNCH("rl_sample_gaussian");
  return RL_OK;
}
