// K4 (replay part) — HBM-resident replay: proportional prioritised sampling on a device
// sum-tree, Atari frame-context gather, generic row gather.
//
// Reference being replaced (PaddlePaddle/PARL):
//   benchmark/fluid/Prioritized_DQN/proportional_per.py:18-70   SumTree (array 2*cap-1, fp64 Python floats)
//   benchmark/fluid/Prioritized_DQN/proportional_per.py:73-157  ProportionalPER store/update/sample
//   benchmark/torch/dqn/replay_memory.py:59-113                 frame-context sampling with
//                                                               episode-boundary zeroing
//   parl/utils/replay_memory.py:51-95                           uniform gather by index
// The tree lives in HBM as double[2*cap-1] with the reference's heap indexing
// (leaf i at cap-1+i).  A batch of leaf writes is followed by a level-synchronous repair
// (parent = left + right) inside one CTA; sampling is one root-to-leaf descent per lane.
#include "common.cuh"
#include "philox.cuh"

namespace rl {

constexpr int kPT = 1024;

// state[0] = _min (running minimum of every priority ever written, proportional_per.py:24,34,42)
// state[1] = _max_priority (proportional_per.py:88,118)
struct PerWriteArgs {
  double* tree;
  double* state;
  const int* tree_idx;      // update: explicit leaf tree indices ; store: NULL
  const float* delta;       // |td| (update) or per-item delta (store; <=0 / NULL -> max_priority)
  int capacity, n, write_pos;
  double alpha, eps;
  int is_update;
};

__global__ void __launch_bounds__(kPT) per_write_kernel(const PerWriteArgs p) {
  __shared__ double s_min[kPT / 32], s_max[kPT / 32];
  const int tid = threadIdx.x;
  const double maxp0 = p.state[1];
  double lmin = 1e300, lmax = -1e300;
  // 1) leaf writes.  Duplicated indices: the LAST occurrence wins, as in the reference's sequential loop.
  for (int i = tid; i < p.n; i += kPT) {
    int idx;
    double pr;
    if (p.is_update) {
      idx = p.tree_idx[i];
      pr = pow((double)p.delta[i] + p.eps, p.alpha);                    // proportional_per.py:114-115
      bool later_dup = false;
      for (int j = i + 1; j < p.n; ++j) later_dup |= (p.tree_idx[j] == idx);
      if (!later_dup) p.tree[idx] = pr;
      lmax = fmax(lmax, pr);
    } else {
      idx = (p.write_pos + i) % p.capacity + p.capacity - 1;            // :30-31
      const double d = (p.delta && p.delta[i] > 0.f) ? (double)p.delta[i] : maxp0;   // :107-108
      pr = pow(d + p.eps, p.alpha);                                     // :110
      if (i + p.capacity >= p.n) p.tree[idx] = pr;                      // ring wrap: later item wins
    }
    lmin = fmin(lmin, pr);
  }
  __syncthreads();
  // 2) level-synchronous repair of the ancestors (see header comment)
  int depth = 0;
  for (int c = 2 * p.capacity - 2; c > 0; c = (c - 1) >> 1) ++depth;
  for (int s = 0; s < depth; ++s) {
    for (int i = tid; i < p.n; i += kPT) {
      // the ancestor exactly s+1 hops above this item's leaf (if the leaf is that deep)
      int j = p.is_update ? p.tree_idx[i] : (p.write_pos + i) % p.capacity + p.capacity - 1;
      int hops = 0;
      while (j > 0 && hops <= s) j = (j - 1) >> 1, ++hops;
      if (hops == s + 1) {
        volatile double* t = p.tree;
        t[j] = t[2 * j + 1] + t[2 * j + 2];
      }
    }
    __syncthreads();
  }
  // 3) running min / max-priority scalars
  for (int o = 16; o > 0; o >>= 1) {
    lmin = fmin(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
    lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  }
  if ((tid & 31) == 0) s_min[tid >> 5] = lmin, s_max[tid >> 5] = lmax;
  __syncthreads();
  if (tid == 0) {
    double mn = p.state[0], mx = p.state[1];
    for (int w = 0; w < kPT / 32; ++w) mn = fmin(mn, s_min[w]), mx = fmax(mx, s_max[w]);
    p.state[0] = mn;
    if (p.is_update) p.state[1] = mx;                                   // store() never raises _max_priority
  }
}

__global__ void __launch_bounds__(128) per_sample_kernel(const double* __restrict__ tree, const double* __restrict__ state,
                                                         int capacity, int seg_num, const float* __restrict__ u,
                                                         uint32_t k0, uint32_t k1, uint32_t draw, double beta,
                                                         double size, int* __restrict__ tree_idx,
                                                         int* __restrict__ elem_idx, float* __restrict__ weights) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= seg_num) return;
  const double total = tree[0];
  const double seg = total / (double)seg_num;                           // proportional_per.py:138
  const double low = seg * (double)i, high = seg * (double)(i + 1);     // :139-140
  const float uf = u ? u[i] : u01_24(philox4x32_10((uint32_t)i, draw, 0u, STREAM_REPLAY, k0, k1).x);
  double value = low + (double)uf * (high - low);                       // np.random.uniform(low, high)
  const int ntree = 2 * capacity - 1;
  int parent = 0;
  while (true) {                                                        // retrieve(), :44-60
    const int left = 2 * parent + 1;
    if (left >= ntree) break;
    const double lv = tree[left];
    if (value <= lv) {
      parent = left;
    } else {
      value -= lv;
      parent = left + 1;
    }
  }
  const double pr = tree[parent];
  tree_idx[i] = parent;
  if (elem_idx) elem_idx[i] = parent - capacity + 1;
  const double prob = size * pr / total;                                // :154
  const double min_prob = size * state[0] / total;                      // :155
  weights[i] = (float)pow(prob / min_prob, -beta);                      // :156
}

// Atari frame-context gather (benchmark/torch/dqn/replay_memory.py:59-85): out[s] = context_len+1
// consecutive frames starting at idx[s]; frames at or before the latest episode end among the
// first context_len-1 slots are zeroed.  One 16-byte block per thread.
__global__ void __launch_bounds__(256) replay_gather_frames_kernel(const uint8_t* __restrict__ frames,
                                                                   const uint8_t* __restrict__ is_over,
                                                                   const int* __restrict__ idx, int n, int curr_size,
                                                                   int ctx, int HW, int lanes, int n_out,
                                                                   uint8_t* __restrict__ out) {
  // Ring of `lanes` interleaved transition streams (one per env, stepped in lock-step): stream position q of lane l
  // is row q*lanes + l, so the plane of one position is a dense [lanes, HW] block the env kernel writes in place.
  // A start index encodes (q0, l) as q0*lanes + l; lanes == 1 is the reference's single ring.
  const int nblk = HW >> 4;
  const long long total = (long long)n * n_out * nblk;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gstride) {
    const int blk = (int)(i % nblk);
    const long long r = i / nblk;
    const int j = (int)(r % n_out);
    const int s = (int)(r / n_out);
    const int start = idx[s];
    const int q0 = start / lanes, lane = start - q0 * lanes;
    int cut = -1;
    for (int k = ctx - 2; k >= 0; --k) {
      if (is_over[(long long)((q0 + k) % curr_size) * lanes + lane]) {
        cut = k;
        break;
      }
    }
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (j > cut)
      v = __ldg(reinterpret_cast<const uint4*>(frames + ((long long)((q0 + j) % curr_size) * lanes + lane) * HW) + blk);
    reinterpret_cast<uint4*>(out + r * HW)[blk] = v;
  }
}

// out[i, :] = src[idx[i], :] for rows of row_bytes (multiple of 4); generic uniform-replay gather.
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint32_t* __restrict__ src, const int* __restrict__ idx,
                                                          long long n, int row_words, uint32_t* __restrict__ out) {
  const long long total = n * row_words;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gstride) {
    const long long r = i / row_words;
    const int w = (int)(i - r * row_words);
    out[i] = __ldg(src + (long long)idx[r] * row_words + w);
  }
}

}  // namespace rl

using namespace rl;

static int per_write(double* tree, double* state, int capacity, const int* tree_idx, const float* delta, int n,
                     int write_pos, double alpha, double eps, int is_update, rl_stream_t stream, const char* name) {
  RL_CHECK_ARG(tree && state && capacity >= 2 && n >= 1, "%s: bad argument", name);
  PerWriteArgs a;
  a.tree = tree, a.state = state, a.tree_idx = tree_idx, a.delta = delta, a.capacity = capacity, a.n = n;
  a.write_pos = write_pos, a.alpha = alpha, a.eps = eps, a.is_update = is_update;
  per_write_kernel<<<1, kPT, 0, (cudaStream_t)stream>>>(a);
  return RL_OK;
}

extern "C" int rl_per_store(double* tree, double* state, int capacity, int write_pos, int n, const float* delta,
                            double alpha, double eps, rl_stream_t stream) {
  RL_CHECK_ARG(write_pos >= 0 && write_pos < capacity, "per_store: write_pos out of range");
  if (per_write(tree, state, capacity, nullptr, delta, n, write_pos, alpha, eps, 0, stream, "per_store")) return RL_ERR_BAD_ARG;
  RL_CHECK_LAUNCH("rl_per_store");
  return RL_OK;
}

extern "C" int rl_per_update(double* tree, double* state, int capacity, const int32_t* tree_idx,
                             const float* priorities, int n, double alpha, double eps, rl_stream_t stream) {
  RL_CHECK_ARG(tree_idx && priorities, "per_update: null pointer");
  if (per_write(tree, state, capacity, tree_idx, priorities, n, 0, alpha, eps, 1, stream, "per_update")) return RL_ERR_BAD_ARG;
  RL_CHECK_LAUNCH("rl_per_update");
  return RL_OK;
}

extern "C" int rl_per_sample(const double* tree, const double* state, int capacity, int seg_num, const float* u,
                             uint64_t seed, uint32_t draw, double beta, double size, int32_t* tree_idx,
                             int32_t* elem_idx, float* weights, rl_stream_t stream) {
  RL_CHECK_ARG(tree && state && tree_idx && weights && capacity >= 2 && seg_num >= 1, "per_sample: bad argument");
  per_sample_kernel<<<(seg_num + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      tree, state, capacity, seg_num, u, (uint32_t)seed, (uint32_t)(seed >> 32), draw, beta, size, tree_idx, elem_idx,
      weights);
  RL_CHECK_LAUNCH("rl_per_sample");
  return RL_OK;
}

extern "C" int rl_replay_gather_frames(const uint8_t* frames, const uint8_t* is_over, const int32_t* idx, int n,
                                       int curr_size, int context_len, int HW, int lanes, int n_out, uint8_t* out,
                                       rl_stream_t stream) {
  RL_CHECK_ARG(frames && is_over && idx && out, "replay_gather_frames: null pointer");
  RL_CHECK_ARG(n >= 1 && curr_size >= 1 && context_len >= 1 && HW % 16 == 0 && aligned16(frames) && aligned16(out),
               "replay_gather_frames: bad shape / alignment");
  RL_CHECK_ARG(lanes >= 1 && n_out >= 1 && n_out <= context_len + 1, "replay_gather_frames: lanes=%d n_out=%d", lanes, n_out);
  const long long total = (long long)n * n_out * (HW / 16);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  replay_gather_frames_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(frames, is_over, idx, n, curr_size,
                                                                                 context_len, HW, lanes, n_out, out);
  RL_CHECK_LAUNCH("rl_replay_gather_frames");
  return RL_OK;
}

extern "C" int rl_gather_rows(const void* src, const int32_t* idx, long long n, int row_bytes, void* out,
                              rl_stream_t stream) {
  RL_CHECK_ARG(src && idx && out && n >= 1 && row_bytes >= 4 && row_bytes % 4 == 0, "gather_rows: bad argument");
  const long long total = n * (row_bytes / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  gather_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)src, idx, n, row_bytes / 4,
                                                                        (uint32_t*)out);
  RL_CHECK_LAUNCH("rl_gather_rows");
  return RL_OK;
}
