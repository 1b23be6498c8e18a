/*
 * parl_b200 — C ABI of the B200 (sm_100a) actor-learner hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no
 * torch / C++ types.  Every entry point names the PaddlePaddle/PARL code whose
 * arithmetic it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - All data pointers are DEVICE pointers owned by the caller (e.g.
 *     torch.Tensor.data_ptr()); the library never allocates or frees
 *     user-visible memory.  `workspace` is caller-allocated scratch.
 *   - Work is enqueued on `stream` (a cudaStream_t passed as void*); no call
 *     synchronises the device or blocks the host.
 *   - Return value: RL_OK (0) or a negative RL_ERR_* code; rl_last_error()
 *     returns a thread-local message.  No exceptions cross the ABI.
 *   - "time-major" = [T, B, ...] contiguous; "env-major" = the reference's flat
 *     [B*T, ...] with index b*T + t (examples/IMPALA/actor.py:79-89).
 *   - Losses are written to device memory (`losses`), never read back here.
 */
#ifndef PARL_B200_H_
#define PARL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RL_OK 0
#define RL_ERR_BAD_ARG (-1)
#define RL_ERR_ALIGN (-2)
#define RL_ERR_CUDA (-3)
#define RL_ERR_UNSUPPORTED (-4)
#define RL_ERR_WORKSPACE (-5)

#define RL_LAYOUT_TIME_MAJOR 0
#define RL_LAYOUT_ENV_MAJOR 1

typedef void* rl_stream_t; /* cudaStream_t */

int rl_abi_version(void);
const char* rl_last_error(void);
/* Number of SMs / device name probe used by the host side for grid sizing. */
int rl_device_sm_count(int device);

/* Bytes of zero-initialised scratch the loss kernels need for a problem with
 * `n_cols` independent columns (B) — partial sums + the last-block ticket.
 * The caller zeroes it ONCE (cudaMemset / torch.zeros); kernels leave it zeroed. */
size_t rl_loss_workspace_bytes(int n_cols);

/* ------------------------------------------------------------------------
 * a1  V-trace returns.
 * Replaces parl/algorithms/paddle/impala/vtrace.py:36-139
 * (from_importance_weights).  All inputs [T,B] float32 time-major,
 * bootstrap [B]; outputs vs, pg_advantages [T,B].  clip thresholds < 0 mean
 * "None" (no clipping), matching vtrace.py:104-107,130-133.
 * ---------------------------------------------------------------------- */
int rl_vtrace_from_importance_weights(
    const float* behaviour_actions_log_probs, const float* target_actions_log_probs,
    const float* discounts, const float* rewards, const float* values,
    const float* bootstrap_value, int T, int B,
    float clip_rho_threshold, float clip_pg_rho_threshold,
    float* vs, float* pg_advantages, rl_stream_t stream);

/* ------------------------------------------------------------------------
 * a1+a2+a3  Fused IMPALA loss: log-softmax, entropy, KL, V-trace backward
 * scan, policy-gradient + value + entropy loss AND its gradient w.r.t. the
 * network outputs, one launch.
 * Replaces parl/algorithms/paddle/impala/impala.py:148-208 (post-network part
 * of IMPALA.learn), impala.py:25-79 (VTraceLoss) and vtrace.py:99-139.
 *
 *   target_logits, behaviour_logits : [T,B,A] f32   (T = sample_batch_steps,
 *                                      the LAST row is only the bootstrap row)
 *   actions  : [T,B] int32 (actions_i64=0) or int64 (actions_i64=1)
 *   rewards  : [T,B] f32 ; dones : [T,B] u8 (bool) ; values : [T,B] f32
 *   layout   : RL_LAYOUT_TIME_MAJOR or RL_LAYOUT_ENV_MAJOR (applies to all)
 *   losses   : [5] f32 device = {total, pi_loss, vf_loss, entropy, kl}
 *              (SUM reductions over the (T-1)*B kept rows; kl = MEAN over T*B)
 *   d_logits : [T,B,A] f32 = d total / d target_logits (row T-1 = 0)
 *   d_values : [T,B]   f32 = d total / d values        (row T-1 = 0)
 *   vs_out, pg_adv_out : optional [T-1,B] time-major f32 (may be NULL)
 * ---------------------------------------------------------------------- */
int rl_vtrace_loss_fwd_bwd(
    const float* target_logits, const float* behaviour_logits, const void* actions, int actions_i64,
    const float* rewards, const uint8_t* dones, const float* values,
    int T, int B, int A, int layout,
    float gamma, float clip_rho_threshold, float clip_pg_rho_threshold,
    float vf_loss_coeff, float entropy_coeff,
    float* losses, float* d_logits, float* d_values, float* vs_out, float* pg_adv_out,
    void* workspace, size_t workspace_bytes, rl_stream_t stream);

/* ------------------------------------------------------------------------
 * a10/a11  On-device vectorised actor pool (K5 env step + K7 sampling).
 * Replaces parl/env/vector_env.py:41-63 (VectorEnv.step, auto-reset),
 * the mock envs parl/tests/gym.py:117-207, FrameStack
 * (parl/env/atari_wrappers.py:270-307) and per-row action sampling
 * (examples/IMPALA/atari_agent.py:39-40, parl/algorithms/torch/ppo.py:164-177).
 * RNG: Philox4x32-10, counter (env_offset+b, step, block, stream), key = seed.
 *
 * Episode bookkeeping (all envs): ep_ret f32[B], ep_len i32[B] running values;
 * totals f32[4] = {#completed, sum return, sum length, -}; optional ring of the
 * most recent completed episodes (ring_ret f32[cap], ring_len i32[cap],
 * ring_head u32[1]) — the device analogue of MonitorEnv.next_episode_results()
 * (examples/IMPALA/actor.py:93-102).
 * ---------------------------------------------------------------------- */

/* Atari-shaped synthetic env.  Writes frame (step+1) of every env into
 * `frame_out` ([B,HW] uint8 plane of the caller's frame ring), reward/done of
 * step `step`, and age_out = done ? 0 : min(age_in+1, 3) (frames of the virtual
 * 4-stack that belong to the current episode).  If `logits` != NULL also samples
 * actions_out[b] ~ Categorical(logits[b,:A]) with the exact inverse-CDF rule.
 * reset=1: only emit frame `step` into frame_out and zero age/episode state. */
int rl_env_atari_synth_step(
    uint8_t* frame_out, float* reward_out, uint8_t* done_out,
    const uint8_t* age_in, uint8_t* age_out,
    const float* logits, int A, int32_t* actions_out,
    float* ep_ret, int32_t* ep_len, float* totals,
    float* ring_ret, int32_t* ring_len, uint32_t* ring_head, int ring_cap,
    int B, int HW, uint64_t seed, uint32_t step, uint32_t env_offset, float p_done,
    int reset, rl_stream_t stream);

/* Materialise observations from the frame ring: obs(t,b) channel j (0 = oldest)
 * = plane[t + 3 - min(3-j, age[t,b])].  planes [P,B,HW] u8, ages [>=t_begin+t_count, B] u8.
 * Output [t_count*B, 4, HW] in time-major or env-major sample order;
 * out_dtype 0 = uint8, 1 = float32 (value * scale). */
int rl_obs_stack_gather(
    const uint8_t* planes, const uint8_t* ages, int B, int HW, int t_begin, int t_count,
    int out_layout, int out_dtype, float scale, void* out, rl_stream_t stream);

/* MuJoCo-shaped synthetic env: obs ~ N(0,1)^obs_dim, reward in {0,1}, done ~ p_done
 * (optionally also at max_episode_steps; 0 = no limit).  On done the returned obs is
 * the reset obs (vector_env.py:56-57) — every obs is a fresh draw. */
int rl_env_mujoco_synth_step(
    float* obs_out, float* reward_out, uint8_t* done_out,
    float* ep_ret, int32_t* ep_len, float* totals,
    float* ring_ret, int32_t* ring_len, uint32_t* ring_head, int ring_cap,
    int B, int obs_dim, int max_episode_steps,
    uint64_t seed, uint32_t step, uint32_t env_offset, float p_done, int reset, rl_stream_t stream);

/* CartPole physics (gym classic-control constants, Euler, float32), auto-reset
 * with U(-0.05,0.05)^4.  state [B,4] in/out, obs_out [B,4], actions [B] int32. */
int rl_env_cartpole_step(
    float* state, float* obs_out, float* reward_out, uint8_t* done_out, const int32_t* actions,
    float* ep_ret, int32_t* ep_len, float* totals,
    float* ring_ret, int32_t* ring_len, uint32_t* ring_head, int ring_cap,
    int B, int max_episode_steps, uint64_t seed, uint32_t step, uint32_t env_offset, int reset,
    rl_stream_t stream);

/* Standalone samplers.  logp_out may be NULL. */
int rl_sample_categorical(
    const float* logits, int N, int A, uint64_t seed, uint32_t step, uint32_t env_offset,
    int32_t* actions, float* logp_out, rl_stream_t stream);
int rl_sample_gaussian(
    const float* mean, const float* logstd, int N, int D, uint64_t seed, uint32_t step, uint32_t env_offset,
    float* action, float* logp_out, rl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PARL_B200_H_ */
