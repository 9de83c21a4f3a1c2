// Reverse linear recurrences over time (kernel family iii-a): discounted return and GAE.
//
// Replaces tf.scan(reverse=True)/tf.foldr in utils/value_ops.py:73-97,146-158 of the reference.
// Both are y_t = a_t * y_{t+1} + b_t with y_T = init:
//   discounted_return: a = discount_t,          b = reward_t,                       init = final_value
//   GAE:               a = lambda * discount_t, b = r_t + d_t * V_{t+1} - V_t,      init = 0
// Batch-major [B,T] (the layout PPOAgent passes, time_major=False): one warp per trajectory,
// 32-step tiles walked from the end; inside a tile the (A,B) affine maps are composed with a
// 5-step shuffle scan, so the T-long dependency chain becomes T/32 carries. HBM-bound:
// 20 B per (b,t) for GAE (read r,d,V; write adv; V re-read hits L1/L2), 12 B for returns.
// Time-major [T,B]: one thread per trajectory (coalesced across b), serial in t.
#include <math.h>

#include "common.cuh"

namespace b200rl {

struct Affine {
  float a, b;
};

// Suffix-compose affine maps across the warp: lane l ends with the map of elements l..31.
__device__ __forceinline__ Affine warp_suffix_compose(Affine m, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float a2 = __shfl_down_sync(0xffffffffu, m.a, o);
    float b2 = __shfl_down_sync(0xffffffffu, m.b, o);
    if (lane + o < 32) {
      // y = a*(a2*y' + b2) + b
      m.b = __fadd_rn(__fmul_rn(m.a, b2), m.b);
      m.a = __fmul_rn(m.a, a2);
    }
  }
  return m;
}

// kind 0: discounted return; kind 1: GAE.
template <int KIND>
__global__ void __launch_bounds__(256) scan_batch_major(const float* __restrict__ rewards,
                                                        const float* __restrict__ discounts,
                                                        const float* __restrict__ values,
                                                        const float* __restrict__ final_value,
                                                        float td_lambda, float* __restrict__ out,
                                                        int64_t B, int64_t T, int provide_all,
                                                        int64_t ld_in, int64_t ld_out,
                                                        int64_t fv_stride) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* r = rewards + b * ld_in;
  const float* d = discounts + b * ld_in;
  const float* v = KIND == 1 ? values + b * ld_in : nullptr;
  const float fin = final_value ? final_value[b * fv_stride] : 0.f;
  float carry = KIND == 1 ? 0.f : fin;
  const int64_t ntiles = (T + 31) / 32;
  for (int64_t tile = ntiles - 1; tile >= 0; --tile) {
    const int64_t t = tile * 32 + lane;
    Affine m{1.f, 0.f};
    if (t < T) {
      const float rt = r[t], dt = d[t];
      if (KIND == 1) {
        const float nv = (t + 1 < T) ? v[t + 1] : fin;
        // delta = r + d * V' - V   (value_ops.py:143), unfused like the reference's op chain
        m.b = __fsub_rn(__fadd_rn(rt, __fmul_rn(dt, nv)), v[t]);
        m.a = __fmul_rn(dt, td_lambda);  // weighted_discounts (:144)
      } else {
        m.a = dt;
        m.b = rt;
      }
    }
    m = warp_suffix_compose(m, lane);
    const float y = __fadd_rn(__fmul_rn(m.a, carry), m.b);
    if (t < T && (provide_all || t == 0)) {
      if (provide_all) out[b * ld_out + t] = y;
      else out[b] = y;
    }
    carry = __shfl_sync(0xffffffffu, y, 0);
  }
}

template <int KIND>
__global__ void __launch_bounds__(128) scan_time_major(const float* __restrict__ rewards,
                                                       const float* __restrict__ discounts,
                                                       const float* __restrict__ values,
                                                       const float* __restrict__ final_value,
                                                       float td_lambda, float* __restrict__ out,
                                                       int64_t B, int64_t T, int provide_all) {
  pdl_prologue();
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float fin = final_value ? final_value[b] : 0.f;
  float acc = KIND == 1 ? 0.f : fin;
  float nv = fin;
#pragma unroll 4
  for (int64_t t = T - 1; t >= 0; --t) {
    const float rt = rewards[t * B + b], dt = discounts[t * B + b];
    if (KIND == 1) {
      const float vt = values[t * B + b];
      const float delta = __fsub_rn(__fadd_rn(rt, __fmul_rn(dt, nv)), vt);
      acc = __fadd_rn(delta, __fmul_rn(__fmul_rn(dt, td_lambda), acc));  // :146-148
      nv = vt;
    } else {
      acc = __fadd_rn(__fmul_rn(acc, dt), rt);  // :73-75
    }
    if (provide_all) out[t * B + b] = acc;
  }
  if (!provide_all) out[b] = acc;
}

// trajectories/trajectory.py:815-832: thread per row, exact serial order of the foldr.
__global__ void nstep_reduce_kernel(const float* __restrict__ reward,
                                    const float* __restrict__ discount, float gamma,
                                    float gamma_pow, float* __restrict__ out_reward,
                                    float* __restrict__ out_discount, int64_t B, int64_t T) {
  pdl_prologue();
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t n = T - 1;
  float acc = 0.f, prod = 1.f;
  for (int64_t t = n - 1; t >= 0; --t)
    acc = __fadd_rn(__fmul_rn(acc, __fmul_rn(gamma, discount[b * T + t])), reward[b * T + t]);
  for (int64_t t = 0; t < n; ++t) prod = __fmul_rn(prod, discount[b * T + t]);
  out_reward[b] = acc;
  out_discount[b] = __fmul_rn(gamma_pow, prod);
}

}  // namespace b200rl

using namespace b200rl;

extern "C" {

int b200rl_discounted_return(const float* rewards, const float* discounts,
                             const float* final_value, float* out, int64_t B, int64_t T,
                             int time_major, int provide_all, void* stream) {
  B200RL_CHECK_ARG(rewards && discounts && out, "discounted_return: NULL argument");
  B200RL_CHECK_ARG(B >= 0 && T >= 1, "discounted_return: B=%lld T=%lld", (long long)B, (long long)T);
  if (B == 0) return B200RL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (time_major) {
    B200RL_LAUNCH(scan_time_major<0>, (unsigned)((B + 127) / 128), 128, 0, st, rewards, discounts, nullptr, final_value, 0.f, out, B, T, provide_all);
  } else {
    B200RL_LAUNCH(scan_batch_major<0>, (unsigned)((B + 7) / 8), 256, 0, st, rewards, discounts, nullptr, final_value, 0.f, out, B, T, provide_all, T, T, 1);
  }
  B200RL_CHECK_LAUNCH("discounted_return");
  return B200RL_OK;
}

int b200rl_gae(const float* values, const float* final_value, const float* discounts,
               const float* rewards, float td_lambda, float* out_adv, int64_t B, int64_t T,
               int time_major, void* stream) {
  B200RL_CHECK_ARG(values && final_value && discounts && rewards && out_adv,
                   "gae: NULL argument");
  B200RL_CHECK_ARG(B >= 0 && T >= 1, "gae: B=%lld T=%lld", (long long)B, (long long)T);
  if (B == 0) return B200RL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (time_major) {
    B200RL_LAUNCH(scan_time_major<1>, (unsigned)((B + 127) / 128), 128, 0, st, rewards, discounts, values, final_value, td_lambda, out_adv, B, T, 1);
  } else {
    B200RL_LAUNCH(scan_batch_major<1>, (unsigned)((B + 7) / 8), 256, 0, st, rewards, discounts, values, final_value, td_lambda, out_adv, B, T, 1, T, T, 1);
  }
  B200RL_CHECK_LAUNCH("gae");
  return B200RL_OK;
}

/* Batch-major scans over the first T columns of [B, ld] arrays (ld >= T): used by PPO, whose
 * returns/advantages cover T-1 of the T collected steps (ppo_agent.py:617-719); out rows have
 * stride ld_out, final_value[b] is read at final_value[b * fv_stride]. */
int b200rl_discounted_return_ld(const float* rewards, const float* discounts,
                                const float* final_value, float* out, int64_t B, int64_t T,
                                int64_t ld_in, int64_t ld_out, int64_t fv_stride, void* stream) {
  B200RL_CHECK_ARG(rewards && discounts && out, "discounted_return_ld: NULL argument");
  B200RL_CHECK_ARG(B >= 1 && T >= 1 && ld_in >= T && ld_out >= T, "discounted_return_ld: sizes");
  B200RL_LAUNCH(scan_batch_major<0>, (unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream, rewards, discounts, nullptr, final_value, 0.f, out, B, T, 1, ld_in, ld_out, fv_stride);
  B200RL_CHECK_LAUNCH("discounted_return_ld");
  return B200RL_OK;
}

int b200rl_gae_ld(const float* values, const float* final_value, const float* discounts,
                  const float* rewards, float td_lambda, float* out_adv, int64_t B, int64_t T,
                  int64_t ld_in, int64_t ld_out, int64_t fv_stride, void* stream) {
  B200RL_CHECK_ARG(values && final_value && discounts && rewards && out_adv, "gae_ld: NULL");
  B200RL_CHECK_ARG(B >= 1 && T >= 1 && ld_in >= T && ld_out >= T, "gae_ld: sizes");
  B200RL_LAUNCH(scan_batch_major<1>, (unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream, rewards, discounts, values, final_value, td_lambda, out_adv, B, T, 1, ld_in, ld_out, fv_stride);
  B200RL_CHECK_LAUNCH("gae_ld");
  return B200RL_OK;
}

int b200rl_nstep_reduce(const float* reward, const float* discount, double gamma,
                        float* out_reward, float* out_discount, int64_t B, int64_t T,
                        void* stream) {
  B200RL_CHECK_ARG(reward && discount && out_reward && out_discount, "nstep_reduce: NULL");
  B200RL_CHECK_ARG(T >= 2, "Trajectory frame count must be at least 2, but saw %lld", (long long)T);
  if (B == 0) return B200RL_OK;
  B200RL_LAUNCH(nstep_reduce_kernel, (unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream, reward, discount, (float)gamma, (float)pow(gamma, (double)(T - 2)), out_reward, out_discount, B, T);
  B200RL_CHECK_LAUNCH("nstep_reduce");
  return B200RL_OK;
}

}  // extern "C"
