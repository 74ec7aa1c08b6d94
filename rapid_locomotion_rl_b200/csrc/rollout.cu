// Rollout-step glue of Runner.learn (mini_gym_learn/ppo/__init__.py:126-141) as two launches per step.
//
// The reference's loop body is  actions = alg.act(obs, priv, hist) -> env.step(actions) -> alg.process_env_step(...):
// per step ~25 small kernels around the policy and the env (two casts, Normal sample, log_prob arithmetic, eleven
// storage copies, history shift, reward clone / bootstrap).  Here:
//   rl_rollout_boundary  one warp per env, BETWEEN two steps: closes transition t (reward [+ time-out bootstrap,
//                        ppo.py:81-83], done flag, env bin; pushes the new observation into the history ring,
//                        history_wrapper.py:23) and opens transition t+1 (observation / privileged observation /
//                        history rows into storage[t+1], rollout_storage.py:57-60, and the bf16 staging of the policy
//                        inputs) - every byte of the new observation is read once;
//   rl_rollout_act       one thread per env, AFTER the policy pass: a = mu + std * N(0,1) (actor_critic.py:142-147),
//                        log-probability, and the transition's actions / mu / sigma / log-prob / value slices
//                        (rollout_storage.py:61-69) next to the action buffer the env step reads.
#include <cuda_bf16.h>

#include "rl_common.cuh"

namespace rl {

struct BoundaryArgs { RlRolloutBoundary q; };

// 8-byte row copies with several independent loads in flight (rows are 8 B aligned: even widths and pitches)
template <int U>
__device__ __forceinline__ void copy_row2(const float* __restrict__ s, float* __restrict__ d, int n, int lane) {
  for (int c0 = 2 * lane; c0 < n; c0 += 64 * U) {
    float2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + 64 * u;
      if (c < n) v[u] = *reinterpret_cast<const float2*>(s + c);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = c0 + 64 * u;
      if (c < n) *reinterpret_cast<float2*>(d + c) = v[u];
    }
  }
}

__global__ void __launch_bounds__(256)
rollout_boundary_kernel(const __grid_constant__ BoundaryArgs a) {
  const RlRolloutBoundary& q = a.q;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= q.N) return;
  const int W = q.obs_dim, H = q.H;
  const size_t pitch = (size_t)2 * H * W;
  const float* obs = q.obs + (size_t)n * W;
  float* ring = q.ring + (size_t)n * pitch;
  // the new observation: read once, used for the ring (two slots), the storage row and the bf16 staging
  float o0 = 0.f, o1 = 0.f;                 // columns lane, lane + 32 (obs_dim <= 64)
  if (lane < W) o0 = obs[lane];
  if (lane + 32 < W) o1 = obs[lane + 32];
  if (q.do_post) {
    if (lane < W) { ring[(size_t)q.push_slot * W + lane] = o0; ring[(size_t)(q.push_slot + H) * W + lane] = o0; }
    if (lane + 32 < W) { ring[(size_t)q.push_slot * W + lane + 32] = o1; ring[(size_t)(q.push_slot + H) * W + lane + 32] = o1; }
    if (lane == 0) {
      float r = q.rew[n];
      if (q.time_outs && q.time_outs[n]) r += q.gamma * q.values_prev[n];      // ppo.py:81-83
      q.dst_rewards[n] = r;
      q.dst_dones[n] = q.dones[n];
      q.dst_bins[n] = q.bins ? q.bins[n] : 0.f;
    }
  }
  if (q.do_pre) {
    if (lane < W) q.dst_obs[(size_t)n * W + lane] = o0;
    if (lane + 32 < W) q.dst_obs[(size_t)n * W + lane + 32] = o1;
    __nv_bfloat16* xac = reinterpret_cast<__nv_bfloat16*>(q.Xac) + (size_t)n * q.ld_xac;
    if (lane < W) xac[lane] = __float2bfloat16(o0);
    if (lane + 32 < W) xac[lane + 32] = __float2bfloat16(o1);
    const int P = q.priv_dim;
    const float pv = lane < P ? q.priv[(size_t)n * P + lane] : 0.f;
    if (lane < P) q.dst_priv[(size_t)n * P + lane] = pv;
    if (lane < q.ld_xp) reinterpret_cast<__nv_bfloat16*>(q.Xp)[(size_t)n * q.ld_xp + lane] = __float2bfloat16(pv);
    // history row = the H-slot span that ends with the observation just pushed: H - 1 older slots from the ring
    // (written by earlier launches) + the new observation from registers
    const float* span = ring + (size_t)q.hist_slot * W;
    float* dh = q.dst_hist + (size_t)n * H * W;
    const int older = (H - 1) * W;
    if ((((uintptr_t)span | (uintptr_t)dh) & 7) == 0 && (older & 1) == 0) copy_row2<5>(span, dh, older, lane);
    else for (int c = lane; c < older; c += 32) dh[c] = span[c];
    if (q.do_post) {
      if (lane < W) dh[older + lane] = o0;
      if (lane + 32 < W) dh[older + lane + 32] = o1;
    } else {                                  // first step of a rollout: the ring already holds the newest observation
      for (int c = lane; c < W; c += 32) dh[older + c] = span[older + c];
    }
  }
}

struct ActArgs { RlRolloutAct q; };

__global__ void __launch_bounds__(128)
rollout_act_kernel(const __grid_constant__ ActArgs a) {
  const RlRolloutAct& q = a.q;
  constexpr int ACT = 12;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t step = q.step + (q.step_state ? q.step_state[0] : 0ull);
  if (i < q.N) {
    float lp = 0.f;
#pragma unroll
    for (int b4 = 0; b4 < ACT / 4; ++b4) {
      float z[4];
      if (q.inj_normal) {
#pragma unroll
        for (int k = 0; k < 4; ++k) z[k] = q.inj_normal[(size_t)i * ACT + b4 * 4 + k];
      } else {
        float u[4];
        rng4(q.seed, (uint32_t)i, step, RNG_POLICY, (uint32_t)b4, u);
        const float r0 = sqrtf(-2.f * __logf(fmaxf(u[0], 5.96e-8f))), r1 = sqrtf(-2.f * __logf(fmaxf(u[2], 5.96e-8f)));
        float s0, c0, s1, c1;
        __sincosf(6.283185307179586f * u[1], &s0, &c0);
        __sincosf(6.283185307179586f * u[3], &s1, &c1);
        z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
      }
      float4 av, mv, sv;
      float* ap = &av.x; float* mp = &mv.x; float* sp = &sv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int d = b4 * 4 + k;
        const float mu = q.mean[(size_t)i * ACT + d], sg = q.std[d];
        const float act = mu + sg * z[k];
        ap[k] = act; mp[k] = mu; sp[k] = sg;
        const float diff = act - mu;
        lp += -(diff * diff) / (2.f * sg * sg) - __logf(sg) - 0.9189385332046727f;
      }
      const size_t o = (size_t)i * ACT + b4 * 4;       // 16 B aligned: [N, 12] fp32 rows
      *reinterpret_cast<float4*>(q.actions_out + o) = av;
      if (q.dst_actions) {
        *reinterpret_cast<float4*>(q.dst_actions + o) = av;
        *reinterpret_cast<float4*>(q.dst_mu + o) = mv;
        *reinterpret_cast<float4*>(q.dst_sigma + o) = sv;
      }
    }
    if (q.logp_out) q.logp_out[i] = lp;
    if (q.dst_actions) {
      q.dst_logp[i] = lp;
      q.dst_values[i] = q.value[i];
    }
  }
  if (q.step_state) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(q.step_state + 1), 1ull);
      if (done == gridDim.x - 1) {
        q.step_state[1] = 0;
        atomicAdd(reinterpret_cast<unsigned long long*>(q.step_state), 1ull);
      }
    }
  }
}

}  // namespace rl

using namespace rl;

extern "C" int rl_rollout_boundary(const RlRolloutBoundary* q, void* stream) {
  RL_REQUIRE(q && q->N > 0 && q->obs && q->ring, RL_ERR_BAD_ARG, "rl_rollout_boundary: null argument");
  RL_REQUIRE(q->obs_dim > 0 && q->obs_dim <= 64 && q->priv_dim > 0 && q->priv_dim <= 32 && q->H > 0, RL_ERR_UNSUPPORTED,
             "rl_rollout_boundary: obs_dim=%d (<= 64) priv_dim=%d (<= 32) H=%d", q->obs_dim, q->priv_dim, q->H);
  RL_REQUIRE(q->do_post || q->do_pre, RL_ERR_BAD_ARG, "rl_rollout_boundary: nothing to do");
  if (q->do_post)
    RL_REQUIRE(q->rew && q->dones && q->dst_rewards && q->dst_dones && q->dst_bins && q->push_slot >= 0 && q->push_slot < q->H &&
               (!q->time_outs || q->values_prev), RL_ERR_BAD_ARG, "rl_rollout_boundary: post part incomplete");
  if (q->do_pre)
    RL_REQUIRE(q->priv && q->dst_obs && q->dst_priv && q->dst_hist && q->Xac && q->Xp && q->ld_xac >= q->obs_dim &&
               q->ld_xp >= q->priv_dim && q->ld_xp <= 32 && q->hist_slot >= 0 && q->hist_slot <= q->H, RL_ERR_BAD_ARG,
               "rl_rollout_boundary: pre part incomplete");
  BoundaryArgs a; a.q = *q;
  rollout_boundary_kernel<<<(q->N * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("rollout_boundary_kernel");
}

extern "C" int rl_rollout_act(const RlRolloutAct* q, void* stream) {
  RL_REQUIRE(q && q->N > 0 && q->mean && q->std && q->actions_out, RL_ERR_BAD_ARG, "rl_rollout_act: null argument");
  RL_REQUIRE(!q->dst_actions || (q->dst_mu && q->dst_sigma && q->dst_logp && q->dst_values && q->value), RL_ERR_BAD_ARG,
             "rl_rollout_act: storage slices incomplete");
  RL_REQUIRE((((uintptr_t)q->actions_out | (uintptr_t)q->dst_actions | (uintptr_t)q->dst_mu | (uintptr_t)q->dst_sigma) & 15) == 0,
             RL_ERR_BAD_ARG, "rl_rollout_act: action rows must be 16 B aligned");
  ActArgs a; a.q = *q;
  rollout_act_kernel<<<(q->N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
  return check_launch("rollout_act_kernel");
}
