// GAE reverse-time scan + advantage normalisation.
// Replaces RolloutStorage.compute_returns (mini_gym_learn/ppo/rollout_storage.py:76-90):
// 24 x ~8 ATen ops + two global reductions become two launches.
//
// Data layout: [T,N] row-major, so for a fixed t consecutive threads (envs) touch
// consecutive addresses - every load/store is a fully coalesced 128 B line per warp.
// Bound: HBM.  Algorithmic bytes: 25 B per (t,env): scan reads r,V (8) + done (1) and writes
// returns, raw advantage (8); normalise reads + writes the advantage (8).
//
// Numerics: the scan uses explicit _rn intrinsics (no FMA contraction) in the reference's
// operation order, so `returns` is bit-identical to the fp32 torch CPU path.  Sum and
// sum-of-squares of the raw advantages are accumulated in double, block partials are
// combined in a fixed order by the last block to finish (deterministic), and the
// normalisation uses the UNBIASED std like torch.Tensor.std() (:90).
#include "rl_common.cuh"

namespace rl {

constexpr int kGaeUnroll = 8;

struct GaeWorkspace {       // lives at the head of the caller's workspace
  unsigned int ticket;      // blocks finished
  unsigned int pad;
  double stats[3];          // sum, sumsq, count (mirrors stats_out)
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
gae_scan_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                const uint8_t* __restrict__ dones, const float* __restrict__ last_values,
                float* __restrict__ returns, float* __restrict__ advantages, int T, int N,
                float gamma, float lam, GaeWorkspace* ws, double2* partials,
                double* stats_out) {
  const int n = blockIdx.x * BLOCK + threadIdx.x;
  double s = 0.0, ss = 0.0;
  if (n < N) {
    float adv = 0.f;
    float next_v = last_values[n];
    int t = T - 1;
    // chunks of kGaeUnroll time steps: issue all loads of a chunk, then run the recurrence
    while (t >= 0) {
      float r[kGaeUnroll], v[kGaeUnroll];
      uint8_t d[kGaeUnroll];
#pragma unroll
      for (int k = 0; k < kGaeUnroll; ++k) {
        const int tt = t - k;
        if (tt >= 0) {
          const size_t idx = (size_t)tt * N + n;
          r[k] = __ldg(rewards + idx);
          v[k] = __ldg(values + idx);
          d[k] = __ldg(dones + idx);
        }
      }
#pragma unroll
      for (int k = 0; k < kGaeUnroll; ++k) {
        const int tt = t - k;
        if (tt >= 0) {
          const size_t idx = (size_t)tt * N + n;
          // rollout_storage.py:83-86, same association order as the eager expression
          const float nt = __fsub_rn(1.0f, (float)d[k]);
          const float delta =
              __fsub_rn(__fadd_rn(r[k], __fmul_rn(__fmul_rn(nt, gamma), next_v)), v[k]);
          adv = __fadd_rn(delta, __fmul_rn(__fmul_rn(__fmul_rn(nt, gamma), lam), adv));
          const float ret = __fadd_rn(adv, v[k]);
          returns[idx] = ret;
          const float a = __fsub_rn(ret, v[k]);  // :89 advantages = returns - values
          advantages[idx] = a;
          s += (double)a;
          ss += (double)a * (double)a;
          next_v = v[k];
        }
      }
      t -= kGaeUnroll;
    }
  }
  // block reduction of (s, ss) in double
  __shared__ double sh_s[BLOCK / 32], sh_ss[BLOCK / 32];
  __shared__ bool is_last;
  s = warp_sum(s);
  ss = warp_sum(ss);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh_s[warp] = s; sh_ss[warp] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double bs = 0.0, bss = 0.0;
#pragma unroll
    for (int w = 0; w < BLOCK / 32; ++w) { bs += sh_s[w]; bss += sh_ss[w]; }
    partials[blockIdx.x] = make_double2(bs, bss);
    __threadfence();
    const unsigned int done = atomicAdd(&ws->ticket, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    // fixed-order combine of the per-block partials by one warp: deterministic result
    if (warp == 0) {
      double a = 0.0, b = 0.0;
      for (unsigned int i = lane; i < gridDim.x; i += 32) {
        const double2 p = __ldcg(partials + i);
        a += p.x; b += p.y;
      }
      a = warp_sum(a);
      b = warp_sum(b);
      if (lane == 0) {
        const double cnt = (double)T * (double)N;
        ws->stats[0] = a; ws->stats[1] = b; ws->stats[2] = cnt;
        if (stats_out) { stats_out[0] = a; stats_out[1] = b; stats_out[2] = cnt; }
        ws->ticket = 0;  // re-arm for the next call (graph replays included)
      }
    }
  }
}

__global__ void __launch_bounds__(256)
gae_normalize_kernel(float* __restrict__ advantages, size_t total, const double* __restrict__ stats) {
  const double sum = stats[0], sumsq = stats[1], cnt = stats[2];
  const double mean_d = sum / cnt;
  double var = (sumsq - cnt * mean_d * mean_d) / (cnt - 1.0);  // unbiased, like torch .std()
  var = var > 0.0 ? var : 0.0;
  const float mean = (float)mean_d;
  const float denom = __fadd_rn((float)sqrt(var), 1e-8f);
  const size_t n4 = total / 4;
  float4* a4 = reinterpret_cast<float4*>(advantages);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = a4[i];
    v.x = __fdiv_rn(__fsub_rn(v.x, mean), denom);
    v.y = __fdiv_rn(__fsub_rn(v.y, mean), denom);
    v.z = __fdiv_rn(__fsub_rn(v.z, mean), denom);
    v.w = __fdiv_rn(__fsub_rn(v.w, mean), denom);
    a4[i] = v;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
    advantages[i] = __fdiv_rn(__fsub_rn(advantages[i], mean), denom);
}

static inline int gae_blocks(int N, int block) { return (N + block - 1) / block; }

}  // namespace rl

using namespace rl;

extern "C" int64_t rl_gae_workspace_bytes(int32_t num_envs) {
  // header + one double2 partial per block at the smallest block size (64)
  const int64_t blocks = (num_envs + 63) / 64;
  return 64 + blocks * (int64_t)sizeof(double2);
}

extern "C" int rl_gae_scan(const float* rewards, const float* values, const uint8_t* dones,
                           const float* last_values, float* returns, float* advantages,
                           int32_t T, int32_t N, float gamma, float lam, void* workspace,
                           double* stats_out, void* stream) {
  RL_REQUIRE(rewards && values && dones && last_values && returns && advantages && workspace,
             RL_ERR_BAD_ARG, "rl_gae_scan: null pointer");
  RL_REQUIRE(T > 0 && N > 0, RL_ERR_BAD_ARG, "rl_gae_scan: T=%d N=%d must be positive", T, N);
  RL_REQUIRE(((uintptr_t)workspace & 15) == 0, RL_ERR_BAD_ARG, "rl_gae_scan: workspace not 16B aligned");
  auto* ws = reinterpret_cast<GaeWorkspace*>(workspace);
  auto* partials = reinterpret_cast<double2*>(reinterpret_cast<char*>(workspace) + 64);
  cudaStream_t st = (cudaStream_t)stream;
  // small N: 64-thread blocks spread the scan over more SMs; large N: 128
  if (N < 148 * 128) {
    gae_scan_kernel<64><<<gae_blocks(N, 64), 64, 0, st>>>(rewards, values, dones, last_values, returns,
                                                        advantages, T, N, gamma, lam, ws, partials,
                                                        stats_out);
  } else {
    gae_scan_kernel<128><<<gae_blocks(N, 128), 128, 0, st>>>(rewards, values, dones, last_values,
                                                          returns, advantages, T, N, gamma, lam, ws,
                                                          partials, stats_out);
  }
  return check_launch("gae_scan_kernel");
}

extern "C" int rl_gae_normalize(float* advantages, int32_t T, int32_t N, const double* stats,
                                void* stream) {
  RL_REQUIRE(advantages && stats, RL_ERR_BAD_ARG, "rl_gae_normalize: null pointer");
  RL_REQUIRE(((uintptr_t)advantages & 15) == 0, RL_ERR_BAD_ARG,
             "rl_gae_normalize: advantages not 16B aligned");
  const size_t total = (size_t)T * N;
  int blocks = (int)((total / 4 + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  gae_normalize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(advantages, total, stats);
  return check_launch("gae_normalize_kernel");
}

extern "C" int rl_gae(const float* rewards, const float* values, const uint8_t* dones,
                      const float* last_values, float* returns, float* advantages, int32_t T,
                      int32_t N, float gamma, float lam, void* workspace, void* stream) {
  int rc = rl_gae_scan(rewards, values, dones, last_values, returns, advantages, T, N, gamma, lam,
                       workspace, nullptr, stream);
  if (rc != RL_OK) return rc;
  auto* ws = reinterpret_cast<GaeWorkspace*>(workspace);
  return rl_gae_normalize(advantages, T, N, ws->stats, stream);
}
