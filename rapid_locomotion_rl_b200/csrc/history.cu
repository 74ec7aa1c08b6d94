// Observation-history ring buffer.
// Replaces HistoryWrapper.step's `torch.cat((obs_history[:, num_obs:], obs), dim=-1)`
// (mini_gym/envs/wrappers/history_wrapper.py:23), which re-reads and re-writes the whole
// [N, H*num_obs] history every step (4.9 KB/env).  The ring keeps 2H slots per env and writes
// each new observation to slots k and k+H, so the newest H observations are always the
// contiguous span [(k+1)*num_obs, (k+1+H)*num_obs) of the row, oldest first - the same
// element order as the reference - at 0.5 KB/env of traffic.
#include "rl_common.cuh"

namespace rl {

__global__ void __launch_bounds__(256)
history_push_kernel(float* __restrict__ hist, const float* __restrict__ obs, int N, int num_obs, int H, int slot) {
  const size_t total = (size_t)N * num_obs;
  const size_t pitch = (size_t)2 * H * num_obs;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / num_obs, c = i - n * num_obs;
    const float v = obs[i];
    float* row = hist + n * pitch;
    row[(size_t)slot * num_obs + c] = v;
    row[(size_t)(slot + H) * num_obs + c] = v;
  }
}

// RolloutStorage.add_transitions (mini_gym_learn/ppo/rollout_storage.py:54-71): the reference issues eleven
// copy_ kernels per step; here one launch writes the whole transition of every env into the [t] slices.
// One warp per env row: every global access is a coalesced run along the row (the 630-float history row
// dominates: 2.5 of the 2.9 KB per transition).
struct StorageAddArgs {
  const float* src[9];      // obs, priv, hist, actions, mu, sigma, rewards, values, logp  ([N, dim] rows, pitch ld[i])
  float* dst[9];            // the [t] slices of the storage, dense [N, dim]
  int dim[9];
  long long ld[9];
  const float* bins; float* dst_bins;          // [N]
  const uint8_t* dones; uint8_t* dst_dones;    // [N] (torch.bool / uint8)
  int N;
};

__global__ void __launch_bounds__(256)
storage_add_kernel(const __grid_constant__ StorageAddArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.N) return;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    if (!a.src[t]) continue;              // rows 0-2 (obs / priv / history) may already be in place: PPO.act stores them
    const float* s = a.src[t] + (size_t)warp * a.ld[t];
    float* d = a.dst[t] + (size_t)warp * a.dim[t];
    for (int c = lane; c < a.dim[t]; c += 32) d[c] = s[c];
  }
  if (lane == 0) {
    a.dst_bins[warp] = a.bins[warp];
    a.dst_dones[warp] = a.dones[warp];
  }
}

}  // namespace rl

extern "C" int rl_storage_add(const RlStorageAdd* q, void* stream) {
  RL_REQUIRE(q && q->N > 0, RL_ERR_BAD_ARG, "rl_storage_add: no envs");
  rl::StorageAddArgs a;
  const void* src[9] = {q->obs, q->priv, q->hist, q->actions, q->mu, q->sigma, q->rewards, q->values, q->logp};
  void* dst[9] = {q->dst_obs, q->dst_priv, q->dst_hist, q->dst_actions, q->dst_mu, q->dst_sigma, q->dst_rewards, q->dst_values, q->dst_logp};
  const int dim[9] = {q->obs_dim, q->priv_dim, q->hist_dim, q->act_dim, q->act_dim, q->act_dim, 1, 1, 1};
  const long long ld[9] = {q->ld_obs, q->ld_priv, q->ld_hist, q->act_dim, q->act_dim, q->act_dim, 1, 1, 1};
  for (int i = 0; i < 9; ++i) {
    RL_REQUIRE((src[i] || i < 3) && (dst[i] || !src[i]) && dim[i] > 0 && ld[i] >= dim[i], RL_ERR_BAD_ARG, "rl_storage_add: field %d", i);
    a.src[i] = reinterpret_cast<const float*>(src[i]); a.dst[i] = reinterpret_cast<float*>(dst[i]);
    a.dim[i] = dim[i]; a.ld[i] = ld[i];
  }
  RL_REQUIRE(q->bins && q->dst_bins && q->dones && q->dst_dones, RL_ERR_BAD_ARG, "rl_storage_add: bins / dones");
  a.bins = q->bins; a.dst_bins = q->dst_bins; a.dones = q->dones; a.dst_dones = q->dst_dones; a.N = q->N;
  const int blocks = (q->N * 32 + 255) / 256;
  rl::storage_add_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
  return rl::check_launch("storage_add_kernel");
}

extern "C" int rl_history_push(float* hist, const float* obs, int32_t N, int32_t num_obs, int32_t H, int32_t slot,
                               void* stream) {
  RL_REQUIRE(hist && obs, RL_ERR_BAD_ARG, "rl_history_push: null pointer");
  RL_REQUIRE(N > 0 && num_obs > 0 && H > 0 && slot >= 0 && slot < H, RL_ERR_BAD_ARG,
             "rl_history_push: N=%d num_obs=%d H=%d slot=%d", N, num_obs, H, slot);
  const size_t total = (size_t)N * num_obs;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  rl::history_push_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(hist, obs, N, num_obs, H, slot);
  return rl::check_launch("history_push_kernel");
}
