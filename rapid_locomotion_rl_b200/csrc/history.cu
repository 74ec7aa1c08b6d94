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

}  // namespace rl

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
