// Standalone terrain-height sampling: LeggedRobot._get_heights (mini_gym/envs/base/legged_robot.py:1469-1503) for a
// list of envs.  The fused step samples the same values inside its own launch; this entry serves direct callers of
// the method.  One warp per env, lanes over the measured points; arithmetic in the reference's order (yaw-only
// rotation of the base-frame points, truncation toward zero, clip to the table, min of three samples).
#include "env_common.cuh"

namespace rl {

__global__ void __launch_bounds__(256)
env_heights_kernel(const __grid_constant__ RlEnvCfg cfg, const float* __restrict__ root_states,
                   const float* __restrict__ height_points, const int16_t* __restrict__ H, const int64_t* __restrict__ ids,
                   int n_ids, float* __restrict__ out) {
  const int wi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wi >= n_ids) return;
  const int e = ids ? (int)ids[wi] : wi;
  const float* r = root_states + (size_t)e * 13;
  const float bx = r[0], by = r[1];
  float yz = r[5], yw = r[6];
  float nrm = sqrtf(yz * yz + yw * yw);
  nrm = fmaxf(nrm, 1e-9f);
  yz = yz / nrm; yw = yw / nrm;
  const int P = cfg.num_height_points;
  for (int p = lane; p < P; p += 32) {
    const float px = height_points[2 * p], py = height_points[2 * p + 1];
    const V3 wp = quat_apply(0.f, 0.f, yz, yw, V3{px, py, 0.f});
    const float fx = (wp.x + bx + cfg.border_size) / cfg.horizontal_scale;
    const float fy = (wp.y + by + cfg.border_size) / cfg.horizontal_scale;
    long long ix = (long long)fx, iy = (long long)fy;
    ix = max(0ll, min(ix, (long long)cfg.hf_rows - 2));
    iy = max(0ll, min(iy, (long long)cfg.hf_cols - 2));
    const int16_t h1 = __ldg(H + ix * cfg.hf_cols + iy);
    const int16_t h2 = __ldg(H + (ix + 1) * cfg.hf_cols + iy);
    const int16_t h3 = __ldg(H + ix * cfg.hf_cols + iy + 1);
    out[(size_t)wi * P + p] = (float)min(min(h1, h2), h3) * cfg.vertical_scale;
  }
}

// The step's height phase as a launch of its own (RlEnvBuffers.height_mean set): one warp per env.  Inside the step kernel
// the four warps of a 32-env CTA take 8 envs each in turn - at 4000 envs that is 125 CTAs, one warp per scheduler, ~2000
// dependent instructions per env (51 us per step); here 4000 warps spread over all SMs hide each other's latency.
__global__ void __launch_bounds__(256)
env_heights_prepass_kernel(const __grid_constant__ StepArgs args) {
  const RlEnvCfg& cfg = args.cfg;
  const RlEnvBuffers& b = args.b;
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= cfg.num_envs) return;
  const uint64_t rng_step = args.step + (b.step_state ? b.step_state[0] : 0ull);
  const float* r = b.root_states + (size_t)e * 13;
  float bx = r[0], by = r[1];
  teleport_xy_env(cfg, e, bx, by);                       // the step kernel applies (and stores) the same teleport
  const float hm = sample_heights_env(cfg, b, args.seed, rng_step, e, bx, by, r[2], r[5], r[6], lane, cfg.add_noise != 0);
  if (lane == 0) b.height_mean[e] = hm;
}

int launch_heights_prepass(const StepArgs& args, cudaStream_t st) {
  const int n = args.cfg.num_envs;
  env_heights_prepass_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(args);
  return check_launch("env_heights_prepass_kernel");
}

}  // namespace rl

extern "C" int rl_env_heights(const RlEnvCfg* cfg, const float* root_states, const float* height_points,
                              const int16_t* height_samples, const int64_t* env_ids, int32_t n_ids, float* out, void* stream) {
  using namespace rl;
  RL_REQUIRE(cfg && root_states && height_points && height_samples && out && n_ids > 0, RL_ERR_BAD_ARG,
             "rl_env_heights: null argument");
  RL_REQUIRE(cfg->num_height_points > 0 && cfg->hf_rows >= 2 && cfg->hf_cols >= 2, RL_ERR_BAD_CFG,
             "rl_env_heights: no height points / table (%d points, %d x %d)", cfg->num_height_points, cfg->hf_rows, cfg->hf_cols);
  env_heights_kernel<<<(n_ids * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*cfg, root_states, height_points, height_samples,
                                                                               env_ids, n_ids, out);
  return check_launch("env_heights_kernel");
}
