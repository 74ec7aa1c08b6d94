// Masked / indexed env reset.
// Replaces LeggedRobot.reset_idx (legged_robot.py:227-290) and the helpers it calls:
// _update_terrain_curriculum :793-818, _randomize_dof_props :544-560, _reset_dofs :690-712,
// _reset_root_states :714-755, buffer zeroing :255-259, episode-sum means :261-267, and
// HistoryWrapper.reset_idx row zeroing (history_wrapper.py:34).
// The reference does this with ~216 ATen ops plus a Python loop calling `.item()` per env.
// One thread per candidate env; per-key episode sums leave through a warp-shuffle reduction
// and one double atomicAdd per warp.  Bound: HBM (touches only the reset rows).
#include "rl_common.cuh"

namespace rl {

constexpr int ND_R = RL_NUM_DOF;
constexpr int MAX_SUM_ROWS = RL_EPISODE_ROWS;

struct ResetArgs {
  RlResetCfg cfg;
  RlResetBuffers b;
  uint64_t seed;
  uint64_t step;
};

__device__ inline bool row_live(const RlResetCfg& cfg, int r) {
  return r < cfg.n_terms || (r == RL_ROW_TERMINATION && cfg.has_termination) || r == RL_ROW_TOTAL;
}

__global__ void __launch_bounds__(128)
env_reset_kernel(const __grid_constant__ ResetArgs args) {
  const RlResetCfg& cfg = args.cfg;
  const RlResetBuffers& b = args.b;
  const int N = cfg.num_envs;
  const size_t Ns = (size_t)N;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int e = -1;
  if (b.ids) { if (i < b.n_ids) e = (int)b.ids[i]; }
  else if (i < N && b.mask[i]) e = i;
  const bool active = e >= 0 && e < N;

  double ksum[MAX_SUM_ROWS];
#pragma unroll
  for (int r = 0; r < MAX_SUM_ROWS; ++r) ksum[r] = 0.0;

  if (active) {
    float* root = b.root_states + (size_t)e * 13;
    float ox = b.env_origins[e * 3 + 0], oy = b.env_origins[e * 3 + 1], oz = b.env_origins[e * 3 + 2];
    // ---- terrain curriculum (:793-818) ----
    if (cfg.terrain_curriculum) {
      const float dx = root[0] - ox, dy = root[1] - oy;
      const float distance = sqrtf(dx * dx + dy * dy);
      const float cx = b.commands[e * 4 + 0], cy = b.commands[e * 4 + 1];
      const bool up = distance > cfg.env_length_half;
      const bool down = (distance < sqrtf(cx * cx + cy * cy) * cfg.episode_length_s_half) && !up;
      long long level = b.terrain_levels[e] + (up ? 1 : 0) - (down ? 1 : 0);
      if (level >= cfg.max_terrain_level) {
        float u;
        if (b.level_u) u = b.level_u[e];
        else { float u4[4]; rng4(args.seed, (uint32_t)e, args.step, RNG_RESET, 1, u4); u = u4[0]; }
        level = (long long)(u * (float)cfg.max_terrain_level);
        if (level >= cfg.max_terrain_level) level = cfg.max_terrain_level - 1;
      } else if (level < 0) {
        level = 0;
      }
      b.terrain_levels[e] = level;
      const long long type = b.terrain_types[e];
      const float* o = b.terrain_origins + ((size_t)level * cfg.num_terrain_cols + type) * 3;
      ox = o[0]; oy = o[1]; oz = o[2];
      b.env_origins[e * 3 + 0] = ox; b.env_origins[e * 3 + 1] = oy; b.env_origins[e * 3 + 2] = oz;
    }
    // ---- DOF property draws (:544-560): one scalar per env broadcast over the 12 DOFs ----
    if (cfg.randomize_motor_strength | cfg.randomize_Kp_factor | cfg.randomize_Kd_factor) {
      float u3[4];
      if (b.dr_u) { u3[0] = b.dr_u[e]; u3[1] = b.dr_u[Ns + e]; u3[2] = b.dr_u[2 * Ns + e]; }
      else rng4(args.seed, (uint32_t)e, args.step, RNG_RESET, 0, u3);
      if (cfg.randomize_motor_strength) {
        const float v = u3[0] * cfg.motor_strength_lo_span[1] + cfg.motor_strength_lo_span[0];
#pragma unroll
        for (int j = 0; j < ND_R; ++j) b.motor_strengths[j * Ns + e] = v;
      }
      if (cfg.randomize_Kp_factor) {
        const float v = u3[1] * cfg.Kp_factor_lo_span[1] + cfg.Kp_factor_lo_span[0];
#pragma unroll
        for (int j = 0; j < ND_R; ++j) b.Kp_factors[j * Ns + e] = v;
      }
      if (cfg.randomize_Kd_factor) {
        const float v = u3[2] * cfg.Kd_factor_lo_span[1] + cfg.Kd_factor_lo_span[0];
#pragma unroll
        for (int j = 0; j < ND_R; ++j) b.Kd_factors[j * Ns + e] = v;
      }
    }
    // ---- DOF state (:702-706) ----
    float* dof = b.dof_state + (size_t)e * 24;
#pragma unroll
    for (int j = 0; j < ND_R; ++j) { dof[2 * j] = cfg.default_dof_pos[j]; dof[2 * j + 1] = 0.f; }
    // ---- root state (:724-734) ----
    float r13[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) r13[k] = cfg.base_init_state[k];
    r13[0] += ox; r13[1] += oy; r13[2] += oz;
    if (cfg.custom_origins) {
      float u0, u1;
      if (b.init_u) { u0 = b.init_u[e]; u1 = b.init_u[Ns + e]; }
      else { float u4[4]; rng4(args.seed, (uint32_t)e, args.step, RNG_RESET, 2, u4); u0 = u4[0]; u1 = u4[1]; }
      // torch_rand_float(lower=x_init_range, upper=y_init_range): quirk kept (:727-729)
      const float span = cfg.y_init_range - cfg.x_init_range;
      r13[0] += span * u0 + cfg.x_init_range;
      r13[1] += span * u1 + cfg.x_init_range;
      r13[0] += cfg.x_init_offset;
      r13[1] += cfg.y_init_offset;
    }
#pragma unroll
    for (int k = 0; k < 13; ++k) root[k] = r13[k];
    // ---- buffers (:255-259) ----
#pragma unroll
    for (int j = 0; j < ND_R; ++j) { b.last_actions[j * Ns + e] = 0.f; b.last_dof_vel[j * Ns + e] = 0.f; }
#pragma unroll
    for (int k = 0; k < RL_NUM_FEET; ++k) b.feet_air_time[k * Ns + e] = 0.f;
    b.episode_length_buf[e] = 0;
    b.reset_buf[e] = 1;
    // ---- episode sums (:261-267) ----
#pragma unroll
    for (int r = 0; r < MAX_SUM_ROWS; ++r) {
      if (row_live(cfg, r)) {
        ksum[r] = (double)b.episode_sums[r * Ns + e];
        b.episode_sums[r * Ns + e] = 0.f;
      }
    }
  }

  // warp reduce + one atomic per warp and key
  const unsigned lane = threadIdx.x & 31;
  const unsigned any = __ballot_sync(0xffffffffu, active);
  if (any && b.episode_sum_out) {
#pragma unroll
    for (int r = 0; r < MAX_SUM_ROWS; ++r) {
      if (row_live(cfg, r)) {
        const double s = warp_sum(ksum[r]);
        if (lane == 0) atomicAdd(b.episode_sum_out + r, s);
      }
    }
    if (lane == 0) atomicAdd(b.episode_sum_out + RL_EPISODE_ROWS, (double)__popc(any));
  }

  // observation-history rows (history_wrapper.py:34): the warp zeroes each active lane's row
  if (b.obs_history && any) {
    const int H = b.obs_history_len;
    unsigned m = any;
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int ee = __shfl_sync(0xffffffffu, e, src);
      float* row = b.obs_history + (size_t)ee * H;
      for (int c = lane; c < H; c += 32) row[c] = 0.f;
    }
  }
}

}  // namespace rl

using namespace rl;

extern "C" int rl_env_reset(const RlResetCfg* cfg, const RlResetBuffers* b, uint64_t seed, uint64_t step,
                            void* stream) {
  RL_REQUIRE(cfg && b, RL_ERR_BAD_ARG, "rl_env_reset: null cfg/buffers");
  RL_REQUIRE(cfg->num_envs > 0, RL_ERR_BAD_CFG, "rl_env_reset: num_envs=%d", cfg->num_envs);
  RL_REQUIRE((b->mask != nullptr) != (b->ids != nullptr), RL_ERR_BAD_ARG,
             "rl_env_reset: exactly one of mask / ids must be given");
  RL_REQUIRE(cfg->n_terms >= 0 && cfg->n_terms <= RL_MAX_TERMS, RL_ERR_BAD_CFG,
             "rl_env_reset: n_terms=%d", cfg->n_terms);
  RL_REQUIRE(b->root_states && b->dof_state && b->env_origins && b->last_actions && b->last_dof_vel &&
             b->feet_air_time && b->episode_length_buf && b->reset_buf && b->Kp_factors && b->Kd_factors &&
             b->motor_strengths && b->episode_sums, RL_ERR_BAD_ARG, "rl_env_reset: a required buffer is null");
  if (cfg->terrain_curriculum)
    RL_REQUIRE(b->terrain_levels && b->terrain_types && b->terrain_origins && b->commands, RL_ERR_BAD_ARG,
               "rl_env_reset: terrain curriculum buffers missing");
  const int n = b->ids ? b->n_ids : cfg->num_envs;
  if (n <= 0) return RL_OK;  // len(env_ids) == 0 (:238)
  ResetArgs args;
  args.cfg = *cfg; args.b = *b; args.seed = seed; args.step = step;
  env_reset_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(args);
  return check_launch("env_reset_kernel");
}
