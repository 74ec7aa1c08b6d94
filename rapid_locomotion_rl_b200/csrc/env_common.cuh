// Shared pieces of the env-step kernels: launch arguments, tile staging, quaternion math in the
// reference's operation order, 16-bit uniform lanes.
#pragma once

#include "rl_common.cuh"

namespace rl {


constexpr int ND = RL_NUM_DOF;

struct StepArgs {
  RlEnvCfg cfg;
  RlEnvBuffers b;
  uint64_t seed;
  uint64_t step;
};

// ---------------------------------------------------------------------------------------
// cooperative tile copies
// ---------------------------------------------------------------------------------------
// (fallback path for ragged tail tiles / unaligned tensors; full tiles use cp.async.bulk)
template <int TILE>
__device__ inline void stage_in(float* __restrict__ dst, const float* __restrict__ src, int n_floats) {
  if ((((uintptr_t)src) & 15) == 0) {
    const int n4 = n_floats >> 2;
    // batches of 4 independent 128-bit loads per thread before the first store
    for (int i = threadIdx.x; i < n4; i += 4 * TILE) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) if (i + k * TILE < n4) v[k] = ldg_stream4(src + 4 * (i + k * TILE));
#pragma unroll
      for (int k = 0; k < 4; ++k) if (i + k * TILE < n4) reinterpret_cast<float4*>(dst)[i + k * TILE] = v[k];
    }
#pragma unroll 1
    for (int i = (n4 << 2) + threadIdx.x; i < n_floats; i += TILE) dst[i] = __ldg(src + i);
  } else {
#pragma unroll 1
    for (int i = threadIdx.x; i < n_floats; i += TILE) dst[i] = __ldg(src + i);
  }
}

template <int TILE>
__device__ inline void stage_out(float* __restrict__ dst, const float* __restrict__ src, int n_floats) {
  if ((((uintptr_t)dst) & 15) == 0) {
    const int n4 = n_floats >> 2;
#pragma unroll 2
    for (int i = threadIdx.x; i < n4; i += TILE)
      stg_stream4(dst + 4 * i, reinterpret_cast<const float4*>(src)[i]);
#pragma unroll 1
    for (int i = (n4 << 2) + threadIdx.x; i < n_floats; i += TILE) dst[i] = src[i];
  } else {
#pragma unroll 1
    for (int i = threadIdx.x; i < n_floats; i += TILE) dst[i] = src[i];
  }
}

// rows of `width` floats in smem (dense) -> global rows with pitch `pitch`
template <int TILE>
__device__ inline void stage_out_rows(float* __restrict__ dst, const float* __restrict__ src, int rows,
                                      int width, int pitch) {
  const int total = rows * width;
#pragma unroll 1
  for (int i = threadIdx.x; i < total; i += TILE) {
    const int r = i / width, c = i - r * width;
    dst[(size_t)r * pitch + c] = src[i];
  }
}

// ---------------------------------------------------------------------------------------
// small math, written in the reference's operation order
// ---------------------------------------------------------------------------------------
struct V3 { float x, y, z; };

// isaacgym.torch_utils.quat_rotate_inverse (xyzw): a - b + c with
// a = v*(2w^2-1), b = cross(qv,v)*w*2, c = qv*dot(qv,v)*2
__device__ inline V3 quat_rotate_inverse(float qx, float qy, float qz, float qw, V3 v) {
  const float s = 2.0f * (qw * qw) - 1.0f;
  V3 a = {v.x * s, v.y * s, v.z * s};
  V3 cr = {qy * v.z - qz * v.y, qz * v.x - qx * v.z, qx * v.y - qy * v.x};
  V3 b = {cr.x * qw * 2.0f, cr.y * qw * 2.0f, cr.z * qw * 2.0f};
  const float d = (qx * v.x + qy * v.y) + qz * v.z;
  V3 c = {qx * d * 2.0f, qy * d * 2.0f, qz * d * 2.0f};
  return {a.x - b.x + c.x, a.y - b.y + c.y, a.z - b.z + c.z};
}

// isaacgym.torch_utils.quat_apply: v + w*t + cross(qv,t), t = 2*cross(qv,v)
__device__ inline V3 quat_apply(float qx, float qy, float qz, float qw, V3 v) {
  V3 t = {(qy * v.z - qz * v.y) * 2.0f, (qz * v.x - qx * v.z) * 2.0f, (qx * v.y - qy * v.x) * 2.0f};
  V3 c = {qy * t.z - qz * t.y, qz * t.x - qx * t.z, qx * t.y - qy * t.x};
  return {v.x + qw * t.x + c.x, v.y + qw * t.y + c.y, v.z + qw * t.z + c.z};
}

__device__ inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ inline float sq(float x) { return x * x; }

// torch.remainder for floats (sign follows the divisor) - math_utils.py:20 `angles %= 2*pi`
__device__ inline float py_mod(float a, float b) {
  float m = fmodf(a, b);
  if (m != 0.0f && ((b < 0.0f) != (m < 0.0f))) m += b;
  return m;
}

// ---------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------
// 16-bit uniform lane k (0..7) of a Philox block -> (u - 0.5) in (-0.5, 0.5), symmetric, never +-0.5
__device__ inline float centered_u16(const uint32_t (&r)[4], int k) {
  const uint32_t x = (r[k >> 1] >> (16 * (k & 1))) & 0xffffu;
  return __fmaf_rn((float)x, 1.0f / 65536.0f, 0.5f / 65536.0f - 0.5f);
}


// Terrain heights under ONE env (legged_robot.py:1469-1503), by one warp - lanes over the measured points: writes
// measured_heights[ge], the height suffix of obs_buf[ge] (:386-389 + noise :392 + clip :134) and returns
// mean(z - heights) (the base_height term's input) on every lane.  (bx, by, bz, yz, yw): base position (after the
// teleport) and the z / w components of the base quaternion.  Shared by the step kernels (four warps take the envs of a
// tile in turn) and by the pre-pass kernel (heights.cu: one warp per env over the whole GPU): same arithmetic, same
// per-lane summation order, identical bits.
// Six 32-point chunks at a time (187 points = one pass): pass 1 computes the cell indices and ISSUES the three int16
// gathers of every chunk (18 independent L2 loads in flight per lane); pass 2 consumes them in per-lane point order.
__device__ inline float sample_heights_env(const RlEnvCfg& cfg, const RlEnvBuffers& b, uint64_t seed, uint64_t rng_step, int ge,
                                           float bx, float by, float bz, float yz, float yw, int lane, bool c_noise) {
  const int P = cfg.num_height_points;
  const int Wc = cfg.num_obs - P;
  const float hscale = cfg.horizontal_scale, vscale = cfg.vertical_scale, co = cfg.clip_obs;
  float nrm = sqrtf(yz * yz + yw * yw);
  nrm = fmaxf(nrm, 1e-9f);
  yz = yz / nrm; yw = yw / nrm;
  float acc = 0.f;
  float* mh = b.measured_heights + (size_t)ge * P;
  float* ob = b.obs_buf + (size_t)ge * cfg.num_obs + Wc;
  const float* nu = b.noise_u ? b.noise_u + (size_t)ge * cfg.num_obs : nullptr;
  // observation noise: Philox block j serves points [8 j, 8 j + 8); lane L computes the blocks L, L + 32, .. once
  // (187 points = 24 blocks: one round), the consumers fetch theirs by shuffle
  const bool philox = c_noise && !nu;
  constexpr int HU = 6;
  const int hf_cols = cfg.hf_cols, ix_max = cfg.hf_rows - 2, iy_max = cfg.hf_cols - 2;
  const int16_t* H = b.height_samples;
#pragma unroll 1
  for (int p0 = 0; p0 < P; p0 += 32 * HU) {
    uint32_t blk[4] = {0u, 0u, 0u, 0u};
    if (philox && p0 + 8 * lane < P)       // blocks p0 / 8 + lane of this pass (32 * HU / 8 = 24 per pass)
      Philox::gen(seed, (uint32_t)ge, (uint32_t)rng_step, (uint32_t)(rng_step >> 32),
                  (RNG_NOISE << 16) | (uint32_t)(64 + (p0 >> 3) + lane), blk);
    int16_t hs[HU][3];
#pragma unroll
    for (int u = 0; u < HU; ++u) {
      const int p = p0 + 32 * u + lane;
      hs[u][0] = hs[u][1] = hs[u][2] = 0;
      if (p < P && !cfg.heights_plane) {
        const float px = b.height_points[2 * p], py = b.height_points[2 * p + 1];
        V3 wp = quat_apply(0.f, 0.f, yz, yw, V3{px, py, 0.f});
        const float fx = (wp.x + bx + cfg.border_size) / hscale;
        const float fy = (wp.y + by + cfg.border_size) / hscale;
        // .long() truncates toward zero; the clip to the table makes the saturating 32-bit conversion equivalent
        const int ix = max(0, min(__float2int_rz(fx), ix_max));
        const int iy = max(0, min(__float2int_rz(fy), iy_max));
        const int16_t* h0 = H + ix * hf_cols + iy;
        hs[u][0] = __ldg(h0);
        hs[u][1] = __ldg(h0 + hf_cols);
        hs[u][2] = __ldg(h0 + 1);
      }
    }
#pragma unroll
    for (int u = 0; u < HU; ++u) {
      const int pbase = p0 + 32 * u;
      if (pbase >= P) break;                     // warp uniform
      const int p = pbase + lane;
      uint32_t r4[4] = {0u, 0u, 0u, 0u};
      if (philox) {
#pragma unroll
        for (int k = 0; k < 4; ++k) r4[k] = __shfl_sync(0xffffffffu, blk[k], 4 * u + (lane >> 3));
      }
      if (p < P) {
        const float h = cfg.heights_plane ? 0.f : (float)min(min(hs[u][0], hs[u][1]), hs[u][2]) * vscale;
        mh[p] = h;
        acc += bz - h;
        float o = clampf(bz - 0.5f - h, -1.f, 1.f) * cfg.obs_scale_height;
        if (c_noise) {
          if (nu) o += (2.0f * nu[Wc + p] - 1.0f) * cfg.noise_scale_height;
          else o = __fmaf_rn(2.0f * centered_u16(r4, p & 7), cfg.noise_scale_height, o);
        }
        ob[p] = clampf(o, -co, co);
      }
    }
  }
  acc = warp_sum(acc);
  return acc / (float)P;
}

// ---- train / eval split (legged_robot.py:456-469): the per-range fields of the configuration -------------------------
// The reference calls _teleport_robots, _push_robots and _randomize_dof_props once per env range with that range's Cfg;
// here an env picks its variant by index.  Without a split (num_train_envs == num_envs or 0) every env is a train env.
struct EnvVariant {
  int teleport_robots;
  float teleport_lo_x, teleport_hi_x, teleport_shift_x, teleport_lo_y, teleport_hi_y, teleport_shift_y;
  int randomize_motor_strength, randomize_Kp_factor, randomize_Kd_factor;
  float motor_strength_lo_span[2], Kp_factor_lo_span[2], Kd_factor_lo_span[2];
  int push_robots, push_interval;
  float push_lo_span[2];
};
__device__ __host__ inline bool has_eval_split(const RlEnvCfg& cfg) {
  return cfg.num_train_envs > 0 && cfg.num_train_envs < cfg.num_envs;
}
__device__ inline EnvVariant env_variant(const RlEnvCfg& cfg, int e) {
  EnvVariant v;
  const bool ev = has_eval_split(cfg) && e >= cfg.num_train_envs;
#define RL_PICK(f) v.f = ev ? cfg.eval_##f : cfg.f
  RL_PICK(teleport_robots);
  RL_PICK(teleport_lo_x); RL_PICK(teleport_hi_x); RL_PICK(teleport_shift_x);
  RL_PICK(teleport_lo_y); RL_PICK(teleport_hi_y); RL_PICK(teleport_shift_y);
  RL_PICK(randomize_motor_strength); RL_PICK(randomize_Kp_factor); RL_PICK(randomize_Kd_factor);
  RL_PICK(motor_strength_lo_span[0]); RL_PICK(motor_strength_lo_span[1]);
  RL_PICK(Kp_factor_lo_span[0]); RL_PICK(Kp_factor_lo_span[1]);
  RL_PICK(Kd_factor_lo_span[0]); RL_PICK(Kd_factor_lo_span[1]);
  RL_PICK(push_robots); RL_PICK(push_interval);
  RL_PICK(push_lo_span[0]); RL_PICK(push_lo_span[1]);
#undef RL_PICK
  return v;
}

// legged_robot.py:768-791 for env e of a split population (the range's own thresholds)
__device__ inline bool teleport_xy_env(const RlEnvCfg& cfg, int e, float& x, float& y) {
  const bool ev = has_eval_split(cfg) && e >= cfg.num_train_envs;
  if (!(ev ? cfg.eval_teleport_robots : cfg.teleport_robots)) return false;
  const float lo_x = ev ? cfg.eval_teleport_lo_x : cfg.teleport_lo_x, hi_x = ev ? cfg.eval_teleport_hi_x : cfg.teleport_hi_x;
  const float sh_x = ev ? cfg.eval_teleport_shift_x : cfg.teleport_shift_x;
  const float lo_y = ev ? cfg.eval_teleport_lo_y : cfg.teleport_lo_y, hi_y = ev ? cfg.eval_teleport_hi_y : cfg.teleport_hi_y;
  const float sh_y = ev ? cfg.eval_teleport_shift_y : cfg.teleport_shift_y;
  const float x0 = x, y0 = y;
  if (x < lo_x) x += sh_x;
  if (x > hi_x) x -= sh_x;
  if (y < lo_y) y += sh_y;
  if (y > hi_y) y -= sh_y;
  return x != x0 || y != y0;
}

// legged_robot.py:768-791: the teleport of one env's base position (idempotent: a teleported position lies inside)
__device__ inline bool teleport_xy(const RlEnvCfg& cfg, float& x, float& y) {
  const float x0 = x, y0 = y;
  if (x < cfg.teleport_lo_x) x += cfg.teleport_shift_x;
  if (x > cfg.teleport_hi_x) x -= cfg.teleport_shift_x;
  if (y < cfg.teleport_lo_y) y += cfg.teleport_shift_y;
  if (y > cfg.teleport_hi_y) y -= cfg.teleport_shift_y;
  return x != x0 || y != y0;
}

// heights.cu: the pre-pass launch (b.height_mean set)
int launch_heights_prepass(const StepArgs& args, cudaStream_t st);

// launcher of the one-warp-per-leg kernel for the standard observation layout (env_step_quad.cu)
int launch_step_quad(const StepArgs& args, bool fuse_torques, cudaStream_t st);
// the same step with all tile traffic on TMA, for the shipped configuration on packed state blocks (env_step_rows.cu)
bool rows_layout_ok(const StepArgs& args);
int launch_step_rows(const StepArgs& args, bool fuse_torques, cudaStream_t st);

}  // namespace rl
