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

// launcher of the one-warp-per-leg kernel for the standard observation layout (env_step_quad.cu)
int launch_step_quad(const StepArgs& args, bool fuse_torques, cudaStream_t st);
// the same step with all tile traffic on TMA, for the shipped configuration on packed state blocks (env_step_rows.cu)
bool rows_layout_ok(const StepArgs& args);
int launch_step_rows(const StepArgs& args, bool fuse_torques, cudaStream_t st);

}  // namespace rl
