// Shared device/host helpers for the rl_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "rl_b200.h"

namespace rl {

// ---- error plumbing (thread-local text behind rl_last_error()) -------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaPeekAtLastError -> RL_ERR_CUDA

#define RL_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ::rl::set_error(__VA_ARGS__);  \
      return (code);                 \
    }                                \
  } while (0)

// ---- Philox4x32-10 counter-based RNG (Salmon et al. 2011) -------------------------------
// Keyed by (seed); counter = (env, step_lo, step_hi, stream<<16 | block).  Every random
// quantity of the env path is addressable without any per-env generator state, so resets,
// noise and curriculum draws are reproducible whatever the launch geometry.
struct Philox {
  static constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u;
  static constexpr uint32_t kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
  __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(kM0, c[0]), hi1 = __umulhi(kM1, c[2]);
#else
    uint32_t hi0 = (uint32_t)(((uint64_t)kM0 * c[0]) >> 32);
    uint32_t hi1 = (uint32_t)(((uint64_t)kM1 * c[2]) >> 32);
#endif
    uint32_t lo0 = kM0 * c[0], lo1 = kM1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __host__ __device__ static inline void gen(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2,
                                             uint32_t c3, uint32_t (&out)[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c[4] = {c0, c1, c2, c3};
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c, k0, k1);
      k0 += kW0; k1 += kW1;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};

enum RngStream : uint32_t {
  RNG_NOISE = 1,    // observation noise (legged_robot.py:392)
  RNG_DR = 2,       // motor/Kp/Kd re-draw (:544-560)
  RNG_PUSH = 3,     // push velocity (:764)
  RNG_RESET = 4,    // reset draws (:727, :813)
  RNG_GAC = 5,      // curriculum sampling (curriculum.py:55-68)
  RNG_POLICY = 6,   // Normal sample in ActorCritic.act (actor_critic.py:144)
};

__host__ __device__ inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
__host__ __device__ inline double u01d(uint32_t hi, uint32_t lo) {
  // 53-bit uniform in [0,1)
  uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)v * (1.0 / 9007199254740992.0);
}

// Four uniforms for (env, step, stream, block).
__device__ inline void rng4(uint64_t seed, uint32_t env, uint64_t step, uint32_t stream,
                            uint32_t block, float (&u)[4]) {
  uint32_t r[4];
  Philox::gen(seed, env, (uint32_t)step, (uint32_t)(step >> 32), (stream << 16) | block, r);
  u[0] = u01(r[0]); u[1] = u01(r[1]); u[2] = u01(r[2]); u[3] = u01(r[3]);
}

// ---- warp helpers ---------------------------------------------------------------------
__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ inline double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 128-bit streaming load / store (read-once data: bypass L1 allocation)
__device__ inline float4 ldg_stream4(const float* p) {
  float4 r;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ inline void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// ---- TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier ---------------------------
__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ inline void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ inline void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion counted in bytes on `bar`; src/dst 16 B aligned, bytes % 16 == 0
__device__ inline void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global (bulk async-group completion)
__device__ inline void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ inline void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ inline void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (before a bulk store)
__device__ inline void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace rl
