// Fused MLP chains on the 5th-generation tensor cores: a whole forward (or dgrad) pass of the learner's
// networks (mini_gym_learn/ppo/actor_critic.py:38-100; the autograd backward behind ppo.py:146-168) for
// a 128-row tile inside ONE persistent CTA.  The reference runs each nn.Linear / nn.ELU as its own cuBLAS
// + elementwise launch with every activation going through HBM; rl_gemm_bf16 (gemm_tc.cu) fuses the
// epilogue but still launches per layer.  Here activations stay in shared memory between layers, weights
// stream from L2 through a TMA ring and accumulators stay in tensor memory.
//
// Execution model (see include/rl_b200.h "Fused MLP chains"): three warp roles interpret three host-built
// op lists, in order, once per tile of the persistent loop; every dependency is an mbarrier:
//   warp 0 (1 elected lane)  LOAD ops: mbarrier wait (unit free) -> expect_tx -> cp.async.bulk.tensor.2d
//   warp 1 (1 elected lane)  MMA ops : waits (stage full / box ready / accumulator free) -> <= 4 tcgen05.mma
//                               (M128 x n x K16, kind::f16, fp32 in TMEM) -> tcgen05.commit on <= 3 barriers
//   warps 2-5, 6-9, .. EPI ops : NW (2..4) workers of four warps, each runs its own list: wait (accumulator full) -> tcgen05.ld ->
//                               bias / ELU / ELU' -> swizzled bf16 box in shared memory (the next layer's A operand) -> TMA store
//                               of the box (saved activation / gradient for wgrad) or fp32 output rows.  The four warps of a
//                               worker never synchronise with each other: warp g owns rows [32 g, 32 g + 32) of every box -
//                               it writes them, fences them and TMA-stores them (a [32 x 64] sub-box, own bulk groups);
//                               barriers the program gives ONE arrival per worker are completed by the last of the four
//                               warps (a shared-memory ticket per barrier)
// Shared memory = n_units x 16 KB dynamic (activation boxes and ring stages, all [rows x 64 bf16] tiles in the
// 128 B-swizzled K-major layout shared by TMA and the UMMA descriptors) + 3 KB static (64 mbarriers, the
// per-warp bias staging rows): 14 units use exactly the 227 KB a CTA can have.
// The host side (ppo/chain.py) builds the op lists and proves them on an emulator (deadlock freedom,
// buffer hazards, parity bookkeeping, numerics) before anything reaches the GPU.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tc_common.cuh"
#include "ppo_loss.cuh"

namespace rl {
namespace tc {

constexpr int MAX_WORKERS = 4;
// warp 0 LOAD, warp 1 MMA, four warps per epilogue worker, then the SECOND issuers: one more LOAD warp and one more MMA
// warp (they land on other scheduler partitions than warps 0 / 1).  The per-op cost of an issuing warp (~1100 cycles
// for an MMA op, ~480 for a load, whatever their size: dependent waits, descriptor arithmetic, commits, all on one
// thread next to two busy epilogue warps) was the critical path of the chains (sensitivity study: profiles/r02_chain_ncu.md),
// so the op lists are split over two issuers each - by accumulator chain for the MMAs, by ring stage for the loads.
constexpr int chain_threads(int nw) { return 128 + 128 * nw; }
constexpr int N_ISSUERS = 2;
constexpr int UNIT_BYTES = 16384;
constexpr int FIXED_SMEM = 3072;            // static: 80 mbarriers [0, 640) | tmem slot 640 | 80 tickets [672, 992) | per-warp bias staging (16 x 128 B) from 1024
static_assert(RL_CHAIN_MAX_BARRIERS * 8 <= 640 && 672 + RL_CHAIN_MAX_BARRIERS * 4 <= 1024, "fixed shared-memory layout");
constexpr int MAX_STORE_MAPS = 16;
constexpr int TRACE_MMA = 8, TRACE_EPI = 5; // stamps per op (loads: 1)
#ifndef RL_CHAIN_WAIT_NS_LOAD
#define RL_CHAIN_WAIT_NS_LOAD 0
#endif
#ifndef RL_CHAIN_WAIT_NS_EPI
#define RL_CHAIN_WAIT_NS_EPI 0
#endif
constexpr int WAIT_NS_LOAD = RL_CHAIN_WAIT_NS_LOAD, WAIT_NS_EPI = RL_CHAIN_WAIT_NS_EPI;

// device-side op formats (built by rl_chain_create from the ABI structs)
struct DevMmaOp {            // 32 B; everything the issuing warp would otherwise have to decode is precomputed
  uint32_t a_lo, b_lo;       // low descriptor words without the shared-memory window base: (off >> 4) | LBO field
  uint32_t idesc;
  uint32_t misc;             // tmem_col | k_steps << 16 | accumulate << 24
  uint16_t wait_off[4];      // byte offset of the barrier in the barrier block + 1, or 0 for "no wait"
  uint16_t commit_off[3];    // same for the barriers tcgen05.commit arrives on
  uint16_t parities;         // bit 2j: parity of wait j when the tile iteration is even, bit 2j+1: when it is odd
};
static_assert(sizeof(DevMmaOp) == 32, "DevMmaOp layout");
constexpr int MAX_OPS_PER_ISSUER = 64;      // an issuing warp keeps its op list in registers: 2 ops per lane

// The LOAD and MMA op lists travel in the kernel parameters (constant bank): indexed by the loop counter
// they are read with uniform loads, so the TMA / tcgen05 instructions get their operands in uniform
// registers directly.  (Ops fetched from global memory are per-thread values: every UTCHMMA then sits in
// an ELECT / 7x R2UR.BROADCAST waterfall loop, ~90 cycles per instruction - measured.)
constexpr int MAX_LOADS = 192, MAX_MMAS = 160;
struct ChainParams {
  CUtensorMap tmaps[RL_CHAIN_MAX_TENSORS];
  CUtensorMap tmaps_st[MAX_STORE_MAPS];   // [32 x 64] boxes over the stored tensors (one warp's rows of a box)
  RlChainLoadOp loads[MAX_LOADS];
  DevMmaOp mmas[MAX_MMAS];
  const RlChainEpiOp* epis[MAX_WORKERS];    // per epilogue worker (global memory)
  const float* params;
  float* outputs[RL_CHAIN_MAX_OUTPUTS];
  int n_loads, n_mmas, n_epis[MAX_WORKERS];
  int load_begin[N_ISSUERS + 1], mma_begin[N_ISSUERS + 1];     // ops of issuer k: [begin[k], begin[k + 1]) (lists sorted by issuer)
  int n_units, n_barriers;
  int num_tiles, rows;           // tiles [tile0, num_tiles) of the `rows`-row batch
  int tile0;
  int dbg_sleep[3];              // sensitivity study (RL_CHAIN_DBG_SLEEP="load,mma,epi" ns per op): where is the critical path?
  unsigned long long* trace;      // profiling aid: clock64 stamps of CTA 0 in tile iteration trace_it (or null)
  int trace_it;
  uint8_t barrier_count[RL_CHAIN_MAX_BARRIERS];
  // PPO loss fused into the value-output epilogue (rl_chain_set_ppo_loss): loss.mean = outputs[mean_out]
  LossArgs loss;
  int loss_value_out;             // outputs[] index whose epilogue op evaluates the loss, or -1
};

// mbarrier wait with a watchdog: a schedule bug must surface as a launch error, never as a hung GPU
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_addr(uint32_t bar_smem_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(bar_smem_addr), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __noinline__ void chain_timeout(uint32_t id, uint32_t parity, int it) {
  printf("mlp_chain_kernel: wait on barrier %u (parity %u, tile iteration %d) timed out, block %d thread %d\n", id, parity, it,
         (int)blockIdx.x, (int)threadIdx.x);
  __trap();
}
// SLEEP_NS > 0: back off between polls (tried: 64 ns in the LOAD warp, 32 ns in the epilogue workers - no gain,
// the dgrad chain got slower; kept as a compile-time knob, default off) so that a waiting warp does not take issue slots from the warps that
// share its scheduler (the LOAD warp and the epilogue workers are not latency critical; the MMA warp is)
template <int SLEEP_NS = 0>
__device__ __forceinline__ void chain_wait(uint64_t* bars, uint32_t spec, int it) {
  const uint32_t id = spec & 0xFFu;
  if (id == RL_CHAIN_NONE) return;
  const uint32_t parity = ((spec >> 8) ^ ((spec >> 9) & (uint32_t)it)) & 1u;
  uint32_t spins = 0;
  while (!mbar_try(&bars[id], parity)) {
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
    if (++spins > (1u << 24)) chain_timeout(id, parity, it);
  }
}

// up to three waits at once: the try_waits are issued back to back (independent), only failed ones are retried
__device__ __forceinline__ void chain_wait3(uint64_t* bars, uint32_t s0, uint32_t s1, uint32_t s2, int it) {
  const uint32_t i0 = s0 & 0xFFu, i1 = s1 & 0xFFu, i2 = s2 & 0xFFu;
  const uint32_t p0 = ((s0 >> 8) ^ ((s0 >> 9) & (uint32_t)it)) & 1u, p1 = ((s1 >> 8) ^ ((s1 >> 9) & (uint32_t)it)) & 1u,
                 p2 = ((s2 >> 8) ^ ((s2 >> 9) & (uint32_t)it)) & 1u;
  bool ok0 = i0 == RL_CHAIN_NONE || mbar_try(&bars[i0], p0);
  bool ok1 = i1 == RL_CHAIN_NONE || mbar_try(&bars[i1], p1);
  bool ok2 = i2 == RL_CHAIN_NONE || mbar_try(&bars[i2], p2);
  uint32_t spins = 0;
  while (!(ok0 && ok1 && ok2)) {
    if (!ok0) ok0 = mbar_try(&bars[i0], p0);
    if (!ok1) ok1 = mbar_try(&bars[i1], p1);
    if (!ok2) ok2 = mbar_try(&bars[i2], p2);
    if (++spins > (1u << 24)) chain_timeout(!ok0 ? i0 : (!ok1 ? i1 : i2), 2, it);
  }
}

// ELU on the fp32 accumulator: exp through ex2.approx.ftz (MUFU), 4 instructions per element
__device__ __forceinline__ float elu1(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
  return x > 0.f ? x : e - 1.f;
}

// tanh on the fp32 accumulator: one MUFU.TANH (high_level_policy networks)
__device__ __forceinline__ float tanh1(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int N>
__device__ __forceinline__ void bulk_wait_read_n() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_read_dyn(int n) {
  switch (n) {
    case 0: bulk_wait_read_n<0>(); break;
    case 1: bulk_wait_read_n<1>(); break;
    case 2: bulk_wait_read_n<2>(); break;
    case 3: bulk_wait_read_n<3>(); break;
    case 4: bulk_wait_read_n<4>(); break;
    case 5: bulk_wait_read_n<5>(); break;
    case 6: bulk_wait_read_n<6>(); break;
    default: bulk_wait_read_n<7>(); break;
  }
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// byte offset of the 16 B chunk holding columns [8c, 8c+8) of row r inside a 128 B-swizzled [rows x 64] box
__device__ __forceinline__ uint32_t sw_chunk(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// one lane of a converged warp (CUTLASS elect_one_sync)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mma_issue(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  // descriptor high word: SBO = 1024 B (8-row group pitch), version 1, SWIZZLE_128B
  constexpr uint32_t HI = 64u | (1u << 14) | (2u << 29);
  const uint64_t da = ((uint64_t)HI << 32) | a_lo, db = ((uint64_t)HI << 32) | b_lo;
  mma_bf16_ss(tmem_d, da, db, idesc, accumulate != 0);
}

__device__ __forceinline__ void mma_wait(uint32_t bar_addr, uint32_t parity, int it) {
  if (mbar_try_addr(bar_addr, parity)) return;
  uint32_t spins = 0;
  while (!mbar_try_addr(bar_addr, parity)) {
    if (++spins > (1u << 24)) chain_timeout((bar_addr & 0x3FFu) / 8, parity, it);
  }
}

// one arrival per WORKER on a barrier every warp of the worker is done with: the last of the four warps arrives
// (lane 0 of each warp calls this after its own reads / writes / bulk reads are complete)
__device__ __forceinline__ void worker_arrive(uint64_t* bars, uint32_t* tickets, uint32_t bar_id) {
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(tickets + bar_id)) : "memory");
  if ((old & 3u) == 3u) mbar_arrive(&bars[bar_id]);
}

// PPO loss of row r (ppo.py:110-144) by the epilogue thread that has just written value[r] (= v) and, earlier in its op
// list, mean[r]: dmean / dvalue rows, then one set of atomics per warp for the loss statistics and d loss / d std.
// Called by whole warps.  (Not inlined: the epilogue's 64 accumulator registers are not live across it.)
__device__ __noinline__ void chain_ppo_loss(const LossArgs& a, int r, int rows, float v) {
  double s_surr = 0, s_val = 0, s_kl = 0;
  float g_std[ACT];
#pragma unroll
  for (int d = 0; d < ACT; ++d) g_std[d] = 0.f;
  const bool valid = r < rows;
  if (valid) ppo_loss_row(a, r, v, s_surr, s_val, s_kl, g_std);
  const int lane = threadIdx.x & 31;
  const int n_valid = __popc(__ballot_sync(0xFFFFFFFFu, valid));
  if (n_valid == 0) return;
  s_surr = warp_sum(s_surr); s_val = warp_sum(s_val); s_kl = warp_sum(s_kl);
  float mine = 0.f;
#pragma unroll
  for (int d = 0; d < ACT; ++d) {
    const float t = warp_sum(g_std[d]);
    if (lane == d) mine = t;
  }
  if (lane == 0) {
    atomicAdd(a.stats + 0, s_surr);
    atomicAdd(a.stats + 1, s_val);
    atomicAdd(a.stats + 2, s_kl);
    if (a.kl_slot) atomicAdd(a.kl_slot, (float)s_kl);
  }
  if (lane < ACT) {
    // entropy (:144): mean over rows of sum_d (0.5 + 0.5 log 2pi + log std_d)  =>  d/dstd_d = 1/std_d per row
    mine += -a.entropy_coef * (1.f / a.std[lane]) * (a.inv_global_B * (float)n_valid);
    atomicAdd(a.dstd + lane, mine);
  }
}

template <bool TRACE, int NW>
__global__ void __launch_bounds__(chain_threads(NW), 1)
mlp_chain_kernel(const __grid_constant__ ChainParams p) {
  __shared__ __align__(1024) uint8_t s_fixed[FIXED_SMEM];
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_fixed);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_fixed + 640);
  uint32_t* tickets = reinterpret_cast<uint32_t*>(s_fixed + 672);          // one per barrier
  float* s_bias = reinterpret_cast<float*>(s_fixed + 1024);                // 16 warps x 32 floats
  // warp-uniform role index (the shuffle tells the compiler so: the LOAD / MMA loops run on the uniform datapath)
  const int warp = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = smem_u32(smem);

  if (threadIdx.x == 0) {
    if (smem_base & 1023u) {
      printf("mlp_chain_kernel: dynamic shared memory is not 1024 B aligned (0x%x)\n", smem_base);
      __trap();
    }
    for (int b = 0; b < p.n_barriers; ++b) mbar_init(&bars[b], p.barrier_count[b]);
    mbar_fence_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + RL_CHAIN_MAX_BARRIERS) tickets[threadIdx.x - 64] = 0;
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int issuer2_warp0 = 2 + 4 * NW;        // warps issuer2_warp0 (LOAD) and issuer2_warp0 + 1 (MMA): the second issuers
  if (warp == 0 || warp == issuer2_warp0) {
    // ===================================== LOAD role =====================================
    const int k = warp == 0 ? 0 : 1;
    const int i0 = p.load_begin[k], i1 = p.load_begin[k + 1];
    const bool leader = elect_one();
    int it = 0;
    for (int tile = p.tile0 + blockIdx.x; tile < p.num_tiles && i1 > i0; tile += gridDim.x, ++it) {
      const int m0 = tile * 128;
      for (int i = i0; i < i1; ++i) {
        const RlChainLoadOp& cur = p.loads[i];
        if (p.dbg_sleep[0]) __nanosleep(p.dbg_sleep[0]);
        chain_wait<WAIT_NS_LOAD>(bars, cur.wait, it);
        if (leader) {
          mbar_expect_tx(&bars[cur.full_bar], cur.expect_bytes);
          tma_load_2d(smem + cur.smem_off, &p.tmaps[cur.tensor], cur.col0, cur.row0 + (cur.tile_rows ? m0 : 0), &bars[cur.full_bar]);
          if (TRACE && p.trace && blockIdx.x == 0 && it == p.trace_it) p.trace[i] = clock64();
        }
      }
    }
  } else if (warp == 1 || warp == issuer2_warp0 + 1) {
    // ===================================== MMA role ======================================
    const int k = warp == 1 ? 0 : 1;
    const int i0 = p.mma_begin[k], i1 = p.mma_begin[k + 1];
    const bool leader = elect_one();
    const uint32_t base16 = smem_base >> 4;
    uint32_t bar0 = smem_u32(bars);
    asm volatile("" : "+r"(bar0));      // (opaque: otherwise the address is re-derived - S2R SR_CgaCtaId - at every use)
    // The op list lives in the kernel parameters; an indexed constant load costs ~100 cycles and an op needs several
    // dependent ones.  The list is the same for every tile, so the warp keeps it in registers: lane l holds ops
    // i0 + l and i0 + 32 + l (8 words each, <= 64 ops per issuer) and hands an op's words out by shuffle.
    uint32_t mine[2][8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = i0 + 32 * h + lane;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(&p.mmas[j < i1 ? j : (i1 > i0 ? i1 - 1 : 0)]);
#pragma unroll
      for (int q = 0; q < 8; ++q) mine[h][q] = (i1 > i0) ? src[q] : 0u;
    }
    int it = 0;
    for (int tile = p.tile0 + blockIdx.x; tile < p.num_tiles && i1 > i0; tile += gridDim.x, ++it) {
      const bool tr = TRACE && p.trace && blockIdx.x == 0 && it == p.trace_it;
      for (int i = i0; i < i1; ++i) {
        const int rel = i - i0, srcl = rel & 31;
        uint32_t w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) w[q] = __shfl_sync(0xFFFFFFFFu, rel < 32 ? mine[0][q] : mine[1][q], srcl);
        // words: a_lo | b_lo | idesc | misc | wait_off 0,1 | wait_off 2,3 | commit_off 0,1 | commit_off 2, parities
        if (p.dbg_sleep[1]) __nanosleep(p.dbg_sleep[1]);
        if (tr && leader) p.trace[p.n_loads + TRACE_MMA * i + 1] = clock64();
        {
          // (sequential spins: all of them must pass anyway and the later ones are normally complete by then)
          const uint32_t par = (w[7] >> 16) >> (it & 1);
          const uint32_t w0 = w[4] & 0xFFFFu, w1 = w[4] >> 16, w2 = w[5] & 0xFFFFu, w3 = w[5] >> 16;
          if (w0) mma_wait(bar0 + w0 - 1, par & 1u, it);
          if (w1) mma_wait(bar0 + w1 - 1, (par >> 2) & 1u, it);
          if (w2) mma_wait(bar0 + w2 - 1, (par >> 4) & 1u, it);
          if (w3) mma_wait(bar0 + w3 - 1, (par >> 6) & 1u, it);
        }
        tc_fence_after();
        if (leader) {
          if (tr) p.trace[p.n_loads + TRACE_MMA * i] = clock64();
          const uint32_t a_lo = w[0] + base16, b_lo = w[1] + base16, idesc = w[2], misc = w[3];
          const uint32_t tmem_d = tmem_base + (misc & 0xFFFFu);
          const uint32_t k_steps = (misc >> 16) & 0xFFu;
          mma_issue(tmem_d, a_lo, b_lo, idesc, misc >> 24);                     // K16 step 0
          if (k_steps > 1) mma_issue(tmem_d, a_lo + 2, b_lo + 2, idesc, 1);     // +32 B per step
          if (k_steps > 2) mma_issue(tmem_d, a_lo + 4, b_lo + 4, idesc, 1);
          if (k_steps > 3) mma_issue(tmem_d, a_lo + 6, b_lo + 6, idesc, 1);
          if (tr) p.trace[p.n_loads + TRACE_MMA * i + 4] = clock64();
          const uint32_t c0 = w[6] & 0xFFFFu, c1 = w[6] >> 16, c2 = w[7] & 0xFFFFu;
          if (c0) mma_commit(nullptr, bar0 + c0 - 1);
          if (c1) mma_commit(nullptr, bar0 + c1 - 1);
          if (c2) mma_commit(nullptr, bar0 + c2 - 1);
          if (tr) p.trace[p.n_loads + TRACE_MMA * i + 7] = clock64();
        }
      }
    }
  } else {
    // ===================================== EPILOGUE workers ==============================
    const int worker = (warp - 2) >> 2;          // warps 2-5: worker 0, warps 6-9: worker 1, ...
    const int g = warp & 3;                      // TMEM lane quarter this warp may read = the box rows it owns
    const int lr = 32 * g + lane;                // row within the tile
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * g) << 16);
    float* my_bias = s_bias + (warp - 2) * 32;
    const int n_ops = p.n_epis[worker];
    const uint4* ops = reinterpret_cast<const uint4*>(p.epis[worker]);
    int trace_base = p.n_loads + TRACE_MMA * p.n_mmas;
    for (int k = 0; k < worker; ++k) trace_base += TRACE_EPI * p.n_epis[k];
    const bool first = (warp - 2) == 4 * worker && lane == 0;     // the worker's trace stamps come from its first thread
    int it = 0;
    for (int tile = p.tile0 + blockIdx.x; tile < p.num_tiles && n_ops > 0; tile += gridDim.x, ++it) {
      const int m0 = tile * 128;
      const int r = m0 + lr;
      // (two / three workers: the next op's words are fetched one op ahead; four workers have no registers to spare)
      uint4 n0 = __ldg(ops), n1 = __ldg(ops + 1), n2 = __ldg(ops + 2);
#pragma unroll 1
      for (int i = 0; i < n_ops; ++i) {
        if (NW == 4 && i > 0) { n0 = __ldg(ops + 3 * i); n1 = __ldg(ops + 3 * i + 1); n2 = __ldg(ops + 3 * i + 2); }
        const uint4 w0 = n0, w1 = n1, w2 = n2;
        if (NW != 4 && i + 1 < n_ops) { n0 = __ldg(ops + 3 * (i + 1)); n1 = __ldg(ops + 3 * (i + 1) + 1); n2 = __ldg(ops + 3 * (i + 1) + 2); }
        const uint32_t wait_acc = w0.x & 0xFFFFu, wait_dst = w0.x >> 16, wait_aux = w0.y & 0xFFFFu;
        const uint32_t arrive_acc_free = (w0.y >> 16) & 0xFFu, arrive_dst_ready = w0.y >> 24;
        const uint32_t release_aux = w0.z & 0xFFu, mode = (w0.z >> 8) & 0xFFu;
        const int ncols = (int)((w0.z >> 16) & 0xFFu), dst_col0 = (int)(w0.z >> 24);
        const uint32_t tmem_col = w0.w & 0xFFFFu, store_tensor = (w0.w >> 16) & 0xFFu;
        const int store_wait_pending = (int)(int8_t)(w0.w >> 24);
        const uint32_t release_after_store = w1.x & 0xFFu, out_id = (w1.x >> 8) & 0xFFu, out_ld = w1.x >> 16;
        const uint32_t bias_off = w1.y, dst_off = w1.z, aux_off = w1.w;
        const int store_col0 = (int)w2.x;
        if (w2.z) __nanosleep(w2.z);            // delay_ns
        if (p.dbg_sleep[2]) __nanosleep(p.dbg_sleep[2]);
        const bool has_bias = mode == RL_CHAIN_EPI_BIAS_ELU || mode == RL_CHAIN_EPI_BIAS || mode == RL_CHAIN_EPI_BIAS_F32 ||
                              mode == RL_CHAIN_EPI_BIAS_TANH;

        const bool tr = TRACE && p.trace && blockIdx.x == 0 && it == p.trace_it && first;
        unsigned long long* tp = p.trace + trace_base + TRACE_EPI * i;
        if (tr) tp[0] = clock64();
        // bias of this op's columns: two coalesced loads per warp, issued before the accumulator wait
        float b_lo = 0.f, b_hi = 0.f;
        if (has_bias) {
          const float* bias = p.params + bias_off;
          if (lane < ncols) b_lo = __ldg(bias + lane);
          if (lane + 32 < ncols) b_hi = __ldg(bias + 32 + lane);
        }
        if constexpr (NW == 4) {
          // ---- four workers: 32 accumulator columns at a time (<= 96 registers per thread without spills) ----
          chain_wait<WAIT_NS_EPI>(bars, wait_acc, it);
          tc_fence_after();
#ifndef RL_CHAIN_TRACE_WRITE
          if (tr) tp[1] = clock64();
#endif
          const int nh = ncols > 32 ? 2 : 1;
          uint8_t* box = smem + dst_off;
#pragma unroll 1
          for (int h = 0; h < nh; ++h) {
            float f[32];
            {
              uint32_t v[32];
              tmem_ld32(lane_addr + tmem_col + 32 * h, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            }
            if (h == nh - 1 && arrive_acc_free != RL_CHAIN_NONE) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&bars[arrive_acc_free]);
            }
#ifndef RL_CHAIN_TRACE_WRITE
            if (tr && h == 0) tp[2] = clock64();
#endif
            if (has_bias) {
              __syncwarp();                            // the previous round's broadcast reads are done
              my_bias[lane] = h ? b_hi : b_lo;
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 bb = *reinterpret_cast<const float4*>(my_bias + j);
                f[j] += bb.x; f[j + 1] += bb.y; f[j + 2] += bb.z; f[j + 3] += bb.w;
              }
              if (mode == RL_CHAIN_EPI_BIAS_ELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = elu1(f[j]);
              } else if (mode == RL_CHAIN_EPI_BIAS_TANH) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = tanh1(f[j]);
              }
            } else if (mode == RL_CHAIN_EPI_DELU || mode == RL_CHAIN_EPI_DTANH) {
              if (h == 0) chain_wait<WAIT_NS_EPI>(bars, wait_aux, it);
              const uint8_t* aux = smem + aux_off;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint4 pk = *reinterpret_cast<const uint4*>(aux + sw_chunk(lr, 4 * h + c));
                const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float2 y = __bfloat1622float2(hh[q]);
                  if (mode == RL_CHAIN_EPI_DTANH) {
                    f[c * 8 + 2 * q] *= 1.f - y.x * y.x;
                    f[c * 8 + 2 * q + 1] *= 1.f - y.y * y.y;
                  } else {
                    f[c * 8 + 2 * q] *= (y.x > 0.f) ? 1.f : (y.x + 1.f);
                    f[c * 8 + 2 * q + 1] *= (y.y > 0.f) ? 1.f : (y.y + 1.f);
                  }
                }
              }
            }
            if (mode == RL_CHAIN_EPI_BIAS_F32) {
              if (r < p.rows) {
                float* dst = p.outputs[out_id] + (size_t)r * out_ld;
#pragma unroll
                for (int j = 0; j < 32; ++j) if (j < ncols) dst[j] = f[j];
              }
              break;
            }
            if (h == 0) {
#ifndef RL_CHAIN_TRACE_WRITE
              if (tr) tp[3] = clock64();
#else
              if (tr) tp[0] = clock64();
#endif
              if (store_wait_pending >= 0) {          // a TMA store this warp issued earlier may still be reading its rows
                if (lane == 0) bulk_wait_read_dyn(store_wait_pending);
                __syncwarp();
              }
#ifdef RL_CHAIN_TRACE_WRITE
              if (tr) tp[1] = clock64();
#endif
              chain_wait<WAIT_NS_EPI>(bars, wait_dst, it);
#ifdef RL_CHAIN_TRACE_WRITE
              if (tr) tp[2] = clock64();
#endif
            }
            if (dst_col0 == 0 && ncols == 64) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(f[c * 8], f[c * 8 + 1]), p1 = __floats2bfloat162_rn(f[c * 8 + 2], f[c * 8 + 3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(f[c * 8 + 4], f[c * 8 + 5]), p3 = __floats2bfloat162_rn(f[c * 8 + 6], f[c * 8 + 7]);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
                pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
                *reinterpret_cast<uint4*>(box + sw_chunk(lr, 4 * h + c)) = pk;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (j < ncols) {
                  const int col = dst_col0 + j;
                  *reinterpret_cast<__nv_bfloat16*>(box + sw_chunk(lr, col >> 3) + (col & 7) * 2) = __float2bfloat16(f[j]);
                }
              }
            }
          }
          if (mode == RL_CHAIN_EPI_BIAS_F32) {
            if (tr) tp[2] = tp[3] = tp[4] = clock64();
            continue;
          }
#ifdef RL_CHAIN_TRACE_WRITE
          if (tr) tp[3] = clock64();
#endif
          fence_async_smem();                      // generic-proxy writes -> visible to tcgen05.mma / TMA
          __syncwarp();
          if (lane == 0) {
            if (arrive_dst_ready != RL_CHAIN_NONE) mbar_arrive(&bars[arrive_dst_ready]);
            if (release_aux != RL_CHAIN_NONE) worker_arrive(bars, tickets, release_aux);
            if (store_tensor != RL_CHAIN_NONE) {
              if (m0 + 32 * g < p.rows) tma_store_2d(&p.tmaps_st[store_tensor], box + 4096 * g, store_col0, m0 + 32 * g);
              bulk_commit();
              if (release_after_store != RL_CHAIN_NONE) {
                bulk_wait_read_n<0>();
                worker_arrive(bars, tickets, release_after_store);
              }
            }
          }
        } else {
        // ---- accumulator columns -> registers ----
        chain_wait<WAIT_NS_EPI>(bars, wait_acc, it);
        tc_fence_after();
        if (tr) tp[1] = clock64();
        float f[64];
        {
          uint32_t v[32];
          tmem_ld32(lane_addr + tmem_col, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (ncols > 32) {
            tmem_ld32(lane_addr + tmem_col + 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) f[32 + j] = __uint_as_float(v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[32 + j] = 0.f;
          }
        }
        if (arrive_acc_free != RL_CHAIN_NONE) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[arrive_acc_free]);
        }
        if (tr) tp[2] = clock64();

        // ---- elementwise ----
        if (has_bias) {
          // (the per-warp staging row holds 32 columns: two rounds)
          __syncwarp();                            // the previous op's broadcast reads are done
          my_bias[lane] = b_lo;
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(my_bias + j);
            f[j] += bb.x; f[j + 1] += bb.y; f[j + 2] += bb.z; f[j + 3] += bb.w;
          }
          if (ncols > 32) {
            __syncwarp();
            my_bias[lane] = b_hi;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(my_bias + j);
              f[32 + j] += bb.x; f[33 + j] += bb.y; f[34 + j] += bb.z; f[35 + j] += bb.w;
            }
          }
          if (mode == RL_CHAIN_EPI_BIAS_ELU) {
#pragma unroll
            for (int j = 0; j < 64; ++j) f[j] = elu1(f[j]);
          } else if (mode == RL_CHAIN_EPI_BIAS_TANH) {
#pragma unroll
            for (int j = 0; j < 64; ++j) f[j] = tanh1(f[j]);
          }
        } else if (mode == RL_CHAIN_EPI_DELU || mode == RL_CHAIN_EPI_DTANH) {
          chain_wait<WAIT_NS_EPI>(bars, wait_aux, it);
          const uint8_t* aux = smem + aux_off;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 pk = *reinterpret_cast<const uint4*>(aux + sw_chunk(lr, c));
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 y = __bfloat1622float2(h[q]);
              if (mode == RL_CHAIN_EPI_DTANH) {
                f[c * 8 + 2 * q] *= 1.f - y.x * y.x;
                f[c * 8 + 2 * q + 1] *= 1.f - y.y * y.y;
              } else {
                f[c * 8 + 2 * q] *= (y.x > 0.f) ? 1.f : (y.x + 1.f);
                f[c * 8 + 2 * q + 1] *= (y.y > 0.f) ? 1.f : (y.y + 1.f);
              }
            }
          }
        }

        if (mode == RL_CHAIN_EPI_BIAS_F32) {
          // ---- fp32 output rows (network outputs: 12 / 1 / 18 columns) ----
          if (r < p.rows) {
            float* dst = p.outputs[out_id] + (size_t)r * out_ld;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncols) dst[j] = f[j];
          }
          if ((int)out_id == p.loss_value_out) chain_ppo_loss(p.loss, r, p.rows, f[0]);
          if (tr) tp[3] = tp[4] = clock64();
          continue;
        }
#ifndef RL_CHAIN_TRACE_WRITE
        if (tr) tp[3] = clock64();
#endif

        // ---- this warp's 32 rows of the bf16 box for the next layer ----
        if (store_wait_pending >= 0) {          // a TMA store this warp issued earlier may still be reading its rows
          if (lane == 0) bulk_wait_read_dyn(store_wait_pending);
          __syncwarp();
        }
#ifdef RL_CHAIN_TRACE_WRITE
        if (tr) tp[0] = clock64();
#endif
        chain_wait<WAIT_NS_EPI>(bars, wait_dst, it);
#ifdef RL_CHAIN_TRACE_WRITE
        if (tr) tp[1] = clock64();
#endif
        uint8_t* box = smem + dst_off;
        if (dst_col0 == 0 && ncols == 64) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(f[c * 8], f[c * 8 + 1]), p1 = __floats2bfloat162_rn(f[c * 8 + 2], f[c * 8 + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(f[c * 8 + 4], f[c * 8 + 5]), p3 = __floats2bfloat162_rn(f[c * 8 + 6], f[c * 8 + 7]);
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
            pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(box + sw_chunk(lr, c)) = pk;
          }
        } else {
          // partial box: `ncols` (<= 32) columns starting at dst_col0 (latent merged next to the observations,
          // or a narrow layer); columns outside the range keep their content
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < ncols) {
              const int col = dst_col0 + j;
              *reinterpret_cast<__nv_bfloat16*>(box + sw_chunk(lr, col >> 3) + (col & 7) * 2) = __float2bfloat16(f[j]);
            }
          }
        }
#ifdef RL_CHAIN_TRACE_WRITE
        if (tr) tp[2] = clock64();
#endif
        fence_async_smem();                      // generic-proxy writes -> visible to tcgen05.mma / TMA
        __syncwarp();
#ifdef RL_CHAIN_TRACE_WRITE
        if (tr) tp[3] = clock64();
#endif
        if (lane == 0) {
          if (arrive_dst_ready != RL_CHAIN_NONE) mbar_arrive(&bars[arrive_dst_ready]);
          if (release_aux != RL_CHAIN_NONE) worker_arrive(bars, tickets, release_aux);
          if (store_tensor != RL_CHAIN_NONE) {
            // (a sub-box that starts below the last row is skipped; the - then empty - group keeps the count)
            if (m0 + 32 * g < p.rows) tma_store_2d(&p.tmaps_st[store_tensor], box + 4096 * g, store_col0, m0 + 32 * g);
            bulk_commit();
            if (release_after_store != RL_CHAIN_NONE) {
              bulk_wait_read_n<0>();
              worker_arrive(bars, tickets, release_after_store);
            }
          }
        }
        }
        if (tr) tp[4] = clock64();
      }
    }
    if (lane == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

struct ChainHandle {
  ChainParams params;
  int n_workers;
  int out_worker[RL_CHAIN_MAX_OUTPUTS], out_pos[RL_CHAIN_MAX_OUTPUTS], out_ld[RL_CHAIN_MAX_OUTPUTS];   // the epilogue op writing outputs[i]: worker, index in the host list (-1: none)
  void* dev_ops;
  size_t smem_bytes;
  unsigned long long* trace;
  size_t trace_len;
};

}  // namespace tc
}  // namespace rl

using namespace rl;
using namespace rl::tc;

extern "C" int rl_chain_create(const RlChainDesc* d, void** handle) {
  RL_REQUIRE(d && handle, RL_ERR_BAD_ARG, "rl_chain_create: null argument");
  RL_REQUIRE(d->n_tensors >= 0 && d->n_tensors <= RL_CHAIN_MAX_TENSORS, RL_ERR_BAD_ARG, "rl_chain_create: n_tensors=%d", d->n_tensors);
  RL_REQUIRE(d->n_units > 0 && d->n_units <= RL_CHAIN_MAX_UNITS, RL_ERR_BAD_ARG, "rl_chain_create: n_units=%d", d->n_units);
  RL_REQUIRE(d->n_barriers > 0 && d->n_barriers <= RL_CHAIN_MAX_BARRIERS, RL_ERR_BAD_ARG, "rl_chain_create: n_barriers=%d", d->n_barriers);
  RL_REQUIRE(d->n_loads >= 0 && d->n_mmas >= 0 && d->n_epis >= 0 && d->loads_host && d->mmas_host && d->epis_host,
             RL_ERR_BAD_ARG, "rl_chain_create: op lists missing");
  RL_REQUIRE(d->n_loads <= MAX_LOADS && d->n_mmas <= MAX_MMAS, RL_ERR_BAD_ARG,
             "rl_chain_create: %d load / %d mma ops exceed the %d / %d that fit in the kernel parameters", d->n_loads, d->n_mmas,
             MAX_LOADS, MAX_MMAS);
  static_assert(sizeof(RlChainLoadOp) == 32 && sizeof(RlChainMmaOp) == 32 && sizeof(RlChainEpiOp) == 48, "op layouts");
  const uint32_t limit = (uint32_t)d->n_units * UNIT_BYTES;
  for (int i = 0; i < d->n_loads; ++i) {
    const RlChainLoadOp& o = d->loads_host[i];
    RL_REQUIRE(o.issuer < N_ISSUERS && (i == 0 || o.issuer >= d->loads_host[i - 1].issuer), RL_ERR_BAD_ARG,
               "rl_chain_create: load op %d: issuer %d (two issuers, list sorted by issuer)", i, (int)o.issuer);
    RL_REQUIRE(o.tensor < d->n_tensors && o.full_bar < d->n_barriers && (o.smem_off & 1023u) == 0 &&
               o.smem_off + o.expect_bytes <= limit && o.expect_bytes > 0 && o.expect_bytes <= 2 * UNIT_BYTES,
               RL_ERR_BAD_ARG, "rl_chain_create: load op %d malformed", i);
  }
  for (int i = 0; i < d->n_mmas; ++i) {
    const RlChainMmaOp& o = d->mmas_host[i];
    RL_REQUIRE(o.issuer < N_ISSUERS && (i == 0 || o.issuer >= d->mmas_host[i - 1].issuer), RL_ERR_BAD_ARG,
               "rl_chain_create: mma op %d: issuer %d (two issuers, list sorted by issuer)", i, (int)o.issuer);
    RL_REQUIRE(o.n >= 16 && o.n <= 256 && (o.n % 16) == 0 && o.tmem_col + o.n <= 512 && o.k_steps >= 1 && o.k_steps <= 4 &&
               (o.a_off & 1023u) == 0 && (o.b_off & 1023u) == 0 && o.a_off + UNIT_BYTES <= limit && o.b_off + (uint32_t)o.n * 128 <= limit,
               RL_ERR_BAD_ARG, "rl_chain_create: mma op %d malformed", i);
  }
  {
    int nm[N_ISSUERS] = {0, 0};
    for (int i = 0; i < d->n_mmas; ++i) ++nm[d->mmas_host[i].issuer];
    RL_REQUIRE(nm[0] <= MAX_OPS_PER_ISSUER && nm[1] <= MAX_OPS_PER_ISSUER, RL_ERR_BAD_ARG,
               "rl_chain_create: %d / %d mma ops per issuer (max %d each: an issuing warp keeps its list in registers)", nm[0], nm[1],
               MAX_OPS_PER_ISSUER);
  }
  int n_epi_w[MAX_WORKERS] = {0, 0, 0, 0};
  int n_workers = 2;
  for (int i = 0; i < d->n_epis; ++i) {
    const RlChainEpiOp& o = d->epis_host[i];
    RL_REQUIRE(o.worker < MAX_WORKERS, RL_ERR_BAD_ARG, "rl_chain_create: epilogue op %d: worker %d (the kernel has at most %d)", i,
               (int)o.worker, MAX_WORKERS);
    if (o.worker + 1 > n_workers) n_workers = o.worker + 1;
    RL_REQUIRE(o.ncols <= 32 || (o.dst_col0 == 0 && o.ncols == 64) || o.mode == RL_CHAIN_EPI_BIAS_F32, RL_ERR_BAD_ARG,
               "rl_chain_create: epilogue op %d: partial boxes hold <= 32 columns", i);
    RL_REQUIRE(o.mode != RL_CHAIN_EPI_BIAS_F32 || o.ncols <= 32, RL_ERR_BAD_ARG, "rl_chain_create: epilogue op %d: <= 32 output columns", i);
    ++n_epi_w[o.worker];
    RL_REQUIRE(o.ncols >= 1 && o.ncols <= 64 && o.tmem_col + (o.ncols > 32 ? 64 : 32) <= 512 && o.mode <= RL_CHAIN_EPI_DTANH, RL_ERR_BAD_ARG,
               "rl_chain_create: epilogue op %d malformed", i);
    if (o.mode == RL_CHAIN_EPI_BIAS_F32)
      RL_REQUIRE(o.out_id < RL_CHAIN_MAX_OUTPUTS && d->outputs[o.out_id] != nullptr, RL_ERR_BAD_ARG, "rl_chain_create: epilogue op %d output", i);
    else
      RL_REQUIRE((o.dst_off & 1023u) == 0 && o.dst_off + UNIT_BYTES <= limit && o.dst_col0 + o.ncols <= 64, RL_ERR_BAD_ARG,
                 "rl_chain_create: epilogue op %d destination", i);
    RL_REQUIRE(o.store_tensor == RL_CHAIN_NONE || o.store_tensor < d->n_tensors, RL_ERR_BAD_ARG, "rl_chain_create: epilogue op %d store", i);
  }
  ChainHandle* h = new ChainHandle();
  memset(&h->params, 0, sizeof(h->params));
  int rc;
  for (int t = 0; t < d->n_tensors; ++t) {
    const RlChainTensor& T = d->tensors[t];
    if ((rc = make_tmap_bf16(&h->params.tmaps[t], T.base, (uint64_t)T.rows, (uint64_t)T.cols, (uint64_t)T.ld, (uint32_t)T.box_rows)) != RL_OK) {
      delete h;
      return rc;
    }
  }
  // device formats: MMA ops with precomputed descriptor words, epilogue ops split per worker
  const size_t ne = (size_t)d->n_epis * sizeof(RlChainEpiOp);
  std::vector<uint8_t> blob(ne + 64, 0);
  for (int i = 0; i < d->n_loads; ++i) h->params.loads[i] = d->loads_host[i];
  DevMmaOp* dm = h->params.mmas;
  for (int i = 0; i < d->n_mmas; ++i) {
    const RlChainMmaOp& o = d->mmas_host[i];
    DevMmaOp& x = dm[i];
    x.a_lo = (o.a_off >> 4) | (1u << 16);        // LBO field = 1 (unused for swizzled K-major)
    x.b_lo = (o.b_off >> 4) | (1u << 16);
    x.idesc = instr_desc_bf16(128, o.n, false, false);
    x.misc = (uint32_t)o.tmem_col | ((uint32_t)o.k_steps << 16) | ((uint32_t)(o.accumulate != 0) << 24);
    const uint16_t ws[4] = {o.wait0, o.wait1, o.wait2, o.wait3};
    const uint8_t cs[3] = {o.commit0, o.commit1, o.commit2};
    uint32_t parities = 0;
    for (int j = 0; j < 4; ++j) {
      const uint32_t id = ws[j] & 0xFFu, base = (ws[j] >> 8) & 1u, flip = (ws[j] >> 9) & 1u;
      x.wait_off[j] = id == RL_CHAIN_NONE ? 0 : (uint16_t)(8 * id + 1);
      parities |= (base << (2 * j)) | ((base ^ flip) << (2 * j + 1));
      if (j < 3) x.commit_off[j] = cs[j] == RL_CHAIN_NONE ? 0 : (uint16_t)(8 * cs[j] + 1);
    }
    x.parities = (uint16_t)parities;
  }
  RlChainEpiOp* de = reinterpret_cast<RlChainEpiOp*>(blob.data());
  {
    int pos[MAX_WORKERS] = {0, 0, 0, 0};
    for (int k = 1; k < MAX_WORKERS; ++k) pos[k] = pos[k - 1] + n_epi_w[k - 1];
    // stored tensors get a second map with a [32 x 64] box (each epilogue warp stores its own rows of a box);
    // the device op names that map
    int st_of[RL_CHAIN_MAX_TENSORS];
    for (int t = 0; t < RL_CHAIN_MAX_TENSORS; ++t) st_of[t] = -1;
    int n_st = 0;
    for (int i = 0; i < d->n_epis; ++i) {
      RlChainEpiOp o = d->epis_host[i];
      if (o.store_tensor != RL_CHAIN_NONE) {
        if (st_of[o.store_tensor] < 0) {
          if (n_st >= MAX_STORE_MAPS) { delete h; set_error("rl_chain_create: more than %d stored tensors", MAX_STORE_MAPS); return RL_ERR_BAD_ARG; }
          const RlChainTensor& T = d->tensors[o.store_tensor];
          if ((rc = make_tmap_bf16(&h->params.tmaps_st[n_st], T.base, (uint64_t)T.rows, (uint64_t)T.cols, (uint64_t)T.ld, 32)) != RL_OK) {
            delete h;
            return rc;
          }
          st_of[o.store_tensor] = n_st++;
        }
        o.store_tensor = (uint8_t)st_of[o.store_tensor];
      }
      de[pos[o.worker]++] = o;
    }
  }
  cudaError_t err = cudaMalloc(&h->dev_ops, blob.size());
  if (err != cudaSuccess) { delete h; set_error("rl_chain_create: cudaMalloc: %s", cudaGetErrorString(err)); return RL_ERR_CUDA; }
  uint8_t* base = reinterpret_cast<uint8_t*>(h->dev_ops);
  err = cudaMemcpy(base, blob.data(), blob.size(), cudaMemcpyHostToDevice);
  if (err != cudaSuccess) { cudaFree(h->dev_ops); delete h; set_error("rl_chain_create: cudaMemcpy: %s", cudaGetErrorString(err)); return RL_ERR_CUDA; }
  h->params.epis[0] = reinterpret_cast<const RlChainEpiOp*>(base);
  for (int k = 1; k < MAX_WORKERS; ++k) h->params.epis[k] = h->params.epis[k - 1] + n_epi_w[k - 1];
  h->n_workers = n_workers;
  h->params.loss_value_out = -1;
  for (int i = 0; i < RL_CHAIN_MAX_OUTPUTS; ++i) h->out_worker[i] = h->out_pos[i] = -1;
  for (int i = 0; i < d->n_epis; ++i) {
    const RlChainEpiOp& o = d->epis_host[i];
    if (o.mode == RL_CHAIN_EPI_BIAS_F32) { h->out_worker[o.out_id] = o.worker; h->out_pos[o.out_id] = i; h->out_ld[o.out_id] = o.out_ld; }
  }
  h->params.params = d->params;
  for (int i = 0; i < RL_CHAIN_MAX_OUTPUTS; ++i) h->params.outputs[i] = d->outputs[i];
  h->params.n_loads = d->n_loads; h->params.n_mmas = d->n_mmas;
  {
    int nl[N_ISSUERS] = {0, 0}, nm[N_ISSUERS] = {0, 0};
    for (int i = 0; i < d->n_loads; ++i) ++nl[d->loads_host[i].issuer];
    for (int i = 0; i < d->n_mmas; ++i) ++nm[d->mmas_host[i].issuer];
    h->params.load_begin[0] = h->params.mma_begin[0] = 0;
    for (int k = 0; k < N_ISSUERS; ++k) {
      h->params.load_begin[k + 1] = h->params.load_begin[k] + nl[k];
      h->params.mma_begin[k + 1] = h->params.mma_begin[k] + nm[k];
    }
  }
  for (int k = 0; k < MAX_WORKERS; ++k) h->params.n_epis[k] = n_epi_w[k];
  h->params.n_units = d->n_units; h->params.n_barriers = d->n_barriers;
  memcpy(h->params.barrier_count, d->barrier_count, RL_CHAIN_MAX_BARRIERS);
  h->smem_bytes = (size_t)d->n_units * UNIT_BYTES;      // dynamic part; FIXED_SMEM bytes are static
  static size_t configured = 0;
  if (h->smem_bytes > configured) {
    const int bytes = (int)h->smem_bytes;
    const cudaFuncAttribute at = cudaFuncAttributeMaxDynamicSharedMemorySize;
    err = cudaFuncSetAttribute(mlp_chain_kernel<false, 2>, at, bytes);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(mlp_chain_kernel<true, 2>, at, bytes);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(mlp_chain_kernel<false, 3>, at, bytes);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(mlp_chain_kernel<true, 3>, at, bytes);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(mlp_chain_kernel<false, 4>, at, bytes);
    if (err == cudaSuccess) err = cudaFuncSetAttribute(mlp_chain_kernel<true, 4>, at, bytes);
    if (err != cudaSuccess) { cudaFree(h->dev_ops); delete h; set_error("rl_chain_create: smem %zu B: %s", h->smem_bytes, cudaGetErrorString(err)); return RL_ERR_CUDA; }
    configured = h->smem_bytes;
  }
  h->trace = nullptr;
  h->trace_len = (size_t)d->n_loads + TRACE_MMA * (size_t)d->n_mmas + TRACE_EPI * (size_t)d->n_epis;
  *handle = h;
  return RL_OK;
}

// Profiling aid: per-op clock64 stamps of CTA 0 in tile iteration `tile_iteration` of the next runs
// (1 stamp per LOAD op, 8 per MMA op: waits passed, after each K16 step (4), after each commit (3), 5 per EPI op: start, accumulator
// ready, registers loaded, elementwise done, end).  tile_iteration < 0 switches tracing off.
extern "C" int rl_chain_trace(void* handle, int32_t tile_iteration) {
  RL_REQUIRE(handle, RL_ERR_BAD_ARG, "rl_chain_trace: null handle");
  ChainHandle* h = reinterpret_cast<ChainHandle*>(handle);
  if (tile_iteration >= 0 && !h->trace) {
    cudaError_t err = cudaMalloc(&h->trace, h->trace_len * 8);
    RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "rl_chain_trace: cudaMalloc: %s", cudaGetErrorString(err));
    cudaMemset(h->trace, 0, h->trace_len * 8);
  }
  h->params.trace = tile_iteration >= 0 ? h->trace : nullptr;
  h->params.trace_it = tile_iteration;
  return RL_OK;
}
// copies the stamps to `out_host` (capacity in stamps); returns the number of stamps or a negative error
extern "C" int64_t rl_chain_read_trace(void* handle, uint64_t* out_host, int64_t capacity) {
  RL_REQUIRE(handle && out_host, RL_ERR_BAD_ARG, "rl_chain_read_trace: null argument");
  ChainHandle* h = reinterpret_cast<ChainHandle*>(handle);
  RL_REQUIRE(h->trace, RL_ERR_BAD_ARG, "rl_chain_read_trace: tracing was never enabled");
  const size_t n = (size_t)capacity < h->trace_len ? (size_t)capacity : h->trace_len;
  cudaError_t err = cudaMemcpy(out_host, h->trace, n * 8, cudaMemcpyDeviceToHost);
  RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "rl_chain_read_trace: %s", cudaGetErrorString(err));
  return (int64_t)n;
}

static int chain_launch(void* handle, int32_t rows, int32_t tile_begin, int32_t tile_end, void* stream);

extern "C" int rl_chain_run(void* handle, int32_t rows, void* stream) {
  RL_REQUIRE(handle && rows > 0, RL_ERR_BAD_ARG, "rl_chain_run: handle=%p rows=%d", handle, rows);
  return chain_launch(handle, rows, 0, (rows + 127) / 128, stream);
}

extern "C" int rl_chain_run_tiles(void* handle, int32_t rows, int32_t tile_begin, int32_t tile_end, void* stream) {
  RL_REQUIRE(handle && rows > 0 && tile_begin >= 0 && tile_begin < tile_end && tile_end <= (rows + 127) / 128, RL_ERR_BAD_ARG,
             "rl_chain_run_tiles: handle=%p rows=%d tiles [%d, %d)", handle, rows, tile_begin, tile_end);
  return chain_launch(handle, rows, tile_begin, tile_end, stream);
}

static int chain_launch(void* handle, int32_t rows, int32_t tile_begin, int32_t tile_end, void* stream) {
  ChainHandle* h = reinterpret_cast<ChainHandle*>(handle);
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  ChainParams p = h->params;
  {
    const char* e = getenv("RL_CHAIN_DBG_SLEEP");
    p.dbg_sleep[0] = p.dbg_sleep[1] = p.dbg_sleep[2] = 0;
    if (e) sscanf(e, "%d,%d,%d", &p.dbg_sleep[0], &p.dbg_sleep[1], &p.dbg_sleep[2]);
  }
  p.rows = rows;
  p.num_tiles = tile_end;
  p.tile0 = tile_begin;
  const int n_tiles = tile_end - tile_begin;
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  cudaStream_t st = (cudaStream_t)stream;
#define RL_CHAIN_LAUNCH(NWK)                                                                                  \
  do {                                                                                                        \
    if (p.trace) mlp_chain_kernel<true, NWK><<<grid, chain_threads(NWK), h->smem_bytes, st>>>(p);             \
    else mlp_chain_kernel<false, NWK><<<grid, chain_threads(NWK), h->smem_bytes, st>>>(p);                    \
  } while (0)
  if (h->n_workers <= 2) RL_CHAIN_LAUNCH(2);
  else if (h->n_workers == 3) RL_CHAIN_LAUNCH(3);
  else RL_CHAIN_LAUNCH(4);
#undef RL_CHAIN_LAUNCH
  return check_launch("mlp_chain_kernel");
}

extern "C" int rl_chain_set_ppo_loss(void* handle, const RlChainPpoLoss* L) {
  RL_REQUIRE(handle, RL_ERR_BAD_ARG, "rl_chain_set_ppo_loss: null handle");
  ChainHandle* h = reinterpret_cast<ChainHandle*>(handle);
  if (!L) { h->params.loss_value_out = -1; return RL_OK; }
  RL_REQUIRE(L->Lrow && L->std && L->dmean && L->dvalue && L->dstd && L->stats && L->inv_global_B > 0.f, RL_ERR_BAD_ARG,
             "rl_chain_set_ppo_loss: null pointer / inv_global_B");
  RL_REQUIRE(L->mean_out >= 0 && L->mean_out < RL_CHAIN_MAX_OUTPUTS && L->value_out >= 0 && L->value_out < RL_CHAIN_MAX_OUTPUTS &&
             L->mean_out != L->value_out && h->out_pos[L->mean_out] >= 0 && h->out_pos[L->value_out] >= 0, RL_ERR_BAD_ARG,
             "rl_chain_set_ppo_loss: outputs %d / %d are not written by this chain", L->mean_out, L->value_out);
  // the thread that evaluates row r reads mean[r] back: it must be the thread that wrote it, earlier in program order
  RL_REQUIRE(h->n_workers <= 3 && h->out_worker[L->mean_out] == h->out_worker[L->value_out] && h->out_pos[L->mean_out] < h->out_pos[L->value_out],
             RL_ERR_BAD_ARG, "rl_chain_set_ppo_loss: the mean rows must be written by the value op's epilogue worker, before it "
             "(workers %d / %d, ops %d / %d, %d workers)", h->out_worker[L->mean_out], h->out_worker[L->value_out],
             h->out_pos[L->mean_out], h->out_pos[L->value_out], h->n_workers);
  RL_REQUIRE(h->out_ld[L->mean_out] == ACT, RL_ERR_BAD_ARG, "rl_chain_set_ppo_loss: mean rows have pitch %d, the loss reads pitch %d",
             h->out_ld[L->mean_out], ACT);
  LossArgs& a = h->params.loss;
  memset(&a, 0, sizeof(a));
  a.mean = h->params.outputs[L->mean_out]; a.value = h->params.outputs[L->value_out];
  a.Lrow = L->Lrow; a.std = L->std; a.clip = L->clip; a.value_coef = L->value_coef; a.entropy_coef = L->entropy_coef;
  a.use_clipped_value = L->use_clipped_value; a.inv_global_B = L->inv_global_B;
  a.dmean = (__nv_bfloat16*)L->dmean; a.dvalue = (__nv_bfloat16*)L->dvalue; a.dstd = L->dstd; a.stats = L->stats; a.kl_slot = L->kl_slot;
  h->params.loss_value_out = L->value_out;
  return RL_OK;
}

extern "C" int rl_chain_destroy(void* handle) {
  if (!handle) return RL_OK;
  ChainHandle* h = reinterpret_cast<ChainHandle*>(handle);
  cudaFree(h->dev_ops);
  if (h->trace) cudaFree(h->trace);
  delete h;
  return RL_OK;
}
