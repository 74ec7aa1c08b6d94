// Persistent grouped wgrad: the weight / bias gradients of every layer touched by one loss.backward()
// (mini_gym_learn/ppo/ppo.py:147 and :166; what autograd computes for nn.Linear in actor_critic.py:38-100):
//     dW_p[M_p, N_p] += dY_p[K, M_p]^T X_p[K, N_p],   db_p[m] += sum_k dY_p[k, m]
// for all problems p in ONE launch of one persistent CTA per SM.
//
// Why not one split-K GEMM per layer (gemm_tc.cu): a wgrad CTA reduces only 15-64 k-blocks (4-16k tensor
// cycles) and then spends as long again draining its 128 x 128 fp32 tile with atomics, with nothing
// overlapped (one CTA per SM: the ring takes the shared memory) - measured 150-270 TFLOP/s.  Here
//   warp 0    TMA producer: streams the k-blocks of work item after work item through a 4-stage ring
//   warp 1    tcgen05.mma issuer: accumulators DOUBLE BUFFERED in tensor memory (2 x (BN + 32) columns)
//   warps 2-5 epilogue: tcgen05.ld -> fp32 staging tile in its own shared memory -> row-contiguous atomics,
//             overlapped with the MMAs of the next work item
// A work item is (problem, m tile, n tile, k range); items are dealt round robin to the CTAs.  Operands
// stay row-major [K, M] / [K, N] (MN-major UMMA descriptors): no transposed copy of any activation exists.
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace rl {
namespace tc {

constexpr int WG_THREADS = 192;
constexpr int WG_STAGES = 4;
constexpr int WG_MAX_PROBLEMS = 16;
constexpr int WG_A_BYTES = 128 * 64 * 2;                 // [64 k, 128 m] bf16
constexpr int WG_B_BYTES = 128 * 64 * 2;                 // [64 k, 128 n] bf16 (half used when the problem's tile is 64 wide)
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_PITCH = 128 + 4;                        // fp32 staging pitch (floats)
constexpr int WG_STAGING_BYTES = 128 * WG_PITCH * 4;
constexpr int WG_ONES_BYTES = 16 * 128;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + WG_STAGING_BYTES + WG_ONES_BYTES + 1024 + 256;
constexpr int WG_ACC_COLS = 128 + 32;                    // accumulator + the ones-MMA columns of the bias gradient

struct WgProblem {
  CUtensorMap tmA, tmB;        // dY [K, M] and X [K, N], boxes [64 k, 64 cols]
  float* C;
  float* db;
  int M, N, ldc;
  int bn;                      // tile width: 64 or 128
  int gx, gy, splits, kb_per_split, total_kb;
  int first_item;              // prefix sum of items
};
struct WgArgs {
  WgProblem p[WG_MAX_PROBLEMS];
  int n_problems, n_items;
};

struct WgItem { int p, m0, n0, ny, kb_begin, num_kb; };

__device__ __forceinline__ WgItem wg_decode(const WgArgs& A, int item) {
  int p = 0;
  while (p + 1 < A.n_problems && item >= A.p[p + 1].first_item) ++p;
  const WgProblem& P = A.p[p];
  const int local = item - P.first_item;
  const int bx = local % P.gx, by = (local / P.gx) % P.gy, bz = local / (P.gx * P.gy);
  WgItem w;
  w.p = p; w.m0 = bx * 128; w.n0 = by * P.bn; w.ny = by;
  w.kb_begin = bz * P.kb_per_split;
  const int kb_end = min(P.total_kb, w.kb_begin + P.kb_per_split);
  w.num_kb = max(0, kb_end - w.kb_begin);
  return w;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_persistent_kernel(const __grid_constant__ WgArgs A) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  float* staging = reinterpret_cast<float*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint8_t* ones = reinterpret_cast<uint8_t*>(staging) + WG_STAGING_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ones + WG_ONES_BYTES);
  uint64_t* empty_bar = full_bar + WG_STAGES;
  uint64_t* acc_full = empty_bar + WG_STAGES;     // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < WG_ONES_BYTES / 2; i += WG_THREADS) reinterpret_cast<__nv_bfloat16*>(ones)[i] = __float2bfloat16(1.0f);
  fence_async_smem();
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer: one continuous ring over all of this CTA's work items =====
    if (lane == 0) {
      uint32_t kcount = 0;
      for (int item = blockIdx.x; item < A.n_items; item += gridDim.x) {
        const WgItem w = wg_decode(A, item);
        const WgProblem& P = A.p[w.p];
        const int nb = P.bn / 64;
        const uint32_t bytes = WG_A_BYTES + nb * 8192;
        for (int i = 0; i < w.num_kb; ++i, ++kcount) {
          const int s = kcount % WG_STAGES;
          const uint32_t use = kcount / WG_STAGES;
          if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1);
          uint8_t* a_dst = tiles + s * WG_STAGE_BYTES;
          uint8_t* b_dst = a_dst + WG_A_BYTES;
          const int k0 = (w.kb_begin + i) * 64;
          mbar_expect_tx(&full_bar[s], bytes);
          tma_load_2d(a_dst, &P.tmA, w.m0, k0, &full_bar[s]);
          tma_load_2d(a_dst + 8192, &P.tmA, w.m0 + 64, k0, &full_bar[s]);
          tma_load_2d(b_dst, &P.tmB, w.n0, k0, &full_bar[s]);
          if (nb > 1) tma_load_2d(b_dst + 8192, &P.tmB, w.n0 + 64, k0, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t ones_addr = smem_u32(ones);
      constexpr uint32_t idesc_db = instr_desc_bf16(128, 16, true, true);
      uint32_t kcount = 0;
      int j = 0;
      for (int item = blockIdx.x; item < A.n_items; item += gridDim.x, ++j) {
        const WgItem w = wg_decode(A, item);
        const WgProblem& P = A.p[w.p];
        const int buf = j & 1;
        const uint32_t use = (uint32_t)j >> 1;
        if (use > 0) mbar_wait(&acc_empty[buf], (use - 1) & 1);     // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * WG_ACC_COLS;
        const uint32_t idesc = instr_desc_bf16(128, P.bn, true, true);
        const bool want_db = P.db != nullptr && w.ny == 0;
        for (int i = 0; i < w.num_kb; ++i, ++kcount) {
          const int s = kcount % WG_STAGES;
          mbar_wait(&full_bar[s], (kcount / WG_STAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + s * WG_STAGE_BYTES);
          const uint32_t b_addr = a_addr + WG_A_BYTES;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = smem_desc_sw128(a_addr + kk * 2048, 8192, 1024);
            const uint64_t db_ = smem_desc_sw128(b_addr + kk * 2048, 8192, 1024);
            mma_bf16_ss(acc, da, db_, idesc, (i | kk) != 0);
            if (want_db) {
              const uint64_t dones = smem_desc_sw128(ones_addr, 8192, 1024);
              mma_bf16_ss(acc + 128, da, dones, idesc_db, (i | kk) != 0);
            }
          }
          mma_commit(&empty_bar[s]);
        }
        mma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===== epilogue: drain accumulator j & 1 while the MMA warp fills the other one =====
    const int g = warp & 3;
    const int lr = 32 * g + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * g) << 16);
    int j = 0;
    for (int item = blockIdx.x; item < A.n_items; item += gridDim.x, ++j) {
      const WgItem w = wg_decode(A, item);
      const WgProblem& P = A.p[w.p];
      const int buf = j & 1;
      mbar_wait(&acc_full[buf], ((uint32_t)j >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = lane_base + buf * WG_ACC_COLS;
      const int nchunks = P.bn / 32;
      // the previous item's atomics have finished reading the staging tile
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c = 0; c < nchunks; ++c) {
        uint32_t v[32];
        tmem_ld32(acc + c * 32, v);
        tmem_ld_wait();
        float* row = staging + lr * WG_PITCH + c * 32;
#pragma unroll
        for (int q = 0; q < 32; q += 4)
          *reinterpret_cast<float4*>(row + q) = make_float4(__uint_as_float(v[q]), __uint_as_float(v[q + 1]), __uint_as_float(v[q + 2]),
                                                            __uint_as_float(v[q + 3]));
      }
      const bool want_db = P.db != nullptr && w.ny == 0;
      float dbv = 0.f;
      if (want_db) {
        uint32_t v[32];
        tmem_ld32(acc + 128, v);
        tmem_ld_wait();
        dbv = __uint_as_float(v[0]);
      }
      // accumulator is in shared memory / registers: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (want_db && w.m0 + lr < P.M) atomicAdd(P.db + w.m0 + lr, dbv);
      asm volatile("bar.sync 1, 128;" ::: "memory");        // staging tile complete
      // warp g drains rows g, g+4, ...: one atomic instruction covers 32 consecutive columns of a row
#pragma unroll 1
      for (int rr = g; rr < 128; rr += 4) {
        const int row = w.m0 + rr;
        if (row >= P.M) break;
        for (int cc = 0; cc < nchunks; ++cc) {
          const int n = w.n0 + cc * 32 + lane;
          if (n < P.N) atomicAdd(P.C + (size_t)row * P.ldc + n, staging[rr * WG_PITCH + cc * 32 + lane]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace rl

using namespace rl;
using namespace rl::tc;

// host: builds the work list for `n` problems and launches one persistent grid
int wgrad_persistent_launch(const RlWgradProblem* pr, int n, cudaStream_t st) {
  RL_REQUIRE(n <= WG_MAX_PROBLEMS, RL_ERR_BAD_ARG, "rl_wgrad_grouped: at most %d problems per call (got %d)", WG_MAX_PROBLEMS, n);
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
    cudaError_t err = cudaFuncSetAttribute(wgrad_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "cudaFuncSetAttribute(wgrad_persistent): %s", cudaGetErrorString(err));
  }
  WgArgs A;
  memset(&A, 0, sizeof(A));
  // k-blocks per work item: about three items per CTA, 48..128 k-blocks each (first measured at 24000 rows: 16 / 32 /
  // 64 k-blocks -> 93 / 69 / 81 us for the 10 policy layers, 44 / 40 / 48 us for the adaptation module)
  long tile_kb = 0;
  for (int i = 0; i < n; ++i) {
    const RlWgradProblem& q = pr[i];
    RL_REQUIRE(q.dY && q.X && q.dW && q.M > 0 && q.N > 0 && q.K > 0, RL_ERR_BAD_ARG, "rl_wgrad_grouped: problem %d", i);
    const int bn = q.N <= 64 ? 64 : 128;
    tile_kb += (long)((q.M + 127) / 128) * ((q.N + bn - 1) / bn) * ((q.K + 63) / 64);
  }
  // (re-measured on the final update schedule, A/B on two boxes, ms per update at 4000 envs: 24 / 32 / 40 / 48 / 56 / 64 / 75
  // k-blocks per item -> 7.54 / 7.32 / 7.31 / 7.22 / 7.37 / 7.42 / 7.71: the lower bound is 48 now; 32768 envs: 96 / 128 / 192 /
  // 256 -> 37.26 / 37.14 / 37.02 / 37.43, the upper bound stays)
  long kb_item = tile_kb / (3L * sm_count);
  if (kb_item < 48) kb_item = 48;
  if (kb_item > 128) kb_item = 128;
  { const char* e = getenv("RL_WGRAD_KB"); if (e && atoi(e) > 0) kb_item = atoi(e); }      // tuning override
  int items = 0;
  for (int i = 0; i < n; ++i) {
    const RlWgradProblem& q = pr[i];
    WgProblem& P = A.p[i];
    int rc;
    if ((rc = make_tmap_bf16(&P.tmA, q.dY, q.K, q.M, q.ld_dy, 64)) != RL_OK) return rc;
    if ((rc = make_tmap_bf16(&P.tmB, q.X, q.K, q.N, q.ld_x, 64)) != RL_OK) return rc;
    P.C = q.dW; P.db = q.db; P.M = q.M; P.N = q.N; P.ldc = q.ld_dw;
    P.bn = q.N <= 64 ? 64 : 128;
    P.gx = (q.M + 127) / 128; P.gy = (q.N + P.bn - 1) / P.bn;
    P.total_kb = (q.K + 63) / 64;
    const int want = q.split_k > 0 ? (P.total_kb + q.split_k - 1) / q.split_k : (int)kb_item;   // split_k > 0: caller's choice
    P.kb_per_split = want < 1 ? 1 : want;
    P.splits = (P.total_kb + P.kb_per_split - 1) / P.kb_per_split;
    P.first_item = items;
    items += P.gx * P.gy * P.splits;
  }
  A.n_problems = n; A.n_items = items;
  const int grid = items < sm_count ? items : sm_count;
  wgrad_persistent_kernel<<<grid, WG_THREADS, WG_SMEM, st>>>(A);
  return check_launch("wgrad_persistent_kernel");
}
