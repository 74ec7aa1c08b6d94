// tcgen05 GEMM with fused epilogues: the dense contraction of the learner
// (mini_gym_learn/ppo/actor_critic.py:38-100 nn.Linear layers, forward and backward in
// mini_gym_learn/ppo/ppo.py:102-168).  The reference runs these as cuBLAS SGEMMs + separate
// bias / ELU / autograd kernels; here one kernel does C = A x B on the 5th-generation tensor cores
// with the bias, ELU, ELU-derivative masking, bf16 down-conversion, split-K accumulation and the
// bias-gradient column sums fused into its epilogue.
//
//   NT  (forward, dgrad):  C[M,N] = A[M,K] * B[N,K]^T      A, B row-major bf16, K contiguous
//   TN  (wgrad):           C[M,N] = A[K,M]^T * B[K,N]      A, B row-major bf16, M / N contiguous
//
// Mapping to B200: one CTA per 128 x BN output tile (x one K split); 192 threads =
//   warp 0    TMA producer: cp.async.bulk.tensor 2-D loads (SWIZZLE_128B) into a 4-stage smem ring
//   warp 1    TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=BN, K=16, kind::f16,
//             fp32 accumulators in TMEM), tcgen05.commit releases smem stages / signals the epilogue
//   warps 2-5 epilogue: tcgen05.ld 32x32b -> registers -> fused epilogue -> global
// Operands stay in the TMA-written swizzled layout; the UMMA shared-memory descriptors address them
// in place (K-major for NT, MN-major for TN), so no transposed copies of activations are ever made.
#include <stdlib.h>

#include "tc_common.cuh"

namespace rl {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 192;

enum Epilogue : int {
  EPI_F32 = 0,          // C(fp32) = acc
  EPI_F32_ATOMIC = 1,   // C(fp32) += acc                      (split-K wgrad)
  EPI_BIAS_ELU_BF16 = 2,// C(bf16) = elu(acc + bias[n])        (hidden layer forward)
  EPI_BIAS_F32 = 3,     // C(fp32) = acc + bias[n]             (output layer forward)
  EPI_DELU_BF16 = 4,    // C(bf16) = acc * elu'(aux[m,n])      (dgrad through the previous ELU)
  EPI_BF16 = 5,         // C(bf16) = acc
  EPI_BIAS_BF16 = 6,    // C(bf16) = acc + bias[n]             (encoder latent into the actor/critic input)
};

struct GemmArgs {
  CUtensorMap tmA, tmB;
  CUtensorMap tmC;         // bf16 output tile store (valid when tma_store)
  CUtensorMap tmAux;       // bf16 aux tile load (valid when epi == EPI_DELU_BF16 and tma_aux)
  int tma_store, tma_aux;
  void* C;
  const float* bias;
  const __nv_bfloat16* aux;
  float* db;               // TN only: db[m] += sum_k A[k,m]  (bias gradient), or null
  int ldc, ld_aux;
  int M, N, K;
  int kblocks_per_split;
  int epi;
};

template <int BN, bool TN, int STAGES>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 2;          // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ONES_BYTES = TN ? 16 * 128 : 0; // 16 k-rows of 128 B filled with bf16 1.0
  static constexpr int BOXES = (BN + 63) / 64;         // 64-column bf16 boxes per output tile
  static constexpr int AUX_BYTES = TN ? 0 : BOXES * 128 * 128;   // aux tile for the dgrad epilogue (NT only)
  static constexpr int F32_PITCH = BN + 4;             // fp32 staging pitch (floats) for the coalesced atomics
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static_assert(BOXES * 128 * 128 <= RING_BYTES, "bf16 staging tile must fit in the stage ring");
  static_assert(!TN || BM * F32_PITCH * 4 <= RING_BYTES, "fp32 staging tile must fit in the stage ring");
  static constexpr int TOTAL = RING_BYTES + AUX_BYTES + ONES_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float elu_f(float x) { return x > 0.f ? x : (__expf(x) - 1.f); }

// STAGES = 2 for short reductions (<= 2 k-blocks per CTA: 48-64 KB of smem, 3 CTAs per SM so that the
// epilogue of one tile overlaps the loads / MMAs of its neighbours), 4 otherwise.
template <int BN, bool TN, int STAGES>
__device__ __forceinline__ void gemm_tile(const GemmArgs& args, const int bx, const int by, const int bz) {
  using S = Smem<BN, TN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  uint8_t* aux_tile = smem + S::RING_BYTES;
  uint8_t* ones = aux_tile + S::AUX_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ones + S::ONES_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_bar = empty_bar + STAGES;
  uint64_t* aux_bar = acc_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = bx * BM, n0 = by * BN;
  const int total_kblocks = (args.K + BK - 1) / BK;
  const int kb_begin = bz * args.kblocks_per_split;
  const int kb_end = min(total_kblocks, kb_begin + args.kblocks_per_split);
  const int num_kb = max(0, kb_end - kb_begin);
  const bool want_db = TN && args.db != nullptr && by == 0;
  constexpr uint32_t TMEM_COLS = (BN + (TN ? 32 : 0)) <= 32 ? 32 : (BN + (TN ? 32 : 0)) <= 64 ? 64 :
                                 (BN + (TN ? 32 : 0)) <= 128 ? 128 : (BN + (TN ? 32 : 0)) <= 256 ? 256 : 512;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc_bar, 1);
    mbar_init(aux_bar, 1);
    mbar_fence_init();
    tma_prefetch_desc(&args.tmA);
    tma_prefetch_desc(&args.tmB);
  }
  if (TN) {  // constant ones tile for the bias-gradient MMA
    for (int i = threadIdx.x; i < S::ONES_BYTES / 2; i += GEMM_THREADS)
      reinterpret_cast<__nv_bfloat16*>(ones)[i] = __float2bfloat16(1.0f);
    fence_async_smem();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      if (!TN && args.tma_aux) {          // aux tile of the dgrad epilogue: lands while the main loop runs
        mbar_expect_tx(aux_bar, S::BOXES * 128 * 128);
#pragma unroll
        for (int bx = 0; bx < S::BOXES; ++bx) tma_load_2d(aux_tile + bx * 16384, &args.tmAux, n0 + 64 * bx, m0, aux_bar);
      }
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES, it = i / STAGES;
        if (it > 0) mbar_wait(&empty_bar[s], (it - 1) & 1);
        uint8_t* a_dst = tiles + s * S::STAGE_BYTES;
        uint8_t* b_dst = a_dst + S::A_BYTES;
        const int k0 = (kb_begin + i) * BK;
        mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
        if (!TN) {
          tma_load_2d(a_dst, &args.tmA, k0, m0, &full_bar[s]);            // box [128 rows, 64 k]
          tma_load_2d(b_dst, &args.tmB, k0, n0, &full_bar[s]);            // box [BN rows, 64 k]
        } else {
#pragma unroll
          for (int c = 0; c < BM / 64; ++c)                                // boxes [64 k-rows, 64 m]
            tma_load_2d(a_dst + c * 8192, &args.tmA, m0 + 64 * c, k0, &full_bar[s]);
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)                                // boxes [64 k-rows, 64 n]
            tma_load_2d(b_dst + c * 8192, &args.tmB, n0 + 64 * c, k0, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_bf16(BM, BN, TN, TN);
      constexpr uint32_t idesc_db = instr_desc_bf16(BM, 16, TN, TN);
      const uint32_t ones_addr = smem_u32(ones);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % STAGES, it = i / STAGES;
        mbar_wait(&full_bar[s], it & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(tiles + s * S::STAGE_BYTES);
        const uint32_t b_addr = a_addr + S::A_BYTES;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          uint64_t da, db_;
          if (!TN) {
            da = smem_desc_sw128(a_addr + kk * 32, 16, 1024);
            db_ = smem_desc_sw128(b_addr + kk * 32, 16, 1024);
          } else {
            da = smem_desc_sw128(a_addr + kk * 2048, 8192, 1024);
            db_ = smem_desc_sw128(b_addr + kk * 2048, 8192, 1024);
          }
          mma_bf16_ss(tmem_base, da, db_, idesc, (i | kk) != 0);
          if (TN && want_db) {
            const uint64_t dones = smem_desc_sw128(ones_addr, 8192, 1024);
            mma_bf16_ss(tmem_base + BN, da, dones, idesc_db, (i | kk) != 0);
          }
        }
        mma_commit(&empty_bar[s]);      // smem stage reusable once these MMAs have read it
      }
      mma_commit(acc_bar);              // accumulators complete
    }
  } else {
    // ===== epilogue warps (TMEM lane group = warp % 4) =====
    // Thread = one accumulator row.  Global traffic is made coalesced by staging the tile in shared
    // memory (the stage ring is free once the accumulators are complete): bf16 tiles leave through TMA
    // tensor stores from the 128 B-swizzled layout, fp32 split-K tiles through row-contiguous atomics.
    const int g = warp & 3;
    const int lr = 32 * g + lane;               // row within the tile
    const int r = m0 + lr;
    const int et = threadIdx.x - 64;            // 0..127 within the epilogue group
    if (num_kb > 0) {
      mbar_wait(acc_bar, 0);
      tc_fence_after();
    }
    const int epi = args.epi;
    const bool bf16_out = !(epi == EPI_F32 || epi == EPI_BIAS_F32 || epi == EPI_F32_ATOMIC);
    const bool staged_bf16 = bf16_out && args.tma_store;
    const bool staged_f32 = TN && epi == EPI_F32_ATOMIC;
    const bool aux_smem = !TN && epi == EPI_DELU_BF16 && args.tma_aux;
    if (aux_smem) mbar_wait(aux_bar, 0);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * g) << 16);
    float* stage_f32 = reinterpret_cast<float*>(tiles);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      if (num_kb > 0) { tmem_ld32(lane_addr + c * 32, v); tmem_ld_wait(); }
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      const int nb = n0 + c * 32;
      const int nvalid = max(0, min(32, args.N - nb));
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (epi == EPI_BIAS_ELU_BF16 || epi == EPI_BIAS_F32 || epi == EPI_BIAS_BF16) {
        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(args.bias + nb) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(args.bias + nb + j));
            f[j] += bb.x; f[j + 1] += bb.y; f[j + 2] += bb.z; f[j + 3] += bb.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < nvalid) f[j] += __ldg(args.bias + nb + j);
        }
      }
      if (epi == EPI_BIAS_ELU_BF16) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = elu_f(f[j]);
      }
      if (epi == EPI_DELU_BF16) {
        if (aux_smem) {
          // 16 B chunks of this row from the swizzled aux tile: box (c>>1), chunk ((c&1)*4 + t) ^ (row & 7)
          const uint8_t* row = aux_tile + (c >> 1) * 16384 + lr * 128;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint4 pk = *reinterpret_cast<const uint4*>(row + ((((c & 1) * 4 + t) ^ (lr & 7)) << 4));
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 y = __bfloat1622float2(h[q]);
              f[t * 8 + 2 * q] *= (y.x > 0.f) ? 1.f : (y.x + 1.f);
              f[t * 8 + 2 * q + 1] *= (y.y > 0.f) ? 1.f : (y.y + 1.f);
            }
          }
        } else if (r < args.M) {
          const __nv_bfloat16* ax = args.aux + (size_t)r * args.ld_aux + nb;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < nvalid) {
              const float y = __bfloat162float(ax[j]);
              f[j] *= (y > 0.f) ? 1.f : (y + 1.f);
            }
          }
        }
      }
      if (staged_bf16) {
        uint8_t* row = tiles + (c >> 1) * 16384 + lr * 128;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          __nv_bfloat162 p0 = __floats2bfloat162_rn(f[t * 8], f[t * 8 + 1]), p1 = __floats2bfloat162_rn(f[t * 8 + 2], f[t * 8 + 3]);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(f[t * 8 + 4], f[t * 8 + 5]), p3 = __floats2bfloat162_rn(f[t * 8 + 6], f[t * 8 + 7]);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
          pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
          *reinterpret_cast<uint4*>(row + ((((c & 1) * 4 + t) ^ (lr & 7)) << 4)) = pk;
        }
      } else if (staged_f32) {
        float* row = stage_f32 + lr * S::F32_PITCH + c * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(row + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
      } else if (r < args.M && nvalid > 0) {
        // direct path: fp32 outputs of the narrow output layers, unaligned bf16 destinations
        if (epi == EPI_F32 || epi == EPI_BIAS_F32) {
          float* dst = reinterpret_cast<float*>(args.C) + (size_t)r * args.ldc + nb;
          if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nvalid) dst[j] = f[j];
          }
        } else if (epi == EPI_F32_ATOMIC) {
          float* dst = reinterpret_cast<float*>(args.C) + (size_t)r * args.ldc + nb;
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < nvalid) atomicAdd(dst + j, f[j]);
        } else {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(args.C) + (size_t)r * args.ldc + nb;
#pragma unroll
          for (int j = 0; j < 32; ++j) if (j < nvalid) dst[j] = __float2bfloat16(f[j]);
        }
      }
    }
    if (staged_bf16) {
      fence_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
#pragma unroll
        for (int bx = 0; bx < S::BOXES; ++bx)
          if (n0 + 64 * bx < args.N) tma_store_2d(&args.tmC, tiles + bx * 16384, n0 + 64 * bx, m0);
        bulk_commit();
        bulk_wait_read0();
      }
    } else if (staged_f32) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // warp g drains rows g, g+4, ...: each atomic instruction covers 32 consecutive columns of one row
      float* Cf = reinterpret_cast<float*>(args.C);
#pragma unroll 1
      for (int rr = g; rr < BM; rr += 4) {
        const int row = m0 + rr;
        if (row >= args.M) break;
#pragma unroll
        for (int cc = 0; cc < BN / 32; ++cc) {
          const int n = n0 + cc * 32 + lane;
          if (n < args.N) atomicAdd(Cf + (size_t)row * args.ldc + n, stage_f32[rr * S::F32_PITCH + cc * 32 + lane]);
        }
      }
    }
    if (TN && want_db && num_kb > 0) {
      // bias gradient: column 0 of the ones-MMA accumulator (all 16 columns are identical)
      uint32_t v[32];
      tmem_ld32(lane_addr + BN, v);
      tmem_ld_wait();
      if (r < args.M) atomicAdd(args.db + r, __uint_as_float(v[0]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, bool TN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ GemmArgs args) {
  gemm_tile<BN, TN, STAGES>(args, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Grouped wgrad: the weight / bias gradients of SEVERAL layers (dW_l = dY_l^T X_l, split-K over the batch
// rows, fp32 atomics into the flat gradient) in ONE launch.  The 13 wgrad GEMMs of a PPO minibatch are
// independent and individually too small to fill 148 SMs (128 x 16 ... 512 x 256 outputs); one grid whose
// CTAs are dealt over (problem, m tile, n tile, k split) keeps every SM busy and pays one launch.
constexpr int MAX_GROUP = 12;
struct GroupedArgs {
  GemmArgs g[MAX_GROUP];
  int first_cta[MAX_GROUP + 1];   // prefix sums of CTAs per problem
  int grid_x[MAX_GROUP], grid_y[MAX_GROUP];
  int n;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
wgrad_grouped_kernel(const __grid_constant__ GroupedArgs G) {
  int p = 0;
  const int cta = blockIdx.x;
  while (p + 1 < G.n && cta >= G.first_cta[p + 1]) ++p;
  const int local = cta - G.first_cta[p];
  // k split slowest: consecutive CTAs share the same K range of dY / X (L2 reuse across the m / n tiles)
  const int gx = G.grid_x[p], gy = G.grid_y[p];
  const int bx = local % gx, by = (local / gx) % gy, bz = local / (gx * gy);
  gemm_tile<BN, true, 4>(G.g[p], bx, by, bz);
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  RL_REQUIRE(fn != nullptr, RL_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  RL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * 2) % 16 == 0, RL_ERR_BAD_ARG,
             "TMA operand must be 16 B aligned with a row pitch that is a multiple of 8 elements (ld=%llu)",
             (unsigned long long)ld);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RL_REQUIRE(rc == CUDA_SUCCESS, RL_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
  return RL_OK;
}

}  // namespace tc
// host: 2-D fp32 tensor map over an SoA block [rows][n] (n = envs, unit stride), box = [box_rows, 32 envs]
// (128 B inner extent, no swizzle: shared memory receives dense [box_rows][32] rows), zero fill / clipping out of bounds
int make_tmap_f32_rows(CUtensorMap* out, const void* base, uint64_t rows, uint64_t n, uint32_t box_rows) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  RL_REQUIRE(fn != nullptr, RL_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  RL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (n * 4) % 16 == 0 && box_rows >= 1 && box_rows <= 256, RL_ERR_BAD_ARG,
             "row-block TMA operand must be 16 B aligned with num_envs %% 4 == 0 (n=%llu, box_rows=%u)", (unsigned long long)n, box_rows);
  cuuint64_t dims[2] = {n, rows};
  cuuint64_t strides[1] = {n * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RL_REQUIRE(rc == CUDA_SUCCESS, RL_ERR_CUDA, "cuTensorMapEncodeTiled (fp32 rows) failed (%d)", (int)rc);
  return RL_OK;
}
namespace tc {

template <int BN, bool TN, int STAGES>
static int configure_one() {
  cudaError_t err = cudaFuncSetAttribute(gemm_bf16_kernel<BN, TN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Smem<BN, TN, STAGES>::TOTAL);
  RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(err));
  return RL_OK;
}

template <int BN, bool TN, int STAGES>
static int launch_gemm_st(const GemmArgs& a, dim3 grid, cudaStream_t st) {
  using S = Smem<BN, TN, STAGES>;
  static bool configured = false;
  if (!configured) {
    int rc = configure_one<BN, TN, STAGES>();
    if (rc != RL_OK) return rc;
    configured = true;
  }
  gemm_bf16_kernel<BN, TN, STAGES><<<grid, GEMM_THREADS, S::TOTAL, st>>>(a);
  return check_launch("gemm_bf16_kernel");
}

template <int BN, bool TN>
static int launch_gemm(const GemmArgs& a, dim3 grid, cudaStream_t st) {
  if constexpr (TN) {
    return launch_gemm_st<BN, TN, 4>(a, grid, st);    // the fp32 staging tile needs the 4-stage ring
  } else {
    static int max_kb2 = -1;     // k-blocks up to which the 2-stage / multi-CTA-per-SM variant is used
    if (max_kb2 < 0) { const char* e = getenv("RL_GEMM_STAGE2_MAXKB"); max_kb2 = e ? atoi(e) : 16; }
    return a.kblocks_per_split <= max_kb2 ? launch_gemm_st<BN, TN, 2>(a, grid, st) : launch_gemm_st<BN, TN, 4>(a, grid, st);
  }
}

}  // namespace tc
}  // namespace rl

using namespace rl;
using namespace rl::tc;

// Sets the dynamic shared-memory limits of every GEMM instantiation and resolves the tensor-map encoder.
// Call once per device before capturing GEMM launches into a CUDA graph.
extern "C" int rl_gemm_init(void) {
  int rc;
  if ((rc = configure_one<32, false, 2>()) || (rc = configure_one<32, false, 4>()) || (rc = configure_one<64, false, 2>()) ||
      (rc = configure_one<64, false, 4>()) || (rc = configure_one<128, false, 2>()) || (rc = configure_one<128, false, 4>()) ||
      (rc = configure_one<64, true, 4>()) || (rc = configure_one<128, true, 4>()))
    return rc;
  {
    cudaError_t e1 = cudaFuncSetAttribute(wgrad_grouped_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<64, true, 4>::TOTAL);
    cudaError_t e2 = cudaFuncSetAttribute(wgrad_grouped_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<128, true, 4>::TOTAL);
    RL_REQUIRE(e1 == cudaSuccess && e2 == cudaSuccess, RL_ERR_CUDA, "cudaFuncSetAttribute(wgrad_grouped)");
  }
  RL_REQUIRE(encode_fn() != nullptr, RL_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  return RL_OK;
}

// fills GemmArgs for the transposed (wgrad) form; returns the grid of that problem
static int build_wgrad_args(GemmArgs& a, const void* A, const void* B, float* C, float* db, int M, int N, int K, int lda, int ldb,
                            int ldc, int split_k, dim3* grid) {
  a.C = C; a.bias = nullptr; a.aux = nullptr; a.db = db;
  a.ldc = ldc; a.ld_aux = 0; a.M = M; a.N = N; a.K = K; a.epi = EPI_F32_ATOMIC;
  const int total_kb = (K + BK - 1) / BK;
  a.kblocks_per_split = (total_kb + split_k - 1) / split_k;
  const int splits = (total_kb + a.kblocks_per_split - 1) / a.kblocks_per_split;
  const int bn = N <= 64 ? 64 : 128;
  int rc;
  if ((rc = make_tmap_bf16(&a.tmA, A, K, M, lda, 64)) != RL_OK) return rc;
  if ((rc = make_tmap_bf16(&a.tmB, B, K, N, ldb, 64)) != RL_OK) return rc;
  a.tma_store = 0; a.tma_aux = 0;
  *grid = dim3((M + BM - 1) / BM, (N + bn - 1) / bn, splits);
  return RL_OK;
}

template <int BN>
static int launch_grouped(const GroupedArgs& G, int total, cudaStream_t st) {
  using S = Smem<BN, true, 4>;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(wgrad_grouped_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "cudaFuncSetAttribute(wgrad_grouped): %s", cudaGetErrorString(err));
    configured = true;
  }
  wgrad_grouped_kernel<BN><<<total, GEMM_THREADS, S::TOTAL, st>>>(G);
  return check_launch("wgrad_grouped_kernel");
}

// C-ABI: n weight-gradient problems dW_i[M_i, N_i] += dY_i[K, M_i]^T X_i[K, N_i] (+ db_i[m] += sum_k dY_i[k, m])
// in at most two launches (one per tile width).  split_k_i: see rl_gemm_bf16.
int wgrad_persistent_launch(const RlWgradProblem* pr, int n, cudaStream_t st);   // wgrad_persistent.cu

extern "C" int rl_wgrad_grouped(const RlWgradProblem* pr, int32_t n, void* stream) {
  RL_REQUIRE(pr && n > 0, RL_ERR_BAD_ARG, "rl_wgrad_grouped: no problems");
  cudaStream_t st = (cudaStream_t)stream;
  static int persistent = -1;     // RL_WGRAD_PERSISTENT=0: one split-K CTA per (tile, k range) instead of the persistent kernel
  if (persistent < 0) { const char* e = getenv("RL_WGRAD_PERSISTENT"); persistent = (e && atoi(e) == 0) ? 0 : 1; }
  if (persistent) return wgrad_persistent_launch(pr, n, st);
  for (int width = 0; width < 2; ++width) {          // 0: outputs with N <= 64 (BN = 64), 1: the rest (BN = 128)
    GroupedArgs G;
    G.n = 0;
    int total = 0;
    for (int i = 0; i < n; ++i) {
      const RlWgradProblem& q = pr[i];
      RL_REQUIRE(q.dY && q.X && q.dW && q.M > 0 && q.N > 0 && q.K > 0 && q.split_k >= 0, RL_ERR_BAD_ARG, "rl_wgrad_grouped: problem %d", i);
      if ((q.N <= 64) != (width == 0)) continue;
      if (G.n == MAX_GROUP) {                         // flush a full group
        G.first_cta[G.n] = total;
        int rc = width == 0 ? launch_grouped<64>(G, total, st) : launch_grouped<128>(G, total, st);
        if (rc != RL_OK) return rc;
        G.n = 0; total = 0;
      }
      dim3 grid;
      int rc = build_wgrad_args(G.g[G.n], q.dY, q.X, q.dW, q.db, q.M, q.N, q.K, q.ld_dy, q.ld_x, q.ld_dw, q.split_k > 0 ? q.split_k : 16, &grid);
      if (rc != RL_OK) return rc;
      G.first_cta[G.n] = total;
      G.grid_x[G.n] = grid.x; G.grid_y[G.n] = grid.y;
      total += grid.x * grid.y * grid.z;
      ++G.n;
    }
    if (G.n > 0) {
      G.first_cta[G.n] = total;
      int rc = width == 0 ? launch_grouped<64>(G, total, st) : launch_grouped<128>(G, total, st);
      if (rc != RL_OK) return rc;
    }
  }
  return RL_OK;
}

// C-ABI.  transposed = 0: C[M,N] = A[M,K] B[N,K]^T (lda, ldb = K pitches);
//         transposed = 1: C[M,N] = A[K,M]^T B[K,N] (lda = pitch of the [K,M] matrix, ldb of [K,N]).
extern "C" int rl_gemm_bf16(const void* A, const void* B, void* C, const float* bias, const void* aux, float* db,
                            int32_t M, int32_t N, int32_t K, int32_t lda, int32_t ldb, int32_t ldc, int32_t ld_aux,
                            int32_t transposed, int32_t epilogue, int32_t split_k, void* stream) {
  RL_REQUIRE(A && B && C, RL_ERR_BAD_ARG, "rl_gemm_bf16: null operand");
  RL_REQUIRE(M > 0 && N > 0 && K > 0, RL_ERR_BAD_ARG, "rl_gemm_bf16: M=%d N=%d K=%d", M, N, K);
  RL_REQUIRE(epilogue >= EPI_F32 && epilogue <= EPI_BIAS_BF16, RL_ERR_BAD_ARG, "rl_gemm_bf16: epilogue=%d", epilogue);
  RL_REQUIRE(!((epilogue == EPI_BIAS_ELU_BF16 || epilogue == EPI_BIAS_F32 || epilogue == EPI_BIAS_BF16) && !bias), RL_ERR_BAD_ARG, "rl_gemm_bf16: bias missing");
  RL_REQUIRE(!(epilogue == EPI_DELU_BF16 && !aux), RL_ERR_BAD_ARG, "rl_gemm_bf16: aux missing");
  RL_REQUIRE(split_k >= 1, RL_ERR_BAD_ARG, "rl_gemm_bf16: split_k=%d", split_k);
  RL_REQUIRE(split_k == 1 || epilogue == EPI_F32_ATOMIC, RL_ERR_BAD_ARG, "rl_gemm_bf16: split-K needs the atomic epilogue");
  RL_REQUIRE(!db || transposed, RL_ERR_BAD_ARG, "rl_gemm_bf16: db only with the transposed (wgrad) form");
  GemmArgs a;
  a.C = C; a.bias = bias; a.aux = reinterpret_cast<const __nv_bfloat16*>(aux); a.db = db;
  a.ldc = ldc; a.ld_aux = ld_aux; a.M = M; a.N = N; a.K = K; a.epi = epilogue;
  const int total_kb = (K + BK - 1) / BK;
  a.kblocks_per_split = (total_kb + split_k - 1) / split_k;
  const int splits = (total_kb + a.kblocks_per_split - 1) / a.kblocks_per_split;
  // tile width: the narrowest of {32, 64, 128} that covers N for small N (the TN form needs >= 64)
  int bn = (N <= 32 && !transposed) ? 32 : (N <= 64 ? 64 : 128);
  int rc;
  if (!transposed) {
    if ((rc = make_tmap_bf16(&a.tmA, A, M, K, lda, BM)) != RL_OK) return rc;
    if ((rc = make_tmap_bf16(&a.tmB, B, N, K, ldb, bn)) != RL_OK) return rc;
  } else {
    if ((rc = make_tmap_bf16(&a.tmA, A, K, M, lda, 64)) != RL_OK) return rc;
    if ((rc = make_tmap_bf16(&a.tmB, B, K, N, ldb, 64)) != RL_OK) return rc;
  }
  a.tma_store = 0; a.tma_aux = 0;
  const bool bf16_out = !(epilogue == EPI_F32 || epilogue == EPI_BIAS_F32 || epilogue == EPI_F32_ATOMIC);
  if (!transposed && bf16_out && (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (ldc % 8) == 0) {
    if ((rc = make_tmap_bf16(&a.tmC, C, M, N, ldc, BM)) != RL_OK) return rc;
    a.tma_store = 1;
  }
  if (!transposed && epilogue == EPI_DELU_BF16 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0 && (ld_aux % 8) == 0) {
    if ((rc = make_tmap_bf16(&a.tmAux, aux, M, N, ld_aux, BM)) != RL_OK) return rc;
    a.tma_aux = 1;
  }
  dim3 grid((M + BM - 1) / BM, (N + bn - 1) / bn, splits);
  cudaStream_t st = (cudaStream_t)stream;
  if (!transposed) {
    if (bn == 32) return launch_gemm<32, false>(a, grid, st);
    if (bn == 64) return launch_gemm<64, false>(a, grid, st);
    return launch_gemm<128, false>(a, grid, st);
  }
  if (bn == 64) return launch_gemm<64, true>(a, grid, st);
  return launch_gemm<128, true>(a, grid, st);
}
