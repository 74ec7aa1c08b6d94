// PPO update support kernels around the tcgen05 GEMMs: minibatch gather + bf16 staging, the fused
// clipped-surrogate / clipped-value / entropy / KL loss with its analytic gradients, gradient-norm
// clipping with the KL-adaptive learning rate kept on the device, fused Adam, bf16 shadow weights,
// and the Normal sampling of ActorCritic.act.
// Replaces mini_gym_learn/ppo/ppo.py:94-178 (PPO.update), rollout_storage.py:100-139 (the twelve
// advanced-index gathers per minibatch) and actor_critic.py:137-147 (Normal sample / log_prob).
// The reference synchronises the host >= 5 times per minibatch (kl_mean compare :118-121, .item()
// :152-153,170, slot_cache :161-162); here nothing leaves the device until update() returns.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "rl_common.cuh"
#include "ppo_loss.cuh"

namespace rl {


// The 630-float history row is 86 % of the gathered bytes.  8 B loads / 4 B bf16x2 stores when the row is 8 B
// aligned (even hist_dim, even pitch); five loads are issued before the first conversion so that every warp keeps
// 1.25 KB in flight (one load at a time left the kernel latency bound at ~45 % of the HBM rate).
__device__ __forceinline__ void gather_history_row(const float* __restrict__ hs, __nv_bfloat16* __restrict__ hd, int hist_dim, int ldh,
                                                   int lane) {
  const __nv_bfloat16 zero = __float2bfloat16(0.f);
  if (((hist_dim | ldh) & 1) == 0 && ((reinterpret_cast<uintptr_t>(hs) | reinterpret_cast<uintptr_t>(hd)) & 7) == 0) {
    constexpr int U = 5;
    for (int c0 = 2 * lane; c0 < ldh; c0 += 64 * U) {
      float2 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + 64 * u;
        v[u] = c < hist_dim ? __ldg(reinterpret_cast<const float2*>(hs + c)) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + 64 * u;
        if (c < ldh) *reinterpret_cast<__nv_bfloat162*>(hd + c) = __floats2bfloat162_rn(v[u].x, v[u].y);
      }
    }
  } else {
    for (int c = lane; c < ldh; c += 32) hd[c] = c < hist_dim ? __float2bfloat16(hs[c]) : zero;
  }
}

// ---- minibatch gather (rollout_storage.py:121-137) + bf16 staging -----------------------------------
// one warp per minibatch row; every global access is a coalesced run along the row (any dimensions)
__global__ void __launch_bounds__(256)
ppo_gather_kernel(const float* __restrict__ obs, const float* __restrict__ priv, const float* __restrict__ hist,
                  const float* __restrict__ actions, const float* __restrict__ values, const float* __restrict__ returns,
                  const float* __restrict__ logp, const float* __restrict__ adv, const float* __restrict__ mu,
                  const float* __restrict__ sigma, const int64_t* __restrict__ idx, int B, int obs_dim, int priv_dim,
                  int hist_dim, __nv_bfloat16* __restrict__ Xp, int ldp, __nv_bfloat16* __restrict__ Xac, int ldac,
                  __nv_bfloat16* __restrict__ Xh, int ldh, float* __restrict__ Lrow) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const size_t src = (size_t)idx[warp];
  const __nv_bfloat16 zero = __float2bfloat16(0.f);
  for (int c = lane; c < ldp; c += 32) Xp[(size_t)warp * ldp + c] = c < priv_dim ? __float2bfloat16(priv[src * priv_dim + c]) : zero;
  for (int c = lane; c < ldac; c += 32) {
    if (c < obs_dim) Xac[(size_t)warp * ldac + c] = __float2bfloat16(obs[src * obs_dim + c]);
    else if (c >= obs_dim + LAT) Xac[(size_t)warp * ldac + c] = zero;   // [obs_dim, obs_dim+18) is the latent slot
  }
  if (Xh) gather_history_row(hist + src * hist_dim, Xh + (size_t)warp * ldh, hist_dim, ldh, lane);
  float* L = Lrow + (size_t)warp * LROW;
  if (lane < ACT) {
    L[lane] = actions[src * ACT + lane];
    L[ACT + lane] = mu[src * ACT + lane];
    L[2 * ACT + lane] = sigma[src * ACT + lane];
  }
  if (lane == 0) { L[36] = logp[src]; L[37] = adv[src]; L[38] = returns[src]; L[39] = values[src]; }
}

// The same rows for the learner's usual shapes (priv_dim <= ldp <= 32, obs_dim <= ldac <= 64, no history in this launch):
// GR rows per warp.  With one row per warp the kernel is a chain of two dependent latencies (the row index, then ~10
// small loads) and ran at ~1/4 of the HBM rate; here the warp fetches its GR indices with one load, issues the loads of
// all GR rows (7 per lane and row, the four per-row scalars on lanes 12 - 15) before the first store, and only then
// converts and stores - the same values to the same places.
constexpr int GATHER_ROWS = 4;
__global__ void __launch_bounds__(256)
ppo_gather_rows_kernel(const float* __restrict__ obs, const float* __restrict__ priv, const float* __restrict__ actions,
                       const float* __restrict__ values, const float* __restrict__ returns, const float* __restrict__ logp,
                       const float* __restrict__ adv, const float* __restrict__ mu, const float* __restrict__ sigma,
                       const int64_t* __restrict__ idx, int B, int obs_dim, int priv_dim, __nv_bfloat16* __restrict__ Xp, int ldp,
                       __nv_bfloat16* __restrict__ Xac, int ldac, float* __restrict__ Lrow) {
  constexpr int GR = GATHER_ROWS;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int row0 = warp * GR;
  if (row0 >= B) return;
  long long mine = 0;
  if (lane < GR && row0 + lane < B) mine = idx[row0 + lane];
  float vp[GR], vo0[GR], vo1[GR], va[GR], vm[GR], vs[GR];
  const float* const scalar_src = lane == 12 ? logp : (lane == 13 ? adv : (lane == 14 ? returns : values));
#pragma unroll
  for (int r = 0; r < GR; ++r) {
    const size_t src = (size_t)__shfl_sync(0xFFFFFFFFu, mine, r);
    vp[r] = vo0[r] = vo1[r] = va[r] = vm[r] = vs[r] = 0.f;
    if (row0 + r < B) {
      if (lane < priv_dim) vp[r] = priv[src * priv_dim + lane];
      if (lane < obs_dim) vo0[r] = obs[src * obs_dim + lane];
      if (lane + 32 < obs_dim) vo1[r] = obs[src * obs_dim + 32 + lane];
      if (lane < ACT) {
        va[r] = actions[src * ACT + lane];
        vm[r] = mu[src * ACT + lane];
        vs[r] = sigma[src * ACT + lane];
      } else if (lane < 16) {
        va[r] = scalar_src[src];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < GR; ++r) {
    const int row = row0 + r;
    if (row >= B) break;
    if (lane < ldp) Xp[(size_t)row * ldp + lane] = __float2bfloat16(vp[r]);            // (zero beyond priv_dim)
    __nv_bfloat16* xa = Xac + (size_t)row * ldac;
    if (lane < obs_dim) xa[lane] = __float2bfloat16(vo0[r]);
    else if (lane >= obs_dim + LAT && lane < ldac) xa[lane] = __float2bfloat16(0.f);
    const int c1 = lane + 32;
    if (c1 < obs_dim) xa[c1] = __float2bfloat16(vo1[r]);
    else if (c1 >= obs_dim + LAT && c1 < ldac) xa[c1] = __float2bfloat16(0.f);         // [obs_dim, obs_dim+18) is the latent slot
    float* L = Lrow + (size_t)row * LROW;
    if (lane < ACT) { L[lane] = va[r]; L[ACT + lane] = vm[r]; L[2 * ACT + lane] = vs[r]; }
    else if (lane < 16) L[36 + lane - 12] = va[r];                                     // logp, adv, returns, values
  }
}

// obs_history_batch alone (rollout_storage.py:124): the adaptation module's input, gathered on its own stream
__global__ void __launch_bounds__(256)
ppo_gather_history_kernel(const float* __restrict__ hist, const int64_t* __restrict__ idx, int B, int hist_dim,
                          __nv_bfloat16* __restrict__ Xh, int ldh) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  gather_history_row(hist + (size_t)idx[warp] * hist_dim, Xh + (size_t)warp * ldh, hist_dim, ldh, lane);
}

// fp32 [rows, cols] (pitch ld_src) -> bf16 [rows, ld_dst], zero padded; optional column offset into dst
__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, int ld_src, __nv_bfloat16* __restrict__ dst, int ld_dst, int rows,
                 int cols, int dst_col0, int pad_to) {
  const size_t total = (size_t)rows * pad_to;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / pad_to;
    const int c = (int)(i - r * pad_to);
    dst[r * ld_dst + dst_col0 + c] = c < cols ? __float2bfloat16(src[r * ld_src + c]) : __float2bfloat16(0.f);
  }
}

__global__ void __launch_bounds__(128)
ppo_loss_kernel(const __grid_constant__ LossArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double s_surr = 0, s_val = 0, s_kl = 0, s_ad = 0;
  float g_std[ACT];
#pragma unroll
  for (int d = 0; d < ACT; ++d) g_std[d] = 0.f;
  if (i < a.B) {
    ppo_loss_row(a, i, a.value[i], s_surr, s_val, s_kl, g_std);
    // adaptation regression (:157-164): mse over B x 18 elements
    if (a.pred) {
      const __nv_bfloat16* tgt = a.Xac + (size_t)i * a.ldac + a.lat_off;
      __nv_bfloat16* dp = a.dpred + (size_t)i * 24;
      const float sc = 2.f * a.inv_global_B / (float)LAT;
      float se = 0.f;
#pragma unroll
      for (int d = 0; d < LAT; ++d) {
        const float e = a.pred[(size_t)i * LAT + d] - __bfloat162float(tgt[d]);
        se += e * e;
        dp[d] = __float2bfloat16(sc * e);
      }
#pragma unroll
      for (int d = LAT; d < 24; ++d) dp[d] = __float2bfloat16(0.f);
      s_ad = se;
    }
  }
  // block reduction -> one atomic per block
  __shared__ double sh[4][4];
  __shared__ float sh_std[4][ACT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  s_surr = warp_sum(s_surr); s_val = warp_sum(s_val); s_kl = warp_sum(s_kl); s_ad = warp_sum(s_ad);
#pragma unroll
  for (int d = 0; d < ACT; ++d) g_std[d] = warp_sum(g_std[d]);
  if (lane == 0) {
    sh[warp][0] = s_surr; sh[warp][1] = s_val; sh[warp][2] = s_kl; sh[warp][3] = s_ad;
#pragma unroll
    for (int d = 0; d < ACT; ++d) sh_std[warp][d] = g_std[d];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const double v = sh[0][threadIdx.x] + sh[1][threadIdx.x] + sh[2][threadIdx.x] + sh[3][threadIdx.x];
    atomicAdd(a.stats + threadIdx.x, v);
    if (threadIdx.x == 2 && a.kl_slot) atomicAdd(a.kl_slot, (float)v);
  }
  if (threadIdx.x >= 32 && threadIdx.x < 32 + ACT) {
    const int d = threadIdx.x - 32;
    float g = sh_std[0][d] + sh_std[1][d] + sh_std[2][d] + sh_std[3][d];
    // entropy (:144): mean over rows of sum_d (0.5 + 0.5 log 2pi + log std_d)  =>  d/dstd_d = 1/std_d, added once
    if (blockIdx.x == 0) g += -a.entropy_coef * (1.f / a.std[d]) * (a.inv_global_B * (float)a.B);
    atomicAdd(a.dstd + d, g);
  }
}

// ---- adaptation-module regression (ppo.py:157-164), run AFTER the policy optimiser step so that the
// target latent comes from the updated encoder, exactly like the reference ---------------------------------
__global__ void __launch_bounds__(128)
adapt_loss_kernel(const float* __restrict__ pred, const __nv_bfloat16* __restrict__ Xac, int ldac, int lat_off, int B,
                  float inv_global_B, __nv_bfloat16* __restrict__ dpred, double* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double se = 0.0;
  if (i < B) {
    const __nv_bfloat16* tgt = Xac + (size_t)i * ldac + lat_off;
    __nv_bfloat16* dp = dpred + (size_t)i * 24;
    const float sc = 2.f * inv_global_B / (float)LAT;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < LAT; ++d) {
      const float e = pred[(size_t)i * LAT + d] - __bfloat162float(tgt[d]);
      acc += e * e;
      dp[d] = __float2bfloat16(sc * e);
    }
#pragma unroll
    for (int d = LAT; d < 24; ++d) dp[d] = __float2bfloat16(0.f);
    se = acc;
  }
  __shared__ double sh[4];
  se = warp_sum(se);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = se;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(stats + 3, sh[0] + sh[1] + sh[2] + sh[3]);
}

// ---- gradient norm + clip coefficient + KL-adaptive learning rate (ppo.py:116-124, 149) ---------------
// ctrl (float): [0] learning rate (persistent), [1] clip coefficient, [2] kl mean of this minibatch
struct FinalizeWs { double sumsq; unsigned int ticket; unsigned int pad; };

// the scalar end of both finalize kernels: clip coefficient (clip_grad_norm_), KL mean -> adaptive learning rate (ppo.py:116-124),
// and - when loss_acc is given - the bookkeeping PPO.update used to do with separate launches: the minibatch's loss
// sums are added to the per-update accumulators and the per-minibatch statistics (and the fp32 KL slot) are zeroed.
__device__ __forceinline__ void finalize_tail(double norm, double* stats, float* ctrl, double global_B, float desired_kl,
                                              float max_norm, int adaptive, double* loss_acc, float* kl_slot) {
  const float coef = (float)((double)max_norm / (norm + 1e-6));       // clip_grad_norm_
  ctrl[1] = coef < 1.f ? coef : 1.f;
  const float kl = kl_slot ? (float)((double)*kl_slot / global_B) : (float)(stats[2] / global_B);
  ctrl[2] = kl;
  if (adaptive) {
    float lr = ctrl[0];
    if (kl > desired_kl * 2.0f) lr = fmaxf(1e-5f, lr / 1.5f);
    else if (kl < desired_kl / 2.0f && kl > 0.0f) lr = fminf(1e-2f, lr * 1.5f);
    ctrl[0] = lr;
  }
  if (loss_acc) {
    loss_acc[0] += stats[0]; loss_acc[1] += stats[1]; loss_acc[2] += stats[2];
    stats[0] = 0.0; stats[1] = 0.0; stats[2] = 0.0;
    if (kl_slot) *kl_slot = 0.f;
  }
}

__global__ void __launch_bounds__(256)
grad_finalize_kernel(const float* __restrict__ grad, size_t n, double* stats, float* ctrl,
                     FinalizeWs* ws, double global_B, float desired_kl, float max_norm, int adaptive, double* loss_acc, float* kl_slot) {
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double g = grad[i];
    s += g * g;
  }
  __shared__ double sh[8];
  __shared__ bool last;
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0;
    for (int w = 0; w < 8; ++w) b += sh[w];
    atomicAdd(&ws->sumsq, b);
    __threadfence();
    last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    const double norm = sqrt(*((volatile double*)&ws->sumsq));
    finalize_tail(norm, stats, ctrl, global_B, desired_kl, max_norm, adaptive, loss_acc, kl_slot);
    ws->sumsq = 0.0;
    ws->ticket = 0;
  }
}

// clip coefficient / KL-adaptive learning rate from an already reduced squared gradient norm (multi-GPU: the
// peer all-reduce kernel computes it while it sums the gradients)
__global__ void finalize_from_norm_kernel(const double* __restrict__ norm2, double* stats, float* ctrl,
                                          double global_B, float desired_kl, float max_norm, int adaptive, double* loss_acc,
                                          float* kl_slot) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  finalize_tail(sqrt(*norm2), stats, ctrl, global_B, desired_kl, max_norm, adaptive, loss_acc, kl_slot);
}

// ---- fused Adam (torch.optim.Adam, eps 1e-8, no weight decay), gradient scaled by the clip coefficient
// step_dev (optional): device-resident optimiser step counter {count, ticket}; the bias corrections are then
// derived on the device and the last CTA advances the counter, so the launch can live in a CUDA graph.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
            const float* __restrict__ ctrl, float lr_fixed, int use_ctrl, float beta1, float beta2, float eps,
            float bc1, float bc2_sqrt, float grad_scale, int* step_dev) {
  if (step_dev) {
    const float t = (float)(step_dev[0] + 1);
    bc1 = 1.f - powf(beta1, t);
    bc2_sqrt = sqrtf(1.f - powf(beta2, t));
  }
  const float lr = use_ctrl ? ctrl[0] : lr_fixed;
  const float coef = (use_ctrl ? ctrl[1] : 1.f) * grad_scale;
  const float step_size = lr / bc1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
    g[i] = 0.f;      // zero_grad for the next minibatch
  }
  if (step_dev) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const int done = atomicAdd(step_dev + 1, 1);
      if (done == (int)gridDim.x - 1) { step_dev[1] = 0; atomicAdd(step_dev, 1); }
    }
  }
}

// ---- bf16 shadow weights for the tensor-core operands ------------------------------------------------
struct ShadowLayer {
  const float* w;          // fp32 master [out, in]
  __nv_bfloat16* wb;       // [out, ld_wb]   forward B operand
  __nv_bfloat16* wbt;      // [in, ld_wbt]   dgrad B operand (transpose)
  int out, in, ld_wb, ld_wbt;
};
constexpr int MAX_SHADOW_LAYERS = 16;
struct ShadowArgs { ShadowLayer L[MAX_SHADOW_LAYERS]; int n; };

__global__ void __launch_bounds__(256)
refresh_shadows_kernel(const __grid_constant__ ShadowArgs a) {
  const ShadowLayer& L = a.L[blockIdx.y];
  __shared__ float tile[32][33];
  const int tiles_x = (L.in + 31) / 32, tiles_y = (L.out + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int t = blockIdx.x; t < tiles_x * tiles_y; t += gridDim.x) {
    const int r0 = (t / tiles_x) * 32, c0 = (t % tiles_x) * 32;
    for (int k = ty; k < 32; k += 8) {
      const int r = r0 + k, c = c0 + tx;
      const float val = (r < L.out && c < L.in) ? L.w[(size_t)r * L.in + c] : 0.f;
      tile[k][tx] = val;
      if (r < L.out && c < L.ld_wb) L.wb[(size_t)r * L.ld_wb + c] = __float2bfloat16(val);
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
      const int c = c0 + k, r = r0 + tx;     // transposed write: row = input column
      if (c < L.in && r < L.ld_wbt) L.wbt[(size_t)c * L.ld_wbt + r] = __float2bfloat16(r < L.out ? tile[tx][k] : 0.f);
    }
    __syncthreads();
  }
}


// ---- Adam + the bf16 shadow copies in ONE launch ------------------------------------------------------
// The optimiser step of a minibatch used to be adam_kernel followed by refresh_shadows_kernel: two latency-bound
// launches (~7 + ~9 us) on the critical path of every minibatch, twice (policy, adaptation module).  The weights are
// views of the flat parameter buffer, so the thread that updates flat element i also knows which weight matrix (if any)
// it belongs to and writes both bf16 operands: wb[r, c] (forward) and wbt[c, r] (dgrad).  The transposed 2 B stores are
// scattered (one 32 B sector per thread): 0.6 M parameters = 19 MB of sector writes absorbed by L2.
struct AdamShadowLayer {
  long long start, end;    // flat element range of the [out, in] master weight
  __nv_bfloat16* wb;       // [out, ld_wb]
  __nv_bfloat16* wbt;      // [in, ld_wbt]
  int in, ld_wb, ld_wbt, pad;
};
struct AdamShadowArgs { AdamShadowLayer L[MAX_SHADOW_LAYERS]; int n; };

__global__ void __launch_bounds__(256)
adam_shadows_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                    const float* __restrict__ ctrl, float lr_fixed, int use_ctrl, float beta1, float beta2, float eps,
                    float bc1, float bc2_sqrt, float grad_scale, int* step_dev, const __grid_constant__ AdamShadowArgs S) {
  if (step_dev) {
    const float t = (float)(step_dev[0] + 1);
    bc1 = 1.f - powf(beta1, t);
    bc2_sqrt = sqrtf(1.f - powf(beta2, t));
  }
  const float lr = use_ctrl ? ctrl[0] : lr_fixed;
  const float coef = (use_ctrl ? ctrl[1] : 1.f) * grad_scale;
  const float step_size = lr / bc1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    const float pi = p[i] - step_size * (mi / denom);
    p[i] = pi;
    g[i] = 0.f;      // zero_grad for the next minibatch
    const long long ii = (long long)i;
#pragma unroll 1
    for (int l = 0; l < S.n; ++l) {
      if (ii >= S.L[l].start && ii < S.L[l].end) {
        const int k = (int)(ii - S.L[l].start), in = S.L[l].in;
        const int r = k / in, c = k - r * in;
        const __nv_bfloat16 b16 = __float2bfloat16(pi);
        S.L[l].wb[(size_t)r * S.L[l].ld_wb + c] = b16;
        S.L[l].wbt[(size_t)c * S.L[l].ld_wbt + r] = b16;
        break;
      }
    }
  }
  if (step_dev) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const int done = atomicAdd(step_dev + 1, 1);
      if (done == (int)gridDim.x - 1) { step_dev[1] = 0; atomicAdd(step_dev, 1); }
    }
  }
}

// ---- ActorCritic.act sampling (actor_critic.py:137-147): a = mu + std * N(0,1), log_prob --------------
__global__ void __launch_bounds__(128)
policy_sample_kernel(const float* __restrict__ mean, const float* __restrict__ std, int N, uint64_t seed, uint64_t step,
                     const float* __restrict__ inj_normal, float* __restrict__ actions, float* __restrict__ logp,
                     float* __restrict__ mu_out, float* __restrict__ sigma_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float lp = 0.f;
#pragma unroll
  for (int b4 = 0; b4 < ACT / 4; ++b4) {
    float z[4];
    if (inj_normal) {
#pragma unroll
      for (int k = 0; k < 4; ++k) z[k] = inj_normal[(size_t)i * ACT + b4 * 4 + k];
    } else {
      float u[4];
      rng4(seed, (uint32_t)i, step, RNG_POLICY, (uint32_t)b4, u);
      // Box-Muller on two pairs
      const float r0 = sqrtf(-2.f * __logf(fmaxf(u[0], 5.96e-8f))), r1 = sqrtf(-2.f * __logf(fmaxf(u[2], 5.96e-8f)));
      float s0, c0, s1, c1;
      __sincosf(6.283185307179586f * u[1], &s0, &c0);
      __sincosf(6.283185307179586f * u[3], &s1, &c1);
      z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int d = b4 * 4 + k;
      const float mu = mean[(size_t)i * ACT + d], sg = std[d];
      const float a = mu + sg * z[k];
      actions[(size_t)i * ACT + d] = a;
      mu_out[(size_t)i * ACT + d] = mu;
      sigma_out[(size_t)i * ACT + d] = sg;
      const float diff = a - mu;
      lp += -(diff * diff) / (2.f * sg * sg) - __logf(sg) - 0.9189385332046727f;
    }
  }
  logp[i] = lp;
}

}  // namespace rl

using namespace rl;

extern "C" int rl_ppo_gather(const float* obs, const float* priv, const float* hist, const float* actions,
                             const float* values, const float* returns, const float* logp, const float* adv,
                             const float* mu, const float* sigma, const int64_t* idx, int32_t B, int32_t obs_dim,
                             int32_t priv_dim, int32_t hist_dim, void* Xp, int32_t ldp, void* Xac, int32_t ldac, void* Xh,
                             int32_t ldh, float* Lrow, void* stream) {
  RL_REQUIRE(obs && priv && actions && values && returns && logp && adv && mu && sigma && idx && Xp && Xac && Lrow,
             RL_ERR_BAD_ARG, "rl_ppo_gather: null pointer");
  RL_REQUIRE(B > 0 && priv_dim <= ldp && obs_dim + LAT <= ldac && (!Xh || (hist && hist_dim <= ldh)), RL_ERR_BAD_ARG,
             "rl_ppo_gather: bad dimensions");
  static const bool rows_off = [] { const char* e = getenv("RL_PPO_GATHER_ROWS"); return e && e[0] == '0'; }();
  if (!Xh && !rows_off && priv_dim <= 32 && ldp <= 32 && obs_dim <= 64 && ldac <= 64) {
    const int warps = (B + GATHER_ROWS - 1) / GATHER_ROWS;
    ppo_gather_rows_kernel<<<(warps * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        obs, priv, actions, values, returns, logp, adv, mu, sigma, idx, B, obs_dim, priv_dim, (__nv_bfloat16*)Xp, ldp,
        (__nv_bfloat16*)Xac, ldac, Lrow);
    return check_launch("ppo_gather_rows_kernel");
  }
  const int blocks = (B * 32 + 255) / 256;
  ppo_gather_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      obs, priv, hist, actions, values, returns, logp, adv, mu, sigma, idx, B, obs_dim, priv_dim, hist_dim,
      (__nv_bfloat16*)Xp, ldp, (__nv_bfloat16*)Xac, ldac, (__nv_bfloat16*)Xh, ldh, Lrow);
  return check_launch("ppo_gather_kernel");
}

extern "C" int rl_ppo_gather_history(const float* hist, const int64_t* idx, int32_t B, int32_t hist_dim, void* Xh, int32_t ldh,
                                     void* stream) {
  RL_REQUIRE(hist && idx && Xh, RL_ERR_BAD_ARG, "rl_ppo_gather_history: null pointer");
  RL_REQUIRE(B > 0 && hist_dim > 0 && hist_dim <= ldh, RL_ERR_BAD_ARG, "rl_ppo_gather_history: bad dimensions");
  ppo_gather_history_kernel<<<(B * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(hist, idx, B, hist_dim, (__nv_bfloat16*)Xh, ldh);
  return check_launch("ppo_gather_history_kernel");
}

extern "C" int rl_cast_bf16(const float* src, int32_t ld_src, void* dst, int32_t ld_dst, int32_t rows, int32_t cols,
                            int32_t dst_col0, int32_t pad_to, void* stream) {
  RL_REQUIRE(src && dst && rows > 0 && cols > 0 && pad_to >= cols && dst_col0 + pad_to <= ld_dst, RL_ERR_BAD_ARG,
             "rl_cast_bf16: bad arguments");
  size_t total = (size_t)rows * pad_to;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cast_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols, dst_col0, pad_to);
  return check_launch("cast_bf16_kernel");
}

extern "C" int rl_ppo_loss(const float* mean, const float* value, const float* pred, const void* Xac, int32_t ldac,
                           int32_t lat_off, const float* Lrow, const float* std, int32_t B, float clip, float value_coef,
                           float entropy_coef, int32_t use_clipped_value, float inv_global_B, void* dmean, void* dvalue,
                           void* dpred, float* dstd, double* stats, float* kl_slot, void* stream) {
  RL_REQUIRE(mean && value && Lrow && std && dmean && dvalue && dstd && stats && Xac, RL_ERR_BAD_ARG, "rl_ppo_loss: null pointer");
  RL_REQUIRE(B > 0 && (!pred || dpred), RL_ERR_BAD_ARG, "rl_ppo_loss: bad arguments");
  LossArgs a;
  a.mean = mean; a.value = value; a.pred = pred; a.Xac = (const __nv_bfloat16*)Xac; a.ldac = ldac; a.lat_off = lat_off;
  a.Lrow = Lrow; a.std = std; a.B = B; a.clip = clip; a.value_coef = value_coef; a.entropy_coef = entropy_coef;
  a.use_clipped_value = use_clipped_value; a.inv_global_B = inv_global_B;
  a.dmean = (__nv_bfloat16*)dmean; a.dvalue = (__nv_bfloat16*)dvalue; a.dpred = (__nv_bfloat16*)dpred;
  a.dstd = dstd; a.stats = stats; a.kl_slot = kl_slot;
  ppo_loss_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
  return check_launch("ppo_loss_kernel");
}

extern "C" int rl_adapt_loss(const float* pred, const void* Xac, int32_t ldac, int32_t lat_off, int32_t B,
                             float inv_global_B, void* dpred, double* stats, void* stream) {
  RL_REQUIRE(pred && Xac && dpred && stats && B > 0, RL_ERR_BAD_ARG, "rl_adapt_loss: bad arguments");
  adapt_loss_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(pred, (const __nv_bfloat16*)Xac, ldac, lat_off, B,
                                                                     inv_global_B, (__nv_bfloat16*)dpred, stats);
  return check_launch("adapt_loss_kernel");
}

extern "C" int rl_grad_finalize(const float* grad, int64_t n, double* stats, float* ctrl, void* workspace,
                                double global_B, float desired_kl, float max_grad_norm, int32_t adaptive, double* loss_acc,
                                float* kl_slot, void* stream) {
  RL_REQUIRE(grad && stats && ctrl && workspace && n > 0, RL_ERR_BAD_ARG, "rl_grad_finalize: bad arguments");
  int blocks = (int)((n + 256 * 8 - 1) / (256 * 8));
  if (blocks > 148 * 2) blocks = 148 * 2;
  grad_finalize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(grad, (size_t)n, stats, ctrl, (FinalizeWs*)workspace,
                                                               global_B, desired_kl, max_grad_norm, adaptive, loss_acc, kl_slot);
  return check_launch("grad_finalize_kernel");
}

extern "C" int rl_grad_finalize_from_norm(const double* norm2, double* stats, float* ctrl, double global_B, float desired_kl,
                                          float max_grad_norm, int32_t adaptive, double* loss_acc, float* kl_slot, void* stream) {
  RL_REQUIRE(norm2 && stats && ctrl && global_B > 0, RL_ERR_BAD_ARG, "rl_grad_finalize_from_norm: bad arguments");
  finalize_from_norm_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(norm2, stats, ctrl, global_B, desired_kl, max_grad_norm, adaptive, loss_acc, kl_slot);
  return check_launch("finalize_from_norm_kernel");
}

extern "C" int rl_adam(float* p, float* g, float* m, float* v, int64_t n, const float* ctrl, float lr_fixed,
                       int32_t use_ctrl, float beta1, float beta2, float eps, int32_t step, float grad_scale,
                       int32_t* step_dev, void* stream) {
  RL_REQUIRE(p && g && m && v && n > 0 && (step >= 1 || step_dev) && (!use_ctrl || ctrl), RL_ERR_BAD_ARG, "rl_adam: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)(step > 0 ? step : 1));
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)(step > 0 ? step : 1)));
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (size_t)n, ctrl, lr_fixed, use_ctrl, beta1, beta2, eps, bc1,
                                                      bc2_sqrt, grad_scale, step_dev);
  return check_launch("adam_kernel");
}

extern "C" int rl_adam_shadows(float* p, float* g, float* m, float* v, int64_t n, const float* ctrl, float lr_fixed,
                               int32_t use_ctrl, float beta1, float beta2, float eps, int32_t step, float grad_scale,
                               int32_t* step_dev, const int64_t* w_start, void* const* wb, void* const* wbt,
                               const int32_t* out_dim, const int32_t* in_dim, const int32_t* ld_wb, const int32_t* ld_wbt,
                               int32_t n_layers, void* stream) {
  RL_REQUIRE(p && g && m && v && n > 0 && (step >= 1 || step_dev) && (!use_ctrl || ctrl), RL_ERR_BAD_ARG, "rl_adam_shadows: bad arguments");
  RL_REQUIRE(w_start && wb && wbt && out_dim && in_dim && ld_wb && ld_wbt && n_layers > 0 && n_layers <= MAX_SHADOW_LAYERS,
             RL_ERR_BAD_ARG, "rl_adam_shadows: bad layer table");
  AdamShadowArgs S;
  S.n = n_layers;
  for (int i = 0; i < n_layers; ++i) {
    RL_REQUIRE(w_start[i] >= 0 && w_start[i] + (int64_t)out_dim[i] * in_dim[i] <= n && ld_wb[i] >= in_dim[i] && ld_wbt[i] >= out_dim[i],
               RL_ERR_BAD_ARG, "rl_adam_shadows: layer %d outside the parameter range / pitch too small", i);
    S.L[i].start = w_start[i]; S.L[i].end = w_start[i] + (long long)out_dim[i] * in_dim[i];
    S.L[i].wb = (__nv_bfloat16*)wb[i]; S.L[i].wbt = (__nv_bfloat16*)wbt[i];
    S.L[i].in = in_dim[i]; S.L[i].ld_wb = ld_wb[i]; S.L[i].ld_wbt = ld_wbt[i]; S.L[i].pad = 0;
  }
  const float bc1 = 1.f - powf(beta1, (float)(step > 0 ? step : 1));
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)(step > 0 ? step : 1)));
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_shadows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (size_t)n, ctrl, lr_fixed, use_ctrl, beta1, beta2, eps,
                                                              bc1, bc2_sqrt, grad_scale, step_dev, S);
  return check_launch("adam_shadows_kernel");
}

extern "C" int rl_refresh_shadows(const void* const* w, void* const* wb, void* const* wbt, const int32_t* out_dim,
                                  const int32_t* in_dim, const int32_t* ld_wb, const int32_t* ld_wbt, int32_t n_layers,
                                  void* stream) {
  RL_REQUIRE(w && wb && wbt && out_dim && in_dim && ld_wb && ld_wbt && n_layers > 0 && n_layers <= MAX_SHADOW_LAYERS,
             RL_ERR_BAD_ARG, "rl_refresh_shadows: bad arguments");
  ShadowArgs a;
  a.n = n_layers;
  for (int i = 0; i < n_layers; ++i) {
    a.L[i].w = (const float*)w[i]; a.L[i].wb = (__nv_bfloat16*)wb[i]; a.L[i].wbt = (__nv_bfloat16*)wbt[i];
    a.L[i].out = out_dim[i]; a.L[i].in = in_dim[i]; a.L[i].ld_wb = ld_wb[i]; a.L[i].ld_wbt = ld_wbt[i];
    RL_REQUIRE(ld_wb[i] >= in_dim[i] && ld_wbt[i] >= out_dim[i], RL_ERR_BAD_ARG, "rl_refresh_shadows: pitch too small");
  }
  refresh_shadows_kernel<<<dim3(64, n_layers), 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("refresh_shadows_kernel");
}

extern "C" int rl_policy_sample(const float* mean, const float* std, int32_t N, uint64_t seed, uint64_t step,
                                const float* inj_normal, float* actions, float* logp, float* mu_out, float* sigma_out,
                                void* stream) {
  RL_REQUIRE(mean && std && actions && logp && mu_out && sigma_out && N > 0, RL_ERR_BAD_ARG, "rl_policy_sample: bad arguments");
  policy_sample_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mean, std, N, seed, step, inj_normal, actions, logp,
                                                                        mu_out, sigma_out);
  return check_launch("policy_sample_kernel");
}
