// Grid Adaptive Curriculum on the device.
// Replaces LeggedRobot._resample_commands (legged_robot.py:595-626) and
// RewardThresholdCurriculum.update / Curriculum.sample (curriculum.py:110-119, 55-68), which the
// reference runs in numpy on the host with 6 device->host copies, a [3,k,5202] boolean
// neighbourhood tensor and a Python loop per env.
//
//  phase 1 (scatter)  one thread per resampled env: success test on the per-command reward
//                     sums, then +1 on the int32 incidence counter of every bin in the
//                     axis-aligned neighbourhood and a flag on the env's own bin.  Integer
//                     atomics: order independent, bit exact, all-reducible across GPUs.
//  phase 2 (update)   one thread per bin applies w <- min(1, w + 0.2) exactly k times,
//                     k = flag + count (the reference applies the same clip once per unique
//                     own bin and once per successful env per neighbour; the result depends
//                     only on k).  Then a block-wide double prefix sum builds the normalised cdf.
//  phase 3 (sample)   one thread per env: inverse-cdf draw (== numpy choice(p): searchsorted
//                     right), uniform in the cell, small-command zeroing, command_sums reset.
#include "rl_common.cuh"

namespace rl {

struct GacArgs {
  RlGacCfg cfg;
  RlGacBuffers b;
  uint64_t seed;
  uint64_t step;
};

__device__ inline int gac_env(const RlGacBuffers& b, int N, int i) {
  if (b.ids) return (i < b.n_ids) ? (int)b.ids[i] : -1;
  return (i < N && b.mask[i]) ? i : -1;
}

__global__ void __launch_bounds__(128)
gac_scatter_kernel(const __grid_constant__ GacArgs args) {
  const RlGacCfg& cfg = args.cfg;
  const RlGacBuffers& b = args.b;
  const size_t Ns = (size_t)cfg.num_envs;
  const int e = gac_env(b, cfg.num_envs, blockIdx.x * blockDim.x + threadIdx.x);
  if (e < 0 || e >= cfg.num_train_envs) return;  // only train envs feed the update (:612)
  const float r_lin = b.command_sums[cfg.lin_slot * Ns + e] / cfg.ep_len;  // :604
  const float r_ang = b.command_sums[cfg.ang_slot * Ns + e] / cfg.ep_len;  // :605
  if (!((r_lin > cfg.lin_threshold) && (r_ang > cfg.ang_threshold))) return;  // curriculum.py:114
  const int bin = (int)b.env_command_bins[e];
  const int ny = cfg.dims[1], nz = cfg.dims[2];
  const int ix = bin / (ny * nz), iy = (bin / nz) % ny, iz = bin % nz;
  b.own_flag[bin] = 1;  // curriculum.py:115 (once per unique bin)
  const int ox = 0, oy = cfg.dims[0], oz = cfg.dims[0] + cfg.dims[1];
  for (int x = b.nbr_lo[ox + ix]; x <= b.nbr_hi[ox + ix]; ++x)
    for (int y = b.nbr_lo[oy + iy]; y <= b.nbr_hi[oy + iy]; ++y)
      for (int z = b.nbr_lo[oz + iz]; z <= b.nbr_hi[oz + iz]; ++z)
        atomicAdd(b.hit_count + (x * ny + y) * nz + z, 1);  // curriculum.py:116-119
}

constexpr int GAC_BLOCK = 1024;

__global__ void __launch_bounds__(GAC_BLOCK)
gac_update_cdf_kernel(const __grid_constant__ GacArgs args) {
  const RlGacCfg& cfg = args.cfg;
  const RlGacBuffers& b = args.b;
  const int nb = cfg.n_bins;
  __shared__ double warp_tot[GAC_BLOCK / 32];
  __shared__ double carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0.0;
  __syncthreads();
  // weights update, then an inclusive scan chunk by chunk (n_bins = 5202 -> 6 chunks)
  for (int base = 0; base < nb; base += GAC_BLOCK) {
    const int i = base + tid;
    double w = 0.0;
    if (i < nb) {
      w = b.weights[i];
      int k = b.hit_count[i] + (b.own_flag[i] ? 1 : 0);
      b.hit_count[i] = 0;
      b.own_flag[i] = 0;
      k = k > 8 ? 8 : k;  // saturates at 1.0 after at most 5 applications from 0
      for (int r = 0; r < k; ++r) w = fmin(fmax(w + 0.2, 0.0), 1.0);  // np.clip(w + 0.2, 0, 1)
      b.weights[i] = w;
    }
    double s = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) warp_tot[warp] = s;
    __syncthreads();
    if (warp == 0) {
      double t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u;
      }
      warp_tot[lane] = t;
    }
    __syncthreads();
    const double prefix = carry + (warp > 0 ? warp_tot[warp - 1] : 0.0) + s;
    if (i < nb) b.cdf[i] = prefix;
    __syncthreads();
    if (tid == GAC_BLOCK - 1) carry = prefix;
    __syncthreads();
  }
  const double total = carry;
  for (int i = tid; i < nb; i += GAC_BLOCK) b.cdf[i] = b.cdf[i] / total;
}

__global__ void __launch_bounds__(128)
gac_sample_kernel(const __grid_constant__ GacArgs args) {
  const RlGacCfg& cfg = args.cfg;
  const RlGacBuffers& b = args.b;
  const size_t Ns = (size_t)cfg.num_envs;
  const int e = gac_env(b, cfg.num_envs, blockIdx.x * blockDim.x + threadIdx.x);
  if (e < 0) return;
  double ub, uc[3];
  if (b.u_bin) {
    ub = b.u_bin[e];
    uc[0] = b.u_cell[e * 3 + 0]; uc[1] = b.u_cell[e * 3 + 1]; uc[2] = b.u_cell[e * 3 + 2];
  } else {
    uint32_t r0[4], r1[4];
    Philox::gen(args.seed, (uint32_t)e, (uint32_t)args.step, (uint32_t)(args.step >> 32), (RNG_GAC << 16) | 0, r0);
    Philox::gen(args.seed, (uint32_t)e, (uint32_t)args.step, (uint32_t)(args.step >> 32), (RNG_GAC << 16) | 1, r1);
    ub = u01d(r0[0], r0[1]); uc[0] = u01d(r0[2], r0[3]); uc[1] = u01d(r1[0], r1[1]); uc[2] = u01d(r1[2], r1[3]);
  }
  // numpy Generator.choice(p=...): cdf.searchsorted(u, side='right')
  int lo = 0, hi = cfg.n_bins;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (b.cdf[mid] <= ub) lo = mid + 1; else hi = mid;
  }
  const int bin = lo < cfg.n_bins ? lo : cfg.n_bins - 1;
  const int ny = cfg.dims[1], nz = cfg.dims[2];
  const int idx[3] = {bin / (ny * nz), (bin / nz) % ny, bin % nz};
  const int off[3] = {0, cfg.dims[0], cfg.dims[0] + cfg.dims[1]};
  float c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double cen = b.centers[off[a] + idx[a]];
    const double low = cen + cfg.bin_size[a] / 2, high = cen - cfg.bin_size[a] / 2;  // curriculum.py:62-63
    c[a] = (float)(low + (high - low) * uc[a]);                                        // rng.uniform(low, high)
  }
  // small commands to zero (:622)
  const float keep = (sqrtf(c[0] * c[0] + c[1] * c[1]) > 0.2f) ? 1.f : 0.f;
  float* cmd = b.commands + (size_t)e * 4;
  cmd[0] = c[0] * keep; cmd[1] = c[1] * keep; cmd[2] = c[2];
  b.env_command_bins[e] = bin;
  for (int r = 0; r < cfg.n_command_sums; ++r) b.command_sums[r * Ns + e] = 0.f;  // :625-626
}

static int gac_validate(const RlGacCfg* cfg, const RlGacBuffers* b) {
  RL_REQUIRE(cfg && b, RL_ERR_BAD_ARG, "gac: null cfg/buffers");
  RL_REQUIRE((b->mask != nullptr) != (b->ids != nullptr), RL_ERR_BAD_ARG, "gac: exactly one of mask / ids");
  RL_REQUIRE(cfg->n_bins == cfg->dims[0] * cfg->dims[1] * cfg->dims[2] && cfg->n_bins > 0, RL_ERR_BAD_CFG,
             "gac: n_bins=%d does not match dims", cfg->n_bins);
  RL_REQUIRE(b->weights && b->centers && b->nbr_lo && b->nbr_hi && b->hit_count && b->own_flag && b->cdf &&
             b->env_command_bins && b->commands && b->command_sums, RL_ERR_BAD_ARG, "gac: a required buffer is null");
  RL_REQUIRE((b->u_bin == nullptr) == (b->u_cell == nullptr), RL_ERR_BAD_ARG, "gac: u_bin and u_cell go together");
  return RL_OK;
}

}  // namespace rl

using namespace rl;

extern "C" int rl_gac_scatter(const RlGacCfg* cfg, const RlGacBuffers* b, void* stream) {
  int rc = gac_validate(cfg, b);
  if (rc != RL_OK) return rc;
  const int n = b->ids ? b->n_ids : cfg->num_envs;
  if (n <= 0) return RL_OK;
  GacArgs args; args.cfg = *cfg; args.b = *b; args.seed = 0; args.step = 0;
  gac_scatter_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(args);
  return check_launch("gac_scatter_kernel");
}

extern "C" int rl_gac_update_sample(const RlGacCfg* cfg, const RlGacBuffers* b, uint64_t seed, uint64_t step,
                                    void* stream) {
  int rc = gac_validate(cfg, b);
  if (rc != RL_OK) return rc;
  const int n = b->ids ? b->n_ids : cfg->num_envs;
  if (n <= 0) return RL_OK;
  GacArgs args; args.cfg = *cfg; args.b = *b; args.seed = seed; args.step = step;
  gac_update_cdf_kernel<<<1, GAC_BLOCK, 0, (cudaStream_t)stream>>>(args);
  rc = check_launch("gac_update_cdf_kernel");
  if (rc != RL_OK) return rc;
  gac_sample_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(args);
  return check_launch("gac_sample_kernel");
}
