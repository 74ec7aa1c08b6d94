// The per-row part of the fused PPO loss (ppo.py:110-144): shared by the stand-alone loss kernel (ppo.cu) and by the
// teacher-forward chain kernel, whose value-output epilogue evaluates it for its own row (chain.cu, rl_chain_set_ppo_loss).
#pragma once
#include <cuda_bf16.h>

namespace rl {

constexpr int ACT = 12;      // num_actions
constexpr int LAT = 18;      // latent / privileged dim
constexpr int LROW = 40;     // per-row loss inputs: actions 12, mu_old 12, sigma_old 12, logp_old, adv, ret, v_old

// ---- fused PPO loss + gradients (ppo.py:110-144, 156-164) --------------------------------------------
// stats (double): [0] sum surrogate, [1] sum value loss, [2] sum kl, [3] sum adaptation sq. error
struct LossArgs {
  const float* mean;      // [B,12]  actor output
  const float* value;     // [B]     critic output
  const float* pred;      // [B,18]  adaptation module output (or null: skip the adaptation loss)
  const __nv_bfloat16* Xac; int ldac; int lat_off;   // latent target = Xac[:, lat_off : lat_off+18]
  const float* Lrow;      // [B,40]
  const float* std;       // [12]    learnable std (actor_critic.py:108)
  int B;
  float clip, value_coef, entropy_coef;
  int use_clipped_value;
  float inv_global_B;     // 1 / (B * world_size): means are taken over the global minibatch
  __nv_bfloat16* dmean;   // [B,16]
  __nv_bfloat16* dvalue;  // [B,8]
  __nv_bfloat16* dpred;   // [B,24]
  float* dstd;            // [12] gradient slot of std (atomic)
  double* stats;          // [4]
  float* kl_slot;         // optional fp32 KL sum (rides in the gradient all-reduce with several GPUs)
};

// One row: reads Lrow[i], mean[i], `v` = value[i]; writes dmean[i] (16 bf16) and dvalue[i] (8 bf16); returns the row's
// surrogate / value-loss / KL terms and its contribution to d loss / d std.
__device__ __forceinline__ void ppo_loss_row(const LossArgs& a, int i, float v, double& s_surr, double& s_val, double& s_kl,
                                             float (&g_std)[ACT]) {
    const float* L = a.Lrow + (size_t)i * LROW;
    const float adv = L[37], ret = L[38], v_old = L[39], logp_old = L[36];
    float logp = 0.f, kl = 0.f;
    float dmu[ACT], dsg[ACT];
#pragma unroll
    for (int d = 0; d < ACT; ++d) {
      const float mu = a.mean[(size_t)i * ACT + d], sg = a.std[d];
      const float act = L[d], mu_o = L[ACT + d], sg_o = L[2 * ACT + d];
      const float diff = act - mu;
      // Normal.log_prob (actor_critic.py:147)
      logp += -(diff * diff) / (2.f * sg * sg) - __logf(sg) - 0.9189385332046727f;
      // ppo.py:112-115
      kl += __logf(sg / sg_o + 1.e-5f) + (sg_o * sg_o + (mu_o - mu) * (mu_o - mu)) / (2.f * sg * sg) - 0.5f;
      dmu[d] = diff / (sg * sg);
      dsg[d] = diff * diff / (sg * sg * sg) - 1.f / sg;
    }
    // surrogate (:127-131); torch.max routes the gradient to the larger argument, halves on ties
    const float ratio = __expf(logp - logp_old);
    const float lo = 1.f - a.clip, hi = 1.f + a.clip;
    const float rc = fminf(fmaxf(ratio, lo), hi);
    const float s1 = -adv * ratio, s2 = -adv * rc;
    const float in_range = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
    float g_ratio;
    if (s1 > s2) g_ratio = -adv;
    else if (s1 < s2) g_ratio = -adv * in_range;
    else g_ratio = 0.5f * (-adv) + 0.5f * (-adv * in_range);
    s_surr = fmaxf(s1, s2);
    const float g_logp = g_ratio * ratio * a.inv_global_B;
    // value loss (:134-142)
    float vl, g_v;
    if (a.use_clipped_value) {
      const float dv = v - v_old;
      const float vc = v_old + fminf(fmaxf(dv, -a.clip), a.clip);
      const float l1 = (v - ret) * (v - ret), l2 = (vc - ret) * (vc - ret);
      const float inr = (dv >= -a.clip && dv <= a.clip) ? 1.f : 0.f;
      vl = fmaxf(l1, l2);
      if (l1 > l2) g_v = 2.f * (v - ret);
      else if (l1 < l2) g_v = 2.f * (vc - ret) * inr;
      else g_v = 0.5f * 2.f * (v - ret) + 0.5f * 2.f * (vc - ret) * inr;
    } else {
      vl = (ret - v) * (ret - v);
      g_v = 2.f * (v - ret);
    }
    s_val = vl; s_kl = kl;
    g_v *= a.value_coef * a.inv_global_B;
    // gradients w.r.t. the network outputs (bf16 operands of the backward GEMMs)
    __nv_bfloat16* dm = a.dmean + (size_t)i * 16;
#pragma unroll
    for (int d = 0; d < ACT; ++d) { dm[d] = __float2bfloat16(g_logp * dmu[d]); g_std[d] = g_logp * dsg[d]; }
#pragma unroll
    for (int d = ACT; d < 16; ++d) dm[d] = __float2bfloat16(0.f);
    __nv_bfloat16* dvp = a.dvalue + (size_t)i * 8;
    dvp[0] = __float2bfloat16(g_v);
#pragma unroll
    for (int d = 1; d < 8; ++d) dvp[d] = __float2bfloat16(0.f);
}

}  // namespace rl
