// Fused env step, one WARP PER LEG: the kernel behind rl_env_step_fused / rl_env_post_physics for the
// standard observation layout (gravity | commands | q | qd | actions [| heights], both shipped robots).
// Same arithmetic as env_step.cu (reference: mini_gym/envs/base/legged_robot.py:106-417, 1469-1646),
// different decomposition.
//
// Why: with one thread per env, 32768 envs are only 1024 warps for 148 SMs x 4 schedulers (ncu: 1.7
// warps per scheduler, every load / instruction-fetch stall exposed, ~4300 serial instructions per warp).
// Here a CTA of 4 warps owns 32 envs (lane = env) and warp w owns the three DOFs of leg w plus one
// slice of the per-env scalar work:
//   phase 1  all warps : PD torques + reward partial sums + q / qd / action observation columns (+noise)
//                        for their 3 DOFs, SoA loads / stores for those DOFs, their accumulator rows
//            warp 0    : frame transforms (base lin/ang velocity, projected gravity), teleport, push
//            warp 1    : contact terms: termination, collision, stumble, contact forces, feet air time
//            warp 2    : privileged-observation scalars
//   barrier
//   phase 2  warp 0    : combines the 4 partial sums, evaluates the enabled terms in reward_names order
//   barrier
//   phase 3  all warps : add the per-term rewards to their rows of episode_sums / command_sums, write their
//                        slices of the obs / priv / torque rows into the (re-used) input tile
//   barrier, cp.async.bulk stores
// 4x the warps, ~1/4 of the instruction stream per warp, ~1.5k SASS instructions in total (fits the
// instruction cache), <= 64 registers, 23 KB of shared memory per CTA -> 8 CTAs (32 warps) per SM.
#include <stdlib.h>

#include "env_common.cuh"

namespace rl {

constexpr int QT = 32;            // envs per CTA
constexpr int QTHREADS = 128;     // 4 warps
constexpr int NPART = 10;         // partial-sum kinds
enum Part { P_TQ2 = 0, P_ACC2, P_RATE2, P_LIM, P_ENERGY, P_ENERGY_POS, P_QD2, P_QD_LIM, P_TQ_LIM, P_STILL };
enum Frame { F_BLV = 0, F_BAV = 3, F_GRAV = 6, F_VWX = 9, F_VWY = 10, F_CMDN = 11, NFRAME = 12 };
enum Ct { C_COLL = 0, C_STUMBLE, C_FCF, C_AIR, C_RESET, C_TIMEOUT, NCT = 6 };
constexpr int NR = RL_MAX_TERMS;

// profiling aid (RL_ENV_TRACE=1): globaltimer stamps of one CTA's phases, read back by rl_debug_env_trace
__device__ unsigned long long g_env_trace[16];
__device__ int g_env_trace_on = 0;
__device__ __forceinline__ void env_stamp(int i) {
#ifdef RL_ENV_TRACE      // compiled out by default: even a predicated-off stamp costs a global load per thread
  if (g_env_trace_on && blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_env_trace[i] = t;
  }
#endif
}

// STD = the shipped training configuration (both robots): 'P' control, exactly the 12 default reward terms in
// their default order, observation noise on, positive-reward clip on, no pushes / heights / time-out resets /
// terminal body height / termination term / global reference.  Those switches become compile-time constants:
// the uniform branches, their constant-bank loads and the run-time term dispatch disappear (the kernel is
// instruction-issue bound: ~1600 warp instructions per 32 envs, 40 % of them integer / branch / constant loads).
constexpr uint32_t STD_TERM_MASK = 0xFFFu;

template <bool FUSE, int MINB, bool STD>
__global__ void __launch_bounds__(QTHREADS, MINB)
env_step_quad_kernel(const __grid_constant__ StepArgs args) {
  env_stamp(0);
  const RlEnvCfg& cfg = args.cfg;
  const int c_control = STD ? 0 : cfg.control_type;
  const bool c_push = STD ? false : cfg.push_robots != 0;
  const bool c_heights = STD ? false : cfg.measure_heights != 0;
  const bool c_noise = STD ? true : cfg.add_noise != 0;
  const bool c_timeout = STD ? false : cfg.timeout_resets != 0;
  const bool c_tbh = STD ? false : cfg.use_terminal_body_height != 0;
  const bool c_term = STD ? false : cfg.has_termination != 0;
  const bool c_pos = STD ? true : cfg.only_positive_rewards != 0;
  const bool c_global = STD ? false : cfg.global_reference != 0;
  const int c_n_terms = STD ? 12 : cfg.n_terms;
  const RlEnvBuffers& b = args.b;
  const int N = cfg.num_envs, NB = cfg.num_bodies;
  const int tile0 = blockIdx.x * QT;
  const int n_valid = min(QT, N - tile0);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int e = tile0 + lane;
  const bool valid = lane < n_valid;
  const int P = c_heights ? cfg.num_height_points : 0;
  constexpr int W = 42;
  const uint64_t rng_step = args.step + (b.step_state ? b.step_state[0] : 0ull);
  const uint32_t tmask = STD ? STD_TERM_MASK : cfg.term_mask;
  const float co = cfg.clip_obs;

  extern __shared__ __align__(16) float smem[];
  float* s_hmean = smem;                          // [32]
  float* s_root = s_hmean + QT;                   // [32][13]
  float* s_dof = s_root + QT * 13;                // [32][24]     } inputs; the output rows
  float* s_con = s_dof + QT * 24;                 // [32][NB*3]   } (obs 42, priv 18, torques 12)
  float* s_act = s_con + QT * NB * 3;             // [32][12]     } re-use this space in phase 3
  float* s_tq_in = s_act + QT * ND;               // [32][12]     (post_physics only)
  float* s_obs = s_dof;
  float* s_priv = s_obs + QT * W;
  float* s_tq = s_priv + QT * RL_PRIV_DIM;
  const int in_floats = 24 + NB * 3 + ND + ND, out_floats = W + RL_PRIV_DIM + ND;
  float* s_x = s_dof + QT * (in_floats > out_floats ? in_floats : out_floats);   // exchange area
  float* s_part = s_x;                            // [NPART][4][32]
  float* s_frame = s_part + NPART * 4 * QT;       // [NFRAME][32]
  float* s_ct = s_frame + NFRAME * QT;            // [NCT][32]
  float* s_r = s_ct + NCT * QT;                   // [NR][32]
  __shared__ int s_root_dirty;
  __shared__ __align__(8) uint64_t s_bar;

  // ---- stage the simulator-owned rows (one cp.async.bulk per tensor for full aligned tiles) ----------
  const float* g_root = b.root_states + (size_t)tile0 * 13;
  const float* g_dof = b.dof_state + (size_t)tile0 * 24;
  const float* g_con = b.contact_forces + (size_t)tile0 * NB * 3;
  const float* g_act = b.actions_in + (size_t)tile0 * ND;
  const float* g_tq = b.torques + (size_t)tile0 * ND;
  const bool bulk_in = (n_valid == QT) &&
      ((((uintptr_t)g_root | (uintptr_t)g_dof | (uintptr_t)g_con | (uintptr_t)g_act | (uintptr_t)g_tq) & 15) == 0);
  if (bulk_in) {
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      mbar_fence_init();
      mbar_expect_tx(&s_bar, (uint32_t)(QT * (13 + 24 + NB * 3 + ND + (FUSE ? 0 : ND)) * sizeof(float)));
      bulk_g2s(s_root, g_root, QT * 13 * 4, &s_bar);
      bulk_g2s(s_dof, g_dof, QT * 24 * 4, &s_bar);
      bulk_g2s(s_con, g_con, (uint32_t)(QT * NB * 3 * 4), &s_bar);
      bulk_g2s(s_act, g_act, QT * ND * 4, &s_bar);
      if (!FUSE) bulk_g2s(s_tq_in, g_tq, QT * ND * 4, &s_bar);
    }
  } else {
    stage_in<QTHREADS>(s_root, g_root, n_valid * 13);
    stage_in<QTHREADS>(s_dof, g_dof, n_valid * 24);
    stage_in<QTHREADS>(s_con, g_con, n_valid * NB * 3);
    stage_in<QTHREADS>(s_act, g_act, n_valid * ND);
    if (!FUSE) stage_in<QTHREADS>(s_tq_in, g_tq, n_valid * ND);
  }
  if (tid == 0) s_root_dirty = 0;

  // ---- early SoA loads: this warp's 3 DOFs and its accumulator rows ------------------------------------
  float kp[3], kd[3], ms[3], la[3], ldv[3];
  constexpr int MAXOWN = (RL_MAX_TERMS + 3) / 4;     // terms i = w, w+4, ... owned by warp w
  float es0 = 0.f, es1 = 0.f, es2 = 0.f, cs0 = 0.f, cs1 = 0.f, cs2 = 0.f;   // rows of the first three owned terms
  float ex0 = 0.f, ex1 = 0.f, ex2 = 0.f, ex3 = 0.f, ex4 = 0.f;   // role specific rows (warp 1: extras, warp 3: total/termination)
  int ep = 0;
  float4 cmd = make_float4(0.f, 0.f, 0.f, 0.f);
  float air[RL_NUM_FEET] = {0.f, 0.f, 0.f, 0.f};
  uint32_t last_contacts = 0;
  float sc6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // warp 2: friction, restitution, payload, com[3]
  const bool air_on = (tmask >> RL_REW_FEET_AIR_TIME) & 1u;
  const float* pe = b.episode_sums + e;
  const float* pc = b.command_sums + e;
  if (valid) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int ix = (3 * w + k) * N + e;
      kp[k] = b.Kp_factors[ix]; kd[k] = b.Kd_factors[ix]; ms[k] = b.motor_strengths[ix];
      la[k] = b.last_actions[ix]; ldv[k] = b.last_dof_vel[ix];
    }
    if (w < c_n_terms) { es0 = pe[w * N]; cs0 = pc[w * N]; }
    if (w + 4 < c_n_terms) { es1 = pe[(w + 4) * N]; cs1 = pc[(w + 4) * N]; }
    if (w + 8 < c_n_terms) { es2 = pe[(w + 8) * N]; cs2 = pc[(w + 8) * N]; }
    ep = (int)b.episode_length_buf[e];
    cmd = *reinterpret_cast<const float4*>(b.commands + (size_t)e * 4);
    if (w == 1) {
      if (air_on) {
#pragma unroll
        for (int k = 0; k < RL_NUM_FEET; ++k) air[k] = b.feet_air_time[k * N + e];
        last_contacts = *reinterpret_cast<const uint32_t*>(b.last_contacts + (size_t)e * 4);
      }
#pragma unroll
      for (int x = 0; x < 5; ++x) {
        const float v = pc[(RL_ROW_EXTRAS + x) * N];
        if (x == 0) ex0 = v; else if (x == 1) ex1 = v; else if (x == 2) ex2 = v; else if (x == 3) ex3 = v; else ex4 = v;
      }
    } else if (w == 2) {
      sc6[0] = b.friction_coeffs[e]; sc6[1] = b.restitutions[e]; sc6[2] = b.payloads[e];
#pragma unroll
      for (int k = 0; k < 3; ++k) sc6[3 + k] = b.com_displacements[k * N + e];
    } else if (w == 3) {
      ex0 = pe[RL_ROW_TOTAL * N];
      if (c_term) { ex1 = pe[RL_ROW_TERMINATION * N]; ex2 = pc[RL_ROW_TERMINATION * N]; }
    }
  }
  env_stamp(1);
  __syncthreads();                      // mbarrier init / cooperative stores visible
  if (b.step_state && tid == 96) {
    // device step counter: every thread of this CTA has read it (above the barrier), so the CTA is counted NOW and the
    // atomic's round trip overlaps with the tile's flight instead of ending the kernel (env_step_rows.cu, profiles/r02_env_step.md)
    const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(b.step_state + 1), 1ull);
    if (done == gridDim.x - 1) {
      b.step_state[1] = 0;
      atomicAdd(reinterpret_cast<unsigned long long*>(b.step_state), 1ull);
    }
  }
  if (bulk_in) mbar_wait(&s_bar, 0);    // all staged bytes have landed
  env_stamp(2);
  ep += 1;                              // :152

  // ---- teleport (:768-791) by warp 0, then the height phase if enabled ---------------------------------
  float* root = s_root + lane * 13;
  bool dirty = false;
  if (w == 0 && valid && cfg.teleport_robots) {
    float x = root[0], y = root[1];
    const float x0 = x, y0 = y;
    if (x < cfg.teleport_lo_x) x += cfg.teleport_shift_x;
    if (x > cfg.teleport_hi_x) x -= cfg.teleport_shift_x;
    if (y < cfg.teleport_lo_y) y += cfg.teleport_shift_y;
    if (y > cfg.teleport_hi_y) y -= cfg.teleport_shift_y;
    if (x != x0 || y != y0) { root[0] = x; root[1] = y; dirty = true; }
  }
  if (c_heights) {
    if (b.height_mean) {
      // sampled by the pre-pass launch (heights.cu), which applied the same teleport to its copy of the position
      if (w == 0 && valid) s_hmean[lane] = b.height_mean[e];
    } else {
      __syncthreads();
#pragma unroll 1
      for (int le = w; le < n_valid; le += 4) {      // one warp per env, lanes over the points (:1469-1503)
        const float* r = s_root + le * 13;
        const float hm = sample_heights_env(cfg, b, args.seed, rng_step, tile0 + le, r[0], r[1], r[2], r[5], r[6], lane, c_noise);
        if (lane == 0) s_hmean[le] = hm;
      }
    }
    __syncthreads();
  }

  env_stamp(3);
  // =================================== phase 1 ===========================================================
  float tq[3], oq[3], oqd[3], oa[3], pm[3];        // this leg's torques / observation / priv columns
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;              // warp 0: noisy gravity columns
  if (valid) {
    // ---- all warps: the three DOFs of leg w (:653-688 and the per-DOF reward sums) ----
    const float2 d0 = *reinterpret_cast<const float2*>(s_dof + lane * 24 + 6 * w);
    const float2 d1 = *reinterpret_cast<const float2*>(s_dof + lane * 24 + 6 * w + 2);
    const float2 d2 = *reinterpret_cast<const float2*>(s_dof + lane * 24 + 6 * w + 4);
    const float q[3] = {d0.x, d1.x, d2.x}, qd[3] = {d0.y, d1.y, d2.y};
    float part[NPART];
#pragma unroll
    for (int t = 0; t < NPART; ++t) part[t] = 0.f;
    const bool on_acc = (tmask >> RL_REW_DOF_ACC) & 1u, on_rate = (tmask >> RL_REW_ACTION_RATE) & 1u;
    const bool on_lim = (tmask >> RL_REW_DOF_POS_LIMITS) & 1u;
    const bool on_energy = ((tmask >> RL_REW_ENERGY) | (tmask >> RL_REW_ENERGY_EXPENDITURE)) & 1u;
    const bool on_qd2 = (tmask >> RL_REW_DOF_VEL) & 1u, on_qd_lim = (tmask >> RL_REW_DOF_VEL_LIMITS) & 1u;
    const bool on_tq_lim = (tmask >> RL_REW_TORQUE_LIMITS) & 1u, on_still = (tmask >> RL_REW_STAND_STILL) & 1u;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int j = 3 * w + k;
      const int ix = j * N + e;
      const float a = clampf(s_act[lane * ND + j], -cfg.clip_actions, cfg.clip_actions);   // :112-113
      float t;
      if (FUSE) {
        float as = a * cfg.action_scale;
        if (k == 0) as *= cfg.hip_scale_reduction;             // dofs 0,3,6,9 (:666)
        if (c_control == 0) {
          const float jpt = as + cfg.default_dof_pos[j];
          b.joint_pos_target[ix] = jpt;
          t = cfg.p_gains[j] * kp[k] * (jpt - q[k]) - cfg.d_gains[j] * kd[k] * qd[k];
        } else if (c_control == 1) {
          t = cfg.p_gains[j] * (as - qd[k]) - cfg.d_gains[j] * (qd[k] - ldv[k]) / cfg.sim_dt;
        } else {
          t = as;
        }
        t = t * ms[k];
        t = clampf(t, -cfg.torque_limits[j], cfg.torque_limits[j]);
      } else {
        t = s_tq_in[lane * ND + j];
      }
      tq[k] = t;
      part[P_TQ2] += sq(t);
      if (on_acc) part[P_ACC2] += sq((ldv[k] - qd[k]) / cfg.dt);
      if (on_rate) part[P_RATE2] += sq(la[k] - a);
      if (on_lim) {
        float ov = -fminf(q[k] - cfg.dof_pos_lo[j], 0.f);
        ov += fmaxf(q[k] - cfg.dof_pos_hi[j], 0.f);
        part[P_LIM] += ov;
      }
      if (on_energy) {
        const float pw = t * qd[k];
        part[P_ENERGY] += pw;
        part[P_ENERGY_POS] += clampf(pw, 0.f, 1e30f);
      }
      if (on_qd2) part[P_QD2] += sq(qd[k]);
      if (on_qd_lim) part[P_QD_LIM] += clampf(fabsf(qd[k]) - cfg.dof_vel_limits[j] * cfg.soft_dof_vel_limit, 0.f, 1.f);
      if (on_tq_lim) part[P_TQ_LIM] += fmaxf(fabsf(t) - cfg.torque_limits[j] * cfg.soft_torque_limit, 0.f);
      if (on_still) part[P_STILL] += fabsf(q[k] - cfg.default_dof_pos[j]);
      oq[k] = (q[k] - cfg.default_dof_pos[j]) * cfg.obs_scale_dof_pos;
      oqd[k] = qd[k] * cfg.obs_scale_dof_vel;
      oa[k] = a;
      b.last_actions[ix] = a;              // :181-182
      b.last_dof_vel[ix] = qd[k];
    }
#pragma unroll
    for (int t = 0; t < NPART; ++t) s_part[(t * 4 + w) * QT + lane] = part[t];

    // DOF-property re-draw (:591-593, :544-560): every warp redraws its own three DOFs from the same uniforms
    if ((ep % cfg.rand_interval) == 0 &&
        (cfg.randomize_motor_strength | cfg.randomize_Kp_factor | cfg.randomize_Kd_factor)) {
      float u3[4];
      if (b.dr_u) { u3[0] = b.dr_u[e]; u3[1] = b.dr_u[N + e]; u3[2] = b.dr_u[2 * N + e]; }
      else rng4(args.seed, (uint32_t)e, rng_step, RNG_DR, 0, u3);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int ix = (3 * w + k) * N + e;
        if (cfg.randomize_motor_strength) {
          ms[k] = u3[0] * cfg.motor_strength_lo_span[1] + cfg.motor_strength_lo_span[0];
          b.motor_strengths[ix] = ms[k];
        }
        if (cfg.randomize_Kp_factor) b.Kp_factors[ix] = u3[1] * cfg.Kp_factor_lo_span[1] + cfg.Kp_factor_lo_span[0];
        if (cfg.randomize_Kd_factor) b.Kd_factors[ix] = u3[2] * cfg.Kd_factor_lo_span[1] + cfg.Kd_factor_lo_span[0];
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) pm[k] = clampf((ms[k] - cfg.priv_shift[4]) * cfg.priv_scale[4], -co, co);

    // ---- observation noise for this leg's q / qd columns (:392): Philox block w, lanes 0-5 ----
    if (c_noise) {
      const float* nu = b.noise_u ? b.noise_u + (size_t)e * cfg.num_obs : nullptr;
      uint32_t r4[4] = {0u, 0u, 0u, 0u};
      if (!nu) Philox::gen(args.seed, (uint32_t)e, (uint32_t)rng_step, (uint32_t)(rng_step >> 32), (RNG_NOISE << 16) | (uint32_t)w, r4);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int cq = 6 + 3 * w + k, cqd = 18 + 3 * w + k;
        if (nu) {
          oq[k] += (2.0f * nu[cq] - 1.0f) * cfg.noise_scale_core[cq];
          oqd[k] += (2.0f * nu[cqd] - 1.0f) * cfg.noise_scale_core[cqd];
        } else {
          oq[k] = __fmaf_rn(2.0f * centered_u16(r4, k), cfg.noise_scale_core[cq], oq[k]);
          oqd[k] = __fmaf_rn(2.0f * centered_u16(r4, 3 + k), cfg.noise_scale_core[cqd], oqd[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { oq[k] = clampf(oq[k], -co, co); oqd[k] = clampf(oqd[k], -co, co); oa[k] = clampf(oa[k], -co, co); }

    if (w == 0) {
      // ---- frames (:159-162), push (:757-766) ----
      const float qx = root[3], qy = root[4], qz = root[5], qw = root[6];
      V3 vw = {root[7], root[8], root[9]};
      const V3 ww = {root[10], root[11], root[12]};
      const V3 blv = quat_rotate_inverse(qx, qy, qz, qw, vw);
      const V3 bav = quat_rotate_inverse(qx, qy, qz, qw, ww);
      const V3 grav = quat_rotate_inverse(qx, qy, qz, qw, V3{0.f, 0.f, -1.f});
      if (c_push && (ep % cfg.push_interval) == 0) {
        float u0, u1;
        if (b.push_u) { u0 = b.push_u[e]; u1 = b.push_u[N + e]; }
        else { float u4[4]; rng4(args.seed, (uint32_t)e, rng_step, RNG_PUSH, 0, u4); u0 = u4[0]; u1 = u4[1]; }
        vw.x = cfg.push_lo_span[1] * u0 + cfg.push_lo_span[0];
        vw.y = cfg.push_lo_span[1] * u1 + cfg.push_lo_span[0];
        root[7] = vw.x; root[8] = vw.y;
        dirty = true;
      }
      s_frame[(F_BLV + 0) * QT + lane] = blv.x; s_frame[(F_BLV + 1) * QT + lane] = blv.y; s_frame[(F_BLV + 2) * QT + lane] = blv.z;
      s_frame[(F_BAV + 0) * QT + lane] = bav.x; s_frame[(F_BAV + 1) * QT + lane] = bav.y; s_frame[(F_BAV + 2) * QT + lane] = bav.z;
      s_frame[(F_GRAV + 0) * QT + lane] = grav.x; s_frame[(F_GRAV + 1) * QT + lane] = grav.y; s_frame[(F_GRAV + 2) * QT + lane] = grav.z;
      s_frame[F_VWX * QT + lane] = vw.x; s_frame[F_VWY * QT + lane] = vw.y;
      b.base_lin_vel[0 * N + e] = blv.x; b.base_lin_vel[1 * N + e] = blv.y; b.base_lin_vel[2 * N + e] = blv.z;
      b.base_ang_vel[0 * N + e] = bav.x; b.base_ang_vel[1 * N + e] = bav.y; b.base_ang_vel[2 * N + e] = bav.z;
      b.projected_gravity[0 * N + e] = grav.x; b.projected_gravity[1 * N + e] = grav.y; b.projected_gravity[2 * N + e] = grav.z;
      b.last_root_vel[0 * N + e] = vw.x; b.last_root_vel[1 * N + e] = vw.y; b.last_root_vel[2 * N + e] = vw.z;
      b.last_root_vel[3 * N + e] = ww.x; b.last_root_vel[4 * N + e] = ww.y; b.last_root_vel[5 * N + e] = ww.z;
      // gravity observation columns 0-2 with noise: Philox block 4, lanes 0-2
      g0 = grav.x; g1 = grav.y; g2 = grav.z;
      if (c_noise) {
        const float* nu = b.noise_u ? b.noise_u + (size_t)e * cfg.num_obs : nullptr;
        if (nu) {
          g0 += (2.0f * nu[0] - 1.0f) * cfg.noise_scale_core[0];
          g1 += (2.0f * nu[1] - 1.0f) * cfg.noise_scale_core[1];
          g2 += (2.0f * nu[2] - 1.0f) * cfg.noise_scale_core[2];
        } else {
          uint32_t r4[4];
          Philox::gen(args.seed, (uint32_t)e, (uint32_t)rng_step, (uint32_t)(rng_step >> 32), (RNG_NOISE << 16) | 4u, r4);
          g0 = __fmaf_rn(2.0f * centered_u16(r4, 0), cfg.noise_scale_core[0], g0);
          g1 = __fmaf_rn(2.0f * centered_u16(r4, 1), cfg.noise_scale_core[1], g1);
          g2 = __fmaf_rn(2.0f * centered_u16(r4, 2), cfg.noise_scale_core[2], g2);
        }
      }
      g0 = clampf(g0, -co, co); g1 = clampf(g1, -co, co); g2 = clampf(g2, -co, co);
    } else if (w == 1) {
      // ---- contact terms: termination (:190-202), collision, stumble, contact forces, air time ----
      const float* con = s_con + lane * NB * 3;
      const float hmean = c_heights ? s_hmean[lane] : root[2];
      bool reset = false;
#pragma unroll 1
      for (int k = 0; k < cfg.n_term_bodies; ++k) {
        const float* f = con + cfg.term_idx[k] * 3;
        reset |= sqrtf((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]) > 1.0f;
      }
      bool time_out = false;
      if (c_timeout) {
        time_out = ep > cfg.max_episode_length;
        reset |= time_out;
        b.time_out_buf[e] = time_out ? 1 : 0;
      }
      if (c_tbh) reset |= hmean < cfg.terminal_body_height;
      b.reset_buf[e] = reset ? 1 : 0;
      b.episode_length_buf[e] = (int64_t)ep;
      float coll = 0.f;
#pragma unroll 1
      for (int k = 0; k < cfg.n_pen_bodies; ++k) {
        const float* f = con + cfg.pen_idx[k] * 3;
        coll += (sqrtf((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]) > 0.1f) ? 1.f : 0.f;
      }
      bool stumble = false;
      float fcf = 0.f;
      if (((tmask >> RL_REW_STUMBLE) | (tmask >> RL_REW_FEET_CONTACT_FORCES)) & 1u) {
#pragma unroll 1
        for (int k = 0; k < RL_NUM_FEET; ++k) {
          const float* f = con + cfg.feet_idx[k] * 3;
          stumble |= sqrtf(f[0] * f[0] + f[1] * f[1]) > 5.0f * fabsf(f[2]);
          fcf += fmaxf(sqrtf((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]) - cfg.max_contact_force, 0.f);
        }
      }
      float r_air = 0.f;
      if (air_on) {               // stateful (:1619-1631)
        uint32_t nc = 0;
#pragma unroll
        for (int k = 0; k < RL_NUM_FEET; ++k) {
          const bool contact = con[cfg.feet_idx[k] * 3 + 2] > 1.0f;
          const bool filt = contact || ((last_contacts >> (8 * k)) & 0xffu);
          nc |= (contact ? 1u : 0u) << (8 * k);
          const bool first = (air[k] > 0.f) && filt;
          air[k] += cfg.dt;
          r_air += (air[k] - 0.5f) * (first ? 1.f : 0.f);
          air[k] *= filt ? 0.f : 1.f;
          b.feet_air_time[k * N + e] = air[k];
        }
        *reinterpret_cast<uint32_t*>(b.last_contacts + (size_t)e * 4) = nc;
      }
      s_ct[C_COLL * QT + lane] = coll; s_ct[C_STUMBLE * QT + lane] = stumble ? 1.f : 0.f;
      s_ct[C_FCF * QT + lane] = fcf; s_ct[C_AIR * QT + lane] = r_air;
      s_ct[C_RESET * QT + lane] = reset ? 1.f : 0.f; s_ct[C_TIMEOUT * QT + lane] = time_out ? 1.f : 0.f;
    } else if (w == 2) {
      // ---- privileged-observation scalars (:398-417) + clip (:136) ----
      sc6[0] = clampf((sc6[0] - cfg.priv_shift[0]) * cfg.priv_scale[0], -co, co);
      sc6[1] = clampf((sc6[1] - cfg.priv_shift[1]) * cfg.priv_scale[1], -co, co);
      sc6[2] = clampf((sc6[2] - cfg.priv_shift[2]) * cfg.priv_scale[2], -co, co);
#pragma unroll
      for (int k = 0; k < 3; ++k) sc6[3 + k] = clampf((sc6[3 + k] - cfg.priv_shift[3]) * cfg.priv_scale[3], -co, co);
    }
  }
  env_stamp(4);
  __syncthreads();
  env_stamp(5);

  // =================================== phase 2 ===========================================================
  // Every warp evaluates ITS OWN reward terms (i = w, w+4, ...) from the exchanged partial sums, frames and
  // contact terms, adds them to its rows of episode_sums / command_sums and publishes r_i; it also writes
  // its slices of the output rows (every thread passed the barrier above: the input tile is dead).
  if (valid) {
    auto psum = [&](int t) {
      const float* p = s_part + t * 4 * QT + lane;
      return (p[0] + p[QT]) + (p[2 * QT] + p[3 * QT]);
    };
    auto F = [&](int i) { return s_frame[i * QT + lane]; };
    auto CT = [&](int i) { return s_ct[i * QT + lane]; };
    const float cmd_xy_norm = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y);
    auto eval_term = [&](int id) -> float {
      switch (id) {
        case RL_REW_TRACKING_LIN_VEL: {
          const float vx = c_global ? F(F_VWX) : F(F_BLV), vy = c_global ? F(F_VWY) : F(F_BLV + 1);
          return expf(-(sq(cmd.x - vx) + sq(cmd.y - vy)) / cfg.tracking_sigma);
        }
        case RL_REW_TRACKING_ANG_VEL: return expf(-sq(cmd.z - F(F_BAV + 2)) / cfg.tracking_sigma_yaw);
        case RL_REW_LIN_VEL_Z: return sq(F(F_BLV + 2));
        case RL_REW_ANG_VEL_XY: return sq(F(F_BAV)) + sq(F(F_BAV + 1));
        case RL_REW_ORIENTATION: return sq(F(F_GRAV)) + sq(F(F_GRAV + 1));
        case RL_REW_TORQUES: return psum(P_TQ2);
        case RL_REW_DOF_ACC: return psum(P_ACC2);
        case RL_REW_BASE_HEIGHT: return sq((c_heights ? s_hmean[lane] : root[2]) - cfg.base_height_target);
        case RL_REW_FEET_AIR_TIME: return CT(C_AIR) * ((cmd_xy_norm > 0.1f) ? 1.f : 0.f);
        case RL_REW_COLLISION: return CT(C_COLL);
        case RL_REW_ACTION_RATE: return psum(P_RATE2);
        case RL_REW_DOF_POS_LIMITS: return psum(P_LIM);
        case RL_REW_ENERGY: return psum(P_ENERGY);
        case RL_REW_ENERGY_EXPENDITURE: return psum(P_ENERGY_POS);
        case RL_REW_DOF_VEL: return psum(P_QD2);
        case RL_REW_SURVIVAL: return (CT(C_RESET) != 0.f && CT(C_TIMEOUT) == 0.f) ? 0.f : 1.f;
        case RL_REW_DOF_VEL_LIMITS: return psum(P_QD_LIM);
        case RL_REW_TORQUE_LIMITS: return psum(P_TQ_LIM);
        case RL_REW_STUMBLE: return CT(C_STUMBLE);
        case RL_REW_STAND_STILL: return psum(P_STILL) * (cmd_xy_norm < 0.1f ? 1.f : 0.f);
        case RL_REW_FEET_CONTACT_FORCES: return CT(C_FCF);
        default: return 0.f;
      }
    };
    float* pew = b.episode_sums + e;
    float* pcw = b.command_sums + e;
    if (STD) {
      // warp w owns terms w, w + 4, w + 8 - compile-time ids, the dispatch folds away
      float r0, r1, r2;
      if (w == 0) { r0 = eval_term(0); r1 = eval_term(4); r2 = eval_term(8); }
      else if (w == 1) { r0 = eval_term(1); r1 = eval_term(5); r2 = eval_term(9); }
      else if (w == 2) { r0 = eval_term(2); r1 = eval_term(6); r2 = eval_term(10); }
      else { r0 = eval_term(3); r1 = eval_term(7); r2 = eval_term(11); }
      r0 *= cfg.term_scale[w]; r1 *= cfg.term_scale[w + 4]; r2 *= cfg.term_scale[w + 8];
      pew[w * N] = es0 + r0; pcw[w * N] = cs0 + r0; s_r[w * QT + lane] = r0;
      pew[(w + 4) * N] = es1 + r1; pcw[(w + 4) * N] = cs1 + r1; s_r[(w + 4) * QT + lane] = r1;
      pew[(w + 8) * N] = es2 + r2; pcw[(w + 8) * N] = cs2 + r2; s_r[(w + 8) * QT + lane] = r2;
    } else
#pragma unroll 1
    for (int k = 0; k < MAXOWN; ++k) {
      const int i = w + 4 * k;
      if (i >= cfg.n_terms) break;
      const float r = eval_term(cfg.term_id[i]) * cfg.term_scale[i];
      float ev, cv;
      if (k == 0) { ev = es0; cv = cs0; } else if (k == 1) { ev = es1; cv = cs1; } else if (k == 2) { ev = es2; cv = cs2; }
      else { ev = pew[i * N]; cv = pcw[i * N]; }       // more than 12 enabled terms: read just in time
      pew[i * N] = ev + r;
      pcw[i * N] = cv + r;
      s_r[i * QT + lane] = r;
    }
    if (w == 1) {
      const float bx = F(F_BLV), wz = F(F_BAV + 2);
      pcw[(RL_ROW_EXTRAS + 0) * N] = ex0 + bx;                  // lin_vel_raw (:336-340)
      pcw[(RL_ROW_EXTRAS + 1) * N] = ex1 + wz;                  // ang_vel_raw
      pcw[(RL_ROW_EXTRAS + 2) * N] = ex2 + sq(bx - cmd.x);      // lin_vel_residual
      pcw[(RL_ROW_EXTRAS + 3) * N] = ex3 + sq(wz - cmd.z);      // ang_vel_residual
      pcw[(RL_ROW_EXTRAS + 4) * N] = ex4 + 1.0f;                // ep_timesteps
    }
    // output rows (re-using the input tile)
    float* obs = s_obs + lane * W;
    float* priv = s_priv + lane * RL_PRIV_DIM;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      obs[6 + 3 * w + k] = oq[k];
      obs[18 + 3 * w + k] = oqd[k];
      obs[30 + 3 * w + k] = oa[k];
      priv[6 + 3 * w + k] = pm[k];
      if (FUSE) s_tq[lane * ND + 3 * w + k] = tq[k];
    }
    if (w == 0) {
      obs[0] = g0; obs[1] = g1; obs[2] = g2;
      obs[3] = clampf(cmd.x * cfg.commands_scale[0], -co, co);
      obs[4] = clampf(cmd.y * cfg.commands_scale[1], -co, co);
      obs[5] = clampf(cmd.z * cfg.commands_scale[2], -co, co);
    } else if (w == 2) {
#pragma unroll
      for (int k = 0; k < 6; ++k) priv[k] = sc6[k];
    }
  }
  if (dirty) s_root_dirty = 1;
  env_stamp(6);
  fence_async_smem();
  __syncthreads();
  env_stamp(7);

  // =================================== phase 3: warp 3 closes compute_reward (:314-340) ===================
  if (w == 3 && valid) {
    float rew = 0.f;
    if (STD) {
#pragma unroll
      for (int i = 0; i < 12; ++i) rew += s_r[i * QT + lane];               // reward_names order
    } else {
#pragma unroll 1
      for (int i = 0; i < cfg.n_terms; ++i) rew += s_r[i * QT + lane];
    }
    if (b.rew_raw) b.rew_raw[e] = rew;
    if (c_pos) rew = fmaxf(rew, 0.f);
    b.episode_sums[RL_ROW_TOTAL * N + e] = ex0 + rew;
    if (c_term) {
      const bool term_flag = s_ct[C_RESET * QT + lane] != 0.f && s_ct[C_TIMEOUT * QT + lane] == 0.f;   // :1554
      const float r = (term_flag ? 1.f : 0.f) * cfg.termination_scale;
      rew += r;
      b.episode_sums[RL_ROW_TERMINATION * N + e] = ex1 + r;
      b.command_sums[RL_ROW_TERMINATION * N + e] = ex2 + r;
    }
    b.rew_buf[e] = rew;
  }

  // ---- write-back of the AoS outputs -----------------------------------------------------------------------
  float* o_obs = b.obs_buf + (size_t)tile0 * cfg.num_obs;
  float* o_priv = b.privileged_obs_buf + (size_t)tile0 * RL_PRIV_DIM;
  float* o_tq = b.torques + (size_t)tile0 * ND;
  float* o_root = b.root_states + (size_t)tile0 * 13;
  const bool bulk_out = (n_valid == QT) && (W == cfg.num_obs) &&
      ((((uintptr_t)o_obs | (uintptr_t)o_priv | (uintptr_t)o_tq | (uintptr_t)o_root) & 15) == 0);
  if (bulk_out) {
    if (tid == 0) {
      bulk_s2g(o_obs, s_obs, QT * W * 4);
      bulk_s2g(o_priv, s_priv, QT * RL_PRIV_DIM * 4);
      if (FUSE) bulk_s2g(o_tq, s_tq, QT * ND * 4);
      if (s_root_dirty) bulk_s2g(o_root, s_root, QT * 13 * 4);
      bulk_commit();
      bulk_wait_read0();
    }
  } else {
    if (W == cfg.num_obs) stage_out<QTHREADS>(o_obs, s_obs, n_valid * W);
    else stage_out_rows<QTHREADS>(o_obs, s_obs, n_valid, W, cfg.num_obs);
    stage_out<QTHREADS>(o_priv, s_priv, n_valid * RL_PRIV_DIM);
    if (FUSE) stage_out<QTHREADS>(o_tq, s_tq, n_valid * ND);
    if (s_root_dirty) stage_out<QTHREADS>(o_root, s_root, n_valid * 13);
  }
  env_stamp(8);
}

static size_t quad_smem_bytes(const RlEnvCfg& cfg) {
  const size_t in_floats = 24 + cfg.num_bodies * 3 + ND + ND, out_floats = 42 + RL_PRIV_DIM + ND;
  const size_t tile = 1 + 13 + (in_floats > out_floats ? in_floats : out_floats);
  const size_t xchg = NPART * 4 + NFRAME + NCT + NR;
  return (size_t)QT * (tile + xchg) * sizeof(float);
}

template <bool FUSE, int MINB, bool STD = false>
static int launch_quad_inst(const StepArgs& args, size_t smem, cudaStream_t st) {
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t err = cudaFuncSetAttribute(env_step_quad_kernel<FUSE, MINB, STD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    configured = smem;
  }
  const int grid = (args.cfg.num_envs + QT - 1) / QT;
  env_step_quad_kernel<FUSE, MINB, STD><<<grid, QTHREADS, smem, st>>>(args);
  return check_launch("env_step_quad_kernel");
}

}  // namespace rl
extern "C" int rl_debug_env_trace(int32_t enable, uint64_t* out_host16) {
  int on = enable;
  cudaMemcpyToSymbol(rl::g_env_trace_on, &on, sizeof(int));
  if (out_host16) cudaMemcpyFromSymbol(out_host16, rl::g_env_trace, 16 * sizeof(unsigned long long));
  return RL_OK;
}
namespace rl {
static int g_rows_mode = -1;
int g_rows_persist_mode = -1;      // env_step_rows.cu: -1 automatic, 0 one tile per CTA, 1 persistent
}
extern "C" int rl_debug_env_rows(int32_t mode) {
  const int prev = rl::g_rows_mode < 0 ? -1 : (rl::g_rows_mode == 0 ? 0 : (rl::g_rows_persist_mode == 1 ? 2 : (rl::g_rows_persist_mode == 0 ? 3 : (rl::g_rows_persist_mode == 2 ? 4 : 1))));
  rl::g_rows_mode = mode < 0 ? -1 : (mode ? 1 : 0);
  rl::g_rows_persist_mode = mode == 2 ? 1 : (mode == 3 ? 0 : (mode == 4 ? 2 : -1));    // 4: the 16-warp wide variant
  return prev;
}
namespace rl {
int launch_step_quad(const StepArgs& args, bool fuse_torques, cudaStream_t st) {
  const size_t smem = quad_smem_bytes(args.cfg);
  static int minb = 0;     // RL_QUAD_MINB=6: 85 registers / 6 CTAs per SM instead of 64 / 8 (tuning knob)
  if (!minb) { const char* m = getenv("RL_QUAD_MINB"); minb = (m && atoi(m) == 6) ? 6 : 8; }
  static int std_ok = -1;   // RL_ENV_STD=0 forces the generic instantiation
  if (std_ok < 0) { const char* e = getenv("RL_ENV_STD"); std_ok = (e && atoi(e) == 0) ? 0 : 1; }
  const RlEnvCfg& c = args.cfg;
  bool is_std = std_ok && c.control_type == 0 && !c.push_robots && !c.measure_heights && c.add_noise && !c.timeout_resets &&
                !c.use_terminal_body_height && !c.has_termination && c.only_positive_rewards && !c.global_reference &&
                c.n_terms == 12 && c.term_mask == STD_TERM_MASK;
  for (int i = 0; is_std && i < 12; ++i) is_std = c.term_id[i] == i;
  if (is_std) {
    // packed state blocks + full aligned tiles: every byte of the tile moves on the async proxy (env_step_rows.cu)
    static int rows_env = -1;   // RL_ENV_ROWS=0 keeps the kernel below
    if (rows_env < 0) { const char* e = getenv("RL_ENV_ROWS"); rows_env = (e && atoi(e) == 0) ? 0 : 1; }
    const int rows_on = g_rows_mode < 0 ? rows_env : g_rows_mode;
    if (rows_on && rows_layout_ok(args)) return launch_step_rows(args, fuse_torques, st);
    // small grids (<= 4 CTAs per SM) are pure latency: 6 CTAs / SM (80 registers, no spills) wins; otherwise
    // 7 CTAs / SM (72 registers; 7 x 148 = 1036 CTAs still hold 32768 envs in one wave).  Measured per launch,
    // 6 / 7 / 8 CTAs per SM: 4000 envs 6.9 / 7.4 / 7.9 us, 32768 envs 16.7 / 13.3 / 14.0 us, 262144 envs 81 / 73 / 80 us
    static int force = -1;
    if (force < 0) { const char* m = getenv("RL_QUAD_MINB"); force = m ? atoi(m) : 0; }
    const int n_cta = (c.num_envs + QT - 1) / QT;
    const bool use6 = force == 6 || (force != 8 && n_cta <= 4 * 148);
    if (use6) return fuse_torques ? launch_quad_inst<true, 6, true>(args, smem, st) : launch_quad_inst<false, 6, true>(args, smem, st);
    if (force == 8) return fuse_torques ? launch_quad_inst<true, 8, true>(args, smem, st) : launch_quad_inst<false, 8, true>(args, smem, st);
    return fuse_torques ? launch_quad_inst<true, 7, true>(args, smem, st) : launch_quad_inst<false, 7, true>(args, smem, st);
  }
  // (the height-sampling phase keeps 18 gathers per lane in flight: it needs the 80-register variant)
  if (minb == 6 || c.measure_heights) return fuse_torques ? launch_quad_inst<true, 6>(args, smem, st) : launch_quad_inst<false, 6>(args, smem, st);
  return fuse_torques ? launch_quad_inst<true, 8>(args, smem, st) : launch_quad_inst<false, 8>(args, smem, st);
}

}  // namespace rl
