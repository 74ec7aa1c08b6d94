// Fused per-step env pipeline of LeggedRobot (mini_gym/envs/base/legged_robot.py).
//
// One launch replaces ~415 ATen ops of LeggedRobot.step (:106-137):
//   _compute_torques :653-688, post_physics_step :139-188, _teleport_robots :768-791,
//   _get_heights :1469-1503, _push_robots :757-766, DOF-property re-draw :591-593,
//   check_termination :190-202, compute_reward :314-340 with every _reward_* :1506-1646,
//   compute_observations :342-417, last_* updates :181-183, observation clip :133-136.
//
// Mapping to B200
//   * one CTA = one tile of TILE=128 consecutive envs, one thread per env for the scalar
//     pipeline; grid = ceil(N/128) (256 CTAs at 32768 envs, 2 resident per SM).
//   * the simulator-owned tensors are AoS rows (13/24/3*NB/12 floats per env).  The CTA
//     copies the tile's contiguous row span into shared memory with 128-bit coalesced
//     streaming loads (or one cp.async.bulk per tensor, see stage_bulk), then each thread
//     reads its own row from shared memory (odd row strides: conflict free).
//   * state we own is SoA [K][N]: thread n touches p[k*N+n], i.e. every access is a fully
//     coalesced 128 B line per warp.
//   * outputs with AoS consumers (obs, privileged obs, torques) are assembled in shared
//     memory and written back with coalesced 128-bit stores.
//   * height sampling: one warp per env, lanes over the 187 points, int16 gathers served
//     from L2 (the 9.4 MB table stays resident), warp-shuffle reduction for the mean.
// Bound: HBM.  Algorithmic bytes per env-step: 1425 B (Mini Cheetah flat).
//
// Numerics: compiled with -fmad=false so every fp32 operation rounds exactly like the
// op-by-op eager reference; integer results (cell indices, resets, counters) are exact.
#include <stdlib.h>

#include "env_common.cuh"

namespace rl {

// Observation noise (:392): obs += (2*u - 1) * scale.  Test mode reads u from the injected
// tensor and keeps the reference's arithmetic exactly; product mode draws 8 sixteen-bit uniforms
// per Philox4x32-10 block (counter = noisy-value index / 8).
struct ObsNoise {
  const float* inj;   // noise_u row or nullptr
  uint64_t seed, step;
  uint32_t env;
  uint32_t r[4];
  int have;           // Philox block currently held (-1: none)
  __device__ ObsNoise(const float* inj_row, uint64_t s, uint64_t st, uint32_t e)
      : inj(inj_row), seed(s), step(st), env(e), have(-1) {}
  // k = running index over noisy values (static after unrolling), c = observation column
  __device__ __forceinline__ float apply(float v, float scale, int k, int c) {
    if (inj) return v + (2.0f * inj[c] - 1.0f) * scale;
    return __fmaf_rn(2.0f * centered_u16(r, k & 7), scale, v);
  }
  __device__ __noinline__ void refill(int blk) {
    Philox::gen(seed, env, (uint32_t)step, (uint32_t)(step >> 32), (RNG_NOISE << 16) | (uint32_t)blk, r);
    have = blk;
  }
  __device__ __forceinline__ void need(int k) {
    if (!inj && (k >> 3) != have) refill(k >> 3);
  }
};

// ---------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------
// STD_OBS: the observation layout of both shipped robots (gravity 3 | commands 3 | q 12 | qd 12 |
// actions 12, optionally followed by height samples) with every column index known at compile
// time; the generic instantiation handles the other observe_* combinations.
template <bool FUSE_TORQUES, int TILE, bool STD_OBS>
__global__ void __launch_bounds__(TILE, 512 / TILE)
env_step_kernel(const __grid_constant__ StepArgs args) {
  const RlEnvCfg& cfg = args.cfg;
  const RlEnvBuffers& b = args.b;
  const int N = cfg.num_envs;
  const int NB = cfg.num_bodies;
  const int tile0 = blockIdx.x * TILE;
  const int n_valid = min(TILE, N - tile0);
  const int tid = threadIdx.x;
  const int e = tile0 + tid;
  const bool valid = tid < n_valid;

  const int P = cfg.measure_heights ? cfg.num_height_points : 0;
  const int W = STD_OBS ? 42 : cfg.num_obs - P;  // width of the non-height part of the observation
  // RNG step key: host argument (+ device counter when replayed from a CUDA graph)
  const uint64_t rng_step = args.step + (b.step_state ? b.step_state[0] : 0ull);

  // Shared memory: [hmean | root | union{ inputs: dof, contact, actions, (torques in) ;
  //                                         outputs: obs, priv, torques out }].
  // The output rows are assembled in registers and only written after a CTA barrier, when every
  // thread is done with the input rows, so they can reuse that space (22.5 KB per 64 envs instead
  // of 41 KB: twice the resident CTAs).
  extern __shared__ __align__(16) float smem[];
  float* s_hmean = smem;                        // [TILE] mean(z - h)
  float* s_root = s_hmean + TILE;               // [TILE][13]
  float* s_dof = s_root + TILE * 13;            // [TILE][24]
  float* s_con = s_dof + TILE * 24;             // [TILE][NB*3]
  float* s_act = s_con + TILE * NB * 3;         // [TILE][12]
  float* s_tq_in = s_act + TILE * ND;           // [TILE][12] (post_physics only)
  float* s_obs = s_dof;                         // [TILE][W]   (aliases the inputs)
  float* s_priv = s_obs + TILE * W;             // [TILE][18]
  float* s_tq = s_priv + TILE * RL_PRIV_DIM;    // [TILE][12]
  __shared__ int s_root_dirty;
  __shared__ __align__(8) uint64_t s_bar;

  // ---- 1. stage the simulator-owned rows of this tile into shared memory ---------------
  // Full tiles of 16 B-aligned tensors: one elected thread issues one cp.async.bulk (TMA 1-D,
  // SASS UBLKCP) per tensor; the bytes land asynchronously while every thread issues its SoA
  // loads.  Ragged tail tile / unaligned views: cooperative 128-bit loads.
  const float* g_root = b.root_states + (size_t)tile0 * 13;
  const float* g_dof = b.dof_state + (size_t)tile0 * 24;
  const float* g_con = b.contact_forces + (size_t)tile0 * NB * 3;
  const float* g_act = b.actions_in + (size_t)tile0 * ND;
  const float* g_tq = b.torques + (size_t)tile0 * ND;
  const bool bulk_in = (n_valid == TILE) &&
      ((((uintptr_t)g_root | (uintptr_t)g_dof | (uintptr_t)g_con | (uintptr_t)g_act | (uintptr_t)g_tq) & 15) == 0);
  if (bulk_in) {
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      mbar_fence_init();
      const uint32_t bytes = (uint32_t)(TILE * (13 + 24 + NB * 3 + ND + (FUSE_TORQUES ? 0 : ND)) * sizeof(float));
      mbar_expect_tx(&s_bar, bytes);
      bulk_g2s(s_root, g_root, TILE * 13 * 4, &s_bar);
      bulk_g2s(s_dof, g_dof, TILE * 24 * 4, &s_bar);
      bulk_g2s(s_con, g_con, (uint32_t)(TILE * NB * 3 * 4), &s_bar);
      bulk_g2s(s_act, g_act, TILE * ND * 4, &s_bar);
      if (!FUSE_TORQUES) bulk_g2s(s_tq_in, g_tq, TILE * ND * 4, &s_bar);
    }
  } else {
    stage_in<TILE>(s_root, g_root, n_valid * 13);
    stage_in<TILE>(s_dof, g_dof, n_valid * 24);
    stage_in<TILE>(s_con, g_con, n_valid * NB * 3);
    stage_in<TILE>(s_act, g_act, n_valid * ND);
    if (!FUSE_TORQUES) stage_in<TILE>(s_tq_in, g_tq, n_valid * ND);
  }
  if (tid == 0) s_root_dirty = 0;

  // ---- 2. issue the SoA state loads early so they overlap the staging ------------------
  // (32-bit element indices: validated N * RL_COMMAND_ROWS < 2^31)
  const uint32_t tmask = cfg.term_mask;
  const bool air_on = (tmask >> RL_REW_FEET_AIR_TIME) & 1u;
  float kp[ND], kd[ND], ms[ND], la[ND], ldv[ND];
  float air[RL_NUM_FEET] = {0.f, 0.f, 0.f, 0.f};
  float fr = 0.f, re = 0.f, pl = 0.f, com[3] = {0.f, 0.f, 0.f};
  uint32_t last_contacts = 0;
  int64_t ep_len = 0;
  float4 cmd = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) {
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const int ix = j * N + e;
      kp[j] = b.Kp_factors[ix];
      kd[j] = b.Kd_factors[ix];
      ms[j] = b.motor_strengths[ix];
      la[j] = b.last_actions[ix];
      ldv[j] = b.last_dof_vel[ix];
    }
    if (air_on) {
#pragma unroll
      for (int k = 0; k < RL_NUM_FEET; ++k) air[k] = b.feet_air_time[k * N + e];
      last_contacts = *reinterpret_cast<const uint32_t*>(b.last_contacts + (size_t)e * 4);
    }
    fr = b.friction_coeffs[e]; re = b.restitutions[e]; pl = b.payloads[e];
#pragma unroll
    for (int k = 0; k < 3; ++k) com[k] = b.com_displacements[k * N + e];
    ep_len = b.episode_length_buf[e];
    cmd = *reinterpret_cast<const float4*>(b.commands + (size_t)e * 4);
  }
  __syncthreads();                      // mbarrier init / cooperative stores visible
  if (bulk_in) mbar_wait(&s_bar, 0);    // all staged bytes have landed

  // ---- 3. counters, teleport (:152, :768-791) --------------------------------------------
  float* root = s_root + tid * 13;
  bool dirty = false;
  if (valid) {
    ep_len += 1;
    // (the range's own thresholds under a train / eval split: legged_robot.py:576 through _call_train_eval)
    float x = root[0], y = root[1];
    if (teleport_xy_env(cfg, e, x, y)) { root[0] = x; root[1] = y; dirty = true; }
  }
  if (cfg.measure_heights) __syncthreads();  // teleported positions visible to the height warps

  // ---- 4. terrain heights (:1469-1503): one warp per env, lanes over points -------------
  if (cfg.measure_heights) {
    const int warp = tid >> 5, lane = tid & 31;
    if (b.height_mean) {
      // sampled by the pre-pass launch (heights.cu), which applied the same teleport to its copy of the position
      if (valid) s_hmean[tid] = b.height_mean[e];
    } else {
#pragma unroll 1
      for (int le = warp; le < n_valid; le += TILE / 32) {
        const float* r = s_root + le * 13;
        const float hm = sample_heights_env(cfg, b, args.seed, rng_step, tile0 + le, r[0], r[1], r[2], r[5], r[6], lane,
                                            cfg.add_noise != 0);
        if (lane == 0) s_hmean[le] = hm;
      }
    }
    __syncthreads();
  }

  // ---- 5. per-env scalar pipeline: everything lands in registers (o, pv, tq) ----------------
  float o[STD_OBS ? 42 : 1];      // observation row (standard layout)
  float pv[RL_PRIV_DIM];          // privileged observation row
  float tq[ND];                   // torques
  float dof[2 * ND], act[ND];     // this env's dof_state / action rows
  V3 blv = {0.f, 0.f, 0.f}, bav = blv, grav = blv, vw = blv;
  float qx = 0.f, qy = 0.f, qz = 0.f, qw = 1.f;
  const float co = cfg.clip_obs;
  if (valid) {
    const float* con = s_con + tid * NB * 3;

    // rows of 24 / 12 floats: 128-bit shared loads (a scalar walk would be 8-way bank conflicted)
#pragma unroll
    for (int k = 0; k < 6; ++k)
      *reinterpret_cast<float4*>(dof + 4 * k) = *reinterpret_cast<const float4*>(s_dof + tid * 24 + 4 * k);
#pragma unroll
    for (int k = 0; k < 3; ++k)
      *reinterpret_cast<float4*>(act + 4 * k) = *reinterpret_cast<const float4*>(s_act + tid * ND + 4 * k);
    if (!FUSE_TORQUES) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        *reinterpret_cast<float4*>(tq + 4 * k) = *reinterpret_cast<const float4*>(s_tq_in + tid * ND + 4 * k);
    }

    qx = root[3]; qy = root[4]; qz = root[5]; qw = root[6];
    vw = V3{root[7], root[8], root[9]};
    const V3 ww = {root[10], root[11], root[12]};
    blv = quat_rotate_inverse(qx, qy, qz, qw, vw);
    bav = quat_rotate_inverse(qx, qy, qz, qw, ww);
    grav = quat_rotate_inverse(qx, qy, qz, qw, V3{0.f, 0.f, -1.f});

    // push (:757-766) - after the body-frame velocity was taken (:160 precedes :588)
    const EnvVariant var = env_variant(cfg, e);          // train / eval split: this env's range (:588, :593)
    if (var.push_robots && (ep_len % var.push_interval) == 0) {
      float u0, u1;
      if (b.push_u) { u0 = b.push_u[e]; u1 = b.push_u[N + e]; }
      else { float u4[4]; rng4(args.seed, (uint32_t)e, rng_step, RNG_PUSH, 0, u4); u0 = u4[0]; u1 = u4[1]; }
      vw.x = var.push_lo_span[1] * u0 + var.push_lo_span[0];
      vw.y = var.push_lo_span[1] * u1 + var.push_lo_span[0];
      root[7] = vw.x; root[8] = vw.y;
      dirty = true;
    }

    // per-DOF sweep: torques (:653-688) and the partial sums of the ENABLED reward terms
    const bool on_energy = ((tmask >> RL_REW_ENERGY) | (tmask >> RL_REW_ENERGY_EXPENDITURE)) & 1u;
    const bool on_qd2 = (tmask >> RL_REW_DOF_VEL) & 1u, on_qd_lim = (tmask >> RL_REW_DOF_VEL_LIMITS) & 1u;
    const bool on_tq_lim = (tmask >> RL_REW_TORQUE_LIMITS) & 1u, on_still = (tmask >> RL_REW_STAND_STILL) & 1u;
    const bool on_acc = (tmask >> RL_REW_DOF_ACC) & 1u, on_rate = (tmask >> RL_REW_ACTION_RATE) & 1u;
    const bool on_lim = (tmask >> RL_REW_DOF_POS_LIMITS) & 1u;
    float sum_tq2 = 0.f, sum_acc2 = 0.f, sum_rate2 = 0.f, sum_lim = 0.f, sum_energy = 0.f,
          sum_energy_pos = 0.f, sum_qd2 = 0.f, sum_qd_lim = 0.f, sum_tq_lim = 0.f, sum_still = 0.f;
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const int ix = j * N + e;
      const float q = dof[2 * j], qd = dof[2 * j + 1];
      const float a = clampf(act[j], -cfg.clip_actions, cfg.clip_actions);  // :112-113
      act[j] = a;
      if (FUSE_TORQUES) {
        float t;
        float as = a * cfg.action_scale;
        if (j % 3 == 0) as *= cfg.hip_scale_reduction;  // dofs 0,3,6,9 (:666)
        if (cfg.control_type == 0) {
          const float jpt = as + cfg.default_dof_pos[j];
          b.joint_pos_target[ix] = jpt;
          t = cfg.p_gains[j] * kp[j] * (jpt - q) - cfg.d_gains[j] * kd[j] * qd;
        } else if (cfg.control_type == 1) {
          t = cfg.p_gains[j] * (as - qd) - cfg.d_gains[j] * (qd - ldv[j]) / cfg.sim_dt;
        } else {
          t = as;
        }
        t = t * ms[j];
        tq[j] = clampf(t, -cfg.torque_limits[j], cfg.torque_limits[j]);
      }
      const float t = tq[j];
      sum_tq2 += sq(t);
      if (on_acc) sum_acc2 += sq((ldv[j] - qd) / cfg.dt);
      if (on_rate) sum_rate2 += sq(la[j] - a);
      if (on_lim) {
        float ov = -fminf(q - cfg.dof_pos_lo[j], 0.f);
        ov += fmaxf(q - cfg.dof_pos_hi[j], 0.f);
        sum_lim += ov;
      }
      if (on_energy) {
        const float pw = t * qd;
        sum_energy += pw;
        sum_energy_pos += clampf(pw, 0.f, 1e30f);
      }
      if (on_qd2) sum_qd2 += sq(qd);
      if (on_qd_lim) sum_qd_lim += clampf(fabsf(qd) - cfg.dof_vel_limits[j] * cfg.soft_dof_vel_limit, 0.f, 1.f);
      if (on_tq_lim) sum_tq_lim += fmaxf(fabsf(t) - cfg.torque_limits[j] * cfg.soft_torque_limit, 0.f);
      if (on_still) sum_still += fabsf(q - cfg.default_dof_pos[j]);
      // last_* updates (:181-182)
      b.last_actions[ix] = a;
      b.last_dof_vel[ix] = qd;
    }

    // DOF-property re-draw for envs whose episode clock hits the interval (:591-593,:544-560)
    if ((ep_len % cfg.rand_interval) == 0 &&
        (var.randomize_motor_strength | var.randomize_Kp_factor | var.randomize_Kd_factor)) {
      float u3[4];
      if (b.dr_u) { u3[0] = b.dr_u[e]; u3[1] = b.dr_u[N + e]; u3[2] = b.dr_u[2 * N + e]; }
      else rng4(args.seed, (uint32_t)e, rng_step, RNG_DR, 0, u3);
      if (var.randomize_motor_strength) {
        const float v = u3[0] * var.motor_strength_lo_span[1] + var.motor_strength_lo_span[0];
#pragma unroll
        for (int j = 0; j < ND; ++j) { ms[j] = v; b.motor_strengths[j * N + e] = v; }
      }
      if (var.randomize_Kp_factor) {
        const float v = u3[1] * var.Kp_factor_lo_span[1] + var.Kp_factor_lo_span[0];
#pragma unroll 1
        for (int j = 0; j < ND; ++j) b.Kp_factors[j * N + e] = v;
      }
      if (var.randomize_Kd_factor) {
        const float v = u3[2] * var.Kd_factor_lo_span[1] + var.Kd_factor_lo_span[0];
#pragma unroll 1
        for (int j = 0; j < ND; ++j) b.Kd_factors[j * N + e] = v;
      }
    }

    // ---- privileged observations (:398-417) + clip (:136) ------------------------------------------
    pv[0] = clampf((fr - cfg.priv_shift[0]) * cfg.priv_scale[0], -co, co);
    pv[1] = clampf((re - cfg.priv_shift[1]) * cfg.priv_scale[1], -co, co);
    pv[2] = clampf((pl - cfg.priv_shift[2]) * cfg.priv_scale[2], -co, co);
#pragma unroll
    for (int k = 0; k < 3; ++k) pv[3 + k] = clampf((com[k] - cfg.priv_shift[3]) * cfg.priv_scale[3], -co, co);
#pragma unroll
    for (int j = 0; j < ND; ++j) pv[6 + j] = clampf((ms[j] - cfg.priv_shift[4]) * cfg.priv_scale[4], -co, co);

    // ---- termination (:190-202) --------------------------------------------------------------
    const float hmean = cfg.measure_heights ? s_hmean[tid] : root[2];  // mean(z - measured_heights)
    bool reset = false;
#pragma unroll 1
    for (int k = 0; k < cfg.n_term_bodies; ++k) {
      const float* f = con + cfg.term_idx[k] * 3;
      const float nrm = sqrtf((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]);
      reset |= nrm > 1.0f;
    }
    bool time_out = false;
    if (cfg.timeout_resets) {
      time_out = ep_len > (int64_t)cfg.max_episode_length;
      reset |= time_out;
      b.time_out_buf[e] = time_out ? 1 : 0;
    }
    if (cfg.use_terminal_body_height) reset |= hmean < cfg.terminal_body_height;
    b.reset_buf[e] = reset ? 1 : 0;

    // ---- reward terms (:1506-1646) -----------------------------------------------------------
    const float cmd_xy_norm = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y);
    float term_val[RL_REW_COUNT];
    {
      const float vx = cfg.global_reference ? vw.x : blv.x, vy = cfg.global_reference ? vw.y : blv.y;
      const float lin_err = sq(cmd.x - vx) + sq(cmd.y - vy);
      term_val[RL_REW_TRACKING_LIN_VEL] = expf(-lin_err / cfg.tracking_sigma);
      term_val[RL_REW_TRACKING_ANG_VEL] = expf(-sq(cmd.z - bav.z) / cfg.tracking_sigma_yaw);
      term_val[RL_REW_LIN_VEL_Z] = sq(blv.z);
      term_val[RL_REW_ANG_VEL_XY] = sq(bav.x) + sq(bav.y);
      term_val[RL_REW_ORIENTATION] = sq(grav.x) + sq(grav.y);
      term_val[RL_REW_TORQUES] = sum_tq2;
      term_val[RL_REW_DOF_ACC] = sum_acc2;
      term_val[RL_REW_BASE_HEIGHT] = sq(hmean - cfg.base_height_target);
      term_val[RL_REW_ACTION_RATE] = sum_rate2;
      term_val[RL_REW_DOF_POS_LIMITS] = sum_lim;
      term_val[RL_REW_ENERGY] = sum_energy;
      term_val[RL_REW_ENERGY_EXPENDITURE] = sum_energy_pos;
      term_val[RL_REW_DOF_VEL] = sum_qd2;
      term_val[RL_REW_DOF_VEL_LIMITS] = sum_qd_lim;
      term_val[RL_REW_TORQUE_LIMITS] = sum_tq_lim;
      term_val[RL_REW_STAND_STILL] = sum_still * (cmd_xy_norm < 0.1f ? 1.f : 0.f);
      const bool term_flag = reset && !time_out;  // :1554 reset_buf * ~time_out_buf
      term_val[RL_REW_TERMINATION] = term_flag ? 1.f : 0.f;
      term_val[RL_REW_SURVIVAL] = term_flag ? 0.f : 1.f;
      float coll = 0.f;
#pragma unroll 1
      for (int k = 0; k < cfg.n_pen_bodies; ++k) {
        const float* f = con + cfg.pen_idx[k] * 3;
        const float nrm = sqrtf((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]);
        coll += (nrm > 0.1f) ? 1.f : 0.f;
      }
      term_val[RL_REW_COLLISION] = coll;
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
      term_val[RL_REW_STUMBLE] = stumble ? 1.f : 0.f;
      term_val[RL_REW_FEET_CONTACT_FORCES] = fcf;
      term_val[RL_REW_FEET_AIR_TIME] = 0.f;
    }
    // stateful feet_air_time term (:1619-1631); its state only advances when the term is enabled
    if (air_on) {
      uint32_t nc = 0;
      float r_air = 0.f;
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
      r_air *= (cmd_xy_norm > 0.1f) ? 1.f : 0.f;
      term_val[RL_REW_FEET_AIR_TIME] = r_air;
    }

    // ---- compute_reward (:314-340): ordered accumulation ----------------------------------------
    // Accumulator rows (fixed layout, rl_b200.h) are read here, just in time, in two batches, so
    // that they do not occupy registers during the sweep; other resident warps cover the latency.
    {
      float* pe = b.episode_sums + e;
      float* pc = b.command_sums + e;
      float rew = 0.f;
      constexpr int HALF = (RL_MAX_TERMS + 1) / 2;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float ev[HALF], cv[HALF];
#pragma unroll
        for (int k = 0; k < HALF; ++k) {
          const int i = h * HALF + k;
          if (i < RL_MAX_TERMS && i < cfg.n_terms) { ev[k] = pe[i * N]; cv[k] = pc[i * N]; }
        }
#pragma unroll
        for (int k = 0; k < HALF; ++k) {
          const int i = h * HALF + k;
          if (i < RL_MAX_TERMS && i < cfg.n_terms) {
            const float r = term_val[cfg.term_id[i]] * cfg.term_scale[i];
            rew += r;
            pe[i * N] = ev[k] + r;
            pc[i * N] = cv[k] + r;
          }
        }
      }
      float e_tot = pe[RL_ROW_TOTAL * N], e_term = 0.f, c_term = 0.f, cx[5];
      if (cfg.has_termination) { e_term = pe[RL_ROW_TERMINATION * N]; c_term = pc[RL_ROW_TERMINATION * N]; }
#pragma unroll
      for (int x = 0; x < 5; ++x) cx[x] = pc[(RL_ROW_EXTRAS + x) * N];
      if (b.rew_raw) b.rew_raw[e] = rew;
      if (cfg.only_positive_rewards) rew = fmaxf(rew, 0.f);
      pe[RL_ROW_TOTAL * N] = e_tot + rew;
      if (cfg.has_termination) {
        const float r = term_val[RL_REW_TERMINATION] * cfg.termination_scale;
        rew += r;
        pe[RL_ROW_TERMINATION * N] = e_term + r;
        pc[RL_ROW_TERMINATION * N] = c_term + r;
      }
      b.rew_buf[e] = rew;
      pc[(RL_ROW_EXTRAS + 0) * N] = cx[0] + blv.x;               // lin_vel_raw
      pc[(RL_ROW_EXTRAS + 1) * N] = cx[1] + bav.z;               // ang_vel_raw
      pc[(RL_ROW_EXTRAS + 2) * N] = cx[2] + sq(blv.x - cmd.x);   // lin_vel_residual
      pc[(RL_ROW_EXTRAS + 3) * N] = cx[3] + sq(bav.z - cmd.z);   // ang_vel_residual
      pc[(RL_ROW_EXTRAS + 4) * N] = cx[4] + 1.0f;                // ep_timesteps
    }

    // ---- remaining state (:152, :160-162, :183) ----------------------------------------------------
    b.episode_length_buf[e] = ep_len;
    b.base_lin_vel[0 * N + e] = blv.x; b.base_lin_vel[1 * N + e] = blv.y; b.base_lin_vel[2 * N + e] = blv.z;
    b.base_ang_vel[0 * N + e] = bav.x; b.base_ang_vel[1 * N + e] = bav.y; b.base_ang_vel[2 * N + e] = bav.z;
    b.projected_gravity[0 * N + e] = grav.x; b.projected_gravity[1 * N + e] = grav.y;
    b.projected_gravity[2 * N + e] = grav.z;
    b.last_root_vel[0 * N + e] = vw.x; b.last_root_vel[1 * N + e] = vw.y; b.last_root_vel[2 * N + e] = vw.z;
    b.last_root_vel[3 * N + e] = ww.x; b.last_root_vel[4 * N + e] = ww.y; b.last_root_vel[5 * N + e] = ww.z;

    // ---- observations (:342-392) with noise (:392) and clip (:134) --------------------------------
    if (STD_OBS) {
      ObsNoise noise(b.noise_u ? b.noise_u + (size_t)e * cfg.num_obs : nullptr, args.seed, rng_step, (uint32_t)e);
      // gravity 0-2 | commands 3-5 | q 6-17 | qd 18-29 | actions 30-41; noisy values are numbered
      // 0-2 (gravity), 3-14 (q), 15-26 (qd): 4 Philox blocks of 8 sixteen-bit uniforms
      o[0] = grav.x; o[1] = grav.y; o[2] = grav.z;
      o[3] = cmd.x * cfg.commands_scale[0]; o[4] = cmd.y * cfg.commands_scale[1]; o[5] = cmd.z * cfg.commands_scale[2];
#pragma unroll
      for (int j = 0; j < ND; ++j) {
        o[6 + j] = (dof[2 * j] - cfg.default_dof_pos[j]) * cfg.obs_scale_dof_pos;
        o[18 + j] = dof[2 * j + 1] * cfg.obs_scale_dof_vel;
        o[30 + j] = act[j];
      }
      if (cfg.add_noise) {
#pragma unroll
        for (int k = 0; k < 27; ++k) {
          const int c = k < 3 ? k : k + 3;     // noisy value k lives in column c
          if ((k & 7) == 0) noise.need(k);
          o[c] = noise.apply(o[c], cfg.noise_scale_core[c], k, c);
        }
      }
#pragma unroll
      for (int c = 0; c < 42; ++c) o[c] = clampf(o[c], -co, co);
    }
  }
  __syncthreads();   // every thread is done with the input rows: the output rows may overwrite them

  if (valid) {
    float* obs = s_obs + tid * W;
    float* priv = s_priv + tid * RL_PRIV_DIM;
    if (STD_OBS) {
      // 42-float rows: 64-bit shared stores are bank-conflict free
#pragma unroll
      for (int c = 0; c < 42; c += 2) *reinterpret_cast<float2*>(obs + c) = make_float2(o[c], o[c + 1]);
    } else {
      // generic layout: [only_lin][only_ang][lin,ang][gravity][cmd][q][qd][a][yaw], runtime offsets
      ObsNoise noise(b.noise_u ? b.noise_u + (size_t)e * cfg.num_obs : nullptr, args.seed, rng_step, (uint32_t)e);
      const float4 cmd4 = *reinterpret_cast<const float4*>(b.commands + (size_t)e * 4);
      int c = 0, k = 0;
      auto emit = [&](float v) {
        const float ns = cfg.add_noise ? cfg.noise_scale_core[c] : 0.f;
        if (ns != 0.f) { noise.need(k); v = noise.apply(v, ns, k, c); ++k; }
        obs[c++] = clampf(v, -co, co);
      };
      if (cfg.observe_only_lin_vel) {
        emit(blv.x * cfg.obs_scale_lin_vel); emit(blv.y * cfg.obs_scale_lin_vel); emit(blv.z * cfg.obs_scale_lin_vel);
      }
      if (cfg.observe_only_ang_vel) {
        emit(bav.x * cfg.obs_scale_ang_vel); emit(bav.y * cfg.obs_scale_ang_vel); emit(bav.z * cfg.obs_scale_ang_vel);
      }
      if (cfg.observe_vel) {
        const V3 lv = cfg.global_reference ? vw : blv;
        emit(lv.x * cfg.obs_scale_lin_vel); emit(lv.y * cfg.obs_scale_lin_vel); emit(lv.z * cfg.obs_scale_lin_vel);
        emit(bav.x * cfg.obs_scale_ang_vel); emit(bav.y * cfg.obs_scale_ang_vel); emit(bav.z * cfg.obs_scale_ang_vel);
      }
      emit(grav.x); emit(grav.y); emit(grav.z);
      if (cfg.observe_command) {
        emit(cmd4.x * cfg.commands_scale[0]); emit(cmd4.y * cfg.commands_scale[1]); emit(cmd4.z * cfg.commands_scale[2]);
      }
#pragma unroll
      for (int j = 0; j < ND; ++j) emit((dof[2 * j] - cfg.default_dof_pos[j]) * cfg.obs_scale_dof_pos);
#pragma unroll
      for (int j = 0; j < ND; ++j) emit(dof[2 * j + 1] * cfg.obs_scale_dof_vel);
#pragma unroll
      for (int j = 0; j < ND; ++j) emit(act[j]);
      if (cfg.observe_yaw) {
        const V3 fwd = quat_apply(qx, qy, qz, qw, V3{1.f, 0.f, 0.f});
        float heading = atan2f(fwd.y, fwd.x);
        const float two_pi = 6.283185307179586f;
        heading = py_mod(heading, two_pi);
        heading -= two_pi * (heading > 3.141592653589793f ? 1.f : 0.f);
        emit(clampf(0.5f * heading, -1.f, 1.f));
      }
    }
#pragma unroll
    for (int k = 0; k < RL_PRIV_DIM / 2; ++k) *reinterpret_cast<float2*>(priv + 2 * k) = make_float2(pv[2 * k], pv[2 * k + 1]);
    if (FUSE_TORQUES) {
#pragma unroll
      for (int k = 0; k < 3; ++k) *reinterpret_cast<float4*>(s_tq + tid * ND + 4 * k) = *reinterpret_cast<float4*>(tq + 4 * k);
    }
  }
  if (dirty) s_root_dirty = 1;
  fence_async_smem();   // generic-proxy smem writes -> visible to the bulk-copy (async) proxy
  __syncthreads();

  // ---- 6. write-back of the AoS outputs: bulk stores for full aligned tiles ------------------------
  float* o_obs = b.obs_buf + (size_t)tile0 * cfg.num_obs;
  float* o_priv = b.privileged_obs_buf + (size_t)tile0 * RL_PRIV_DIM;
  float* o_tq = b.torques + (size_t)tile0 * ND;
  float* o_root = b.root_states + (size_t)tile0 * 13;
  const bool bulk_out = (n_valid == TILE) && (W == cfg.num_obs) && ((W & 3) == 0) &&
      ((((uintptr_t)o_obs | (uintptr_t)o_priv | (uintptr_t)o_tq | (uintptr_t)o_root) & 15) == 0);
  if (bulk_out) {
    if (tid == 0) {
      bulk_s2g(o_obs, s_obs, (uint32_t)(TILE * W * 4));
      bulk_s2g(o_priv, s_priv, TILE * RL_PRIV_DIM * 4);
      if (FUSE_TORQUES) bulk_s2g(o_tq, s_tq, TILE * ND * 4);
      if (s_root_dirty) bulk_s2g(o_root, s_root, TILE * 13 * 4);
      bulk_commit();
      bulk_wait_read0();   // shared memory must stay alive until the copy engine has read it
    }
  } else {
    if (W == cfg.num_obs) stage_out<TILE>(o_obs, s_obs, n_valid * W);
    else stage_out_rows<TILE>(o_obs, s_obs, n_valid, W, cfg.num_obs);
    stage_out<TILE>(o_priv, s_priv, n_valid * RL_PRIV_DIM);
    if (FUSE_TORQUES) stage_out<TILE>(o_tq, s_tq, n_valid * ND);
    if (s_root_dirty) stage_out<TILE>(o_root, s_root, n_valid * 13);
  }

  // device step counter: every CTA read it on entry; the last one to leave advances it
  if (b.step_state && tid == 0) {
    __threadfence();
    const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(b.step_state + 1), 1ull);
    if (done == gridDim.x - 1) {
      b.step_state[1] = 0;
      atomicAdd(reinterpret_cast<unsigned long long*>(b.step_state), 1ull);
    }
  }
}

// standalone PD torques (:653-688) for the real-simulator path: called `decimation` times
__global__ void __launch_bounds__(128)
env_torques_kernel(const __grid_constant__ StepArgs args) {
  // one thread per (env, leg): 4x the threads of a thread-per-env mapping and a quarter of the dependent work
  // each - this launch runs `decimation` times per step between physics sub-steps, its latency is what counts
  const RlEnvCfg& cfg = args.cfg;
  const RlEnvBuffers& b = args.b;
  const int N = cfg.num_envs;
  const size_t Ns = (size_t)N;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int leg = blockIdx.y;
  if (e >= N) return;
  const float* ap = b.actions_in + (size_t)e * ND + 3 * leg;
  const float2* dp = reinterpret_cast<const float2*>(b.dof_state + (size_t)e * 24 + 6 * leg);     // 8 B aligned
  float act[3], q[3], qd[3], kp[3], kd[3], ms[3], ldv[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int j = 3 * leg + k;
    act[k] = ap[k];
    const float2 d = dp[k];
    q[k] = d.x; qd[k] = d.y;
    kp[k] = b.Kp_factors[j * Ns + e]; kd[k] = b.Kd_factors[j * Ns + e]; ms[k] = b.motor_strengths[j * Ns + e];
    ldv[k] = cfg.control_type == 1 ? b.last_dof_vel[j * Ns + e] : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int j = 3 * leg + k;
    const float a = clampf(act[k], -cfg.clip_actions, cfg.clip_actions);
    float as = a * cfg.action_scale;
    if (k == 0) as *= cfg.hip_scale_reduction;
    float t;
    if (cfg.control_type == 0) {
      const float jpt = as + cfg.default_dof_pos[j];
      b.joint_pos_target[j * Ns + e] = jpt;
      t = cfg.p_gains[j] * kp[k] * (jpt - q[k]) - cfg.d_gains[j] * kd[k] * qd[k];
    } else if (cfg.control_type == 1) {
      t = cfg.p_gains[j] * (as - qd[k]) - cfg.d_gains[j] * (qd[k] - ldv[k]) / cfg.sim_dt;
    } else {
      t = as;
    }
    t = t * ms[k];
    b.torques[(size_t)e * ND + j] = clampf(t, -cfg.torque_limits[j], cfg.torque_limits[j]);
  }
}

static size_t step_smem_bytes(const RlEnvCfg& cfg, int tile) {
  const int P = cfg.measure_heights ? cfg.num_height_points : 0;
  const int W = cfg.num_obs - P;
  const size_t in_floats = 24 + cfg.num_bodies * 3 + ND + ND;       // dof, contact, actions, torques in
  const size_t out_floats = (size_t)W + RL_PRIV_DIM + ND;            // obs, priv, torques out (alias the inputs)
  const size_t floats = (size_t)tile * (1 + 13 + (in_floats > out_floats ? in_floats : out_floats));
  return floats * sizeof(float);
}

static int validate(const RlEnvCfg* cfg, const RlEnvBuffers* b, bool need_torques_in) {
  RL_REQUIRE(cfg && b, RL_ERR_BAD_ARG, "env step: null cfg/buffers");
  RL_REQUIRE(cfg->num_envs > 0, RL_ERR_BAD_CFG, "env step: num_envs=%d", cfg->num_envs);
  RL_REQUIRE(cfg->num_actions == ND, RL_ERR_UNSUPPORTED, "env step: num_actions=%d (only 12 supported)", cfg->num_actions);
  RL_REQUIRE(cfg->num_bodies > 0 && cfg->num_bodies <= RL_MAX_BODIES, RL_ERR_BAD_CFG, "env step: num_bodies=%d", cfg->num_bodies);
  RL_REQUIRE(cfg->n_terms >= 0 && cfg->n_terms <= RL_MAX_TERMS, RL_ERR_UNSUPPORTED, "env step: n_terms=%d (max %d enabled reward terms)", cfg->n_terms, RL_MAX_TERMS);
  RL_REQUIRE((long long)cfg->num_envs * RL_COMMAND_ROWS < (1ll << 31), RL_ERR_UNSUPPORTED,
             "env step: num_envs=%d too large for 32-bit element indices", cfg->num_envs);
  RL_REQUIRE(cfg->num_obs - (cfg->measure_heights ? cfg->num_height_points : 0) <= RL_MAX_CORE_OBS, RL_ERR_UNSUPPORTED,
             "env step: more than %d non-height observation columns", RL_MAX_CORE_OBS);
  RL_REQUIRE(cfg->control_type >= 0 && cfg->control_type <= 2, RL_ERR_BAD_CFG, "env step: control_type=%d", cfg->control_type);
  RL_REQUIRE(cfg->rand_interval > 0 && (!cfg->push_robots || cfg->push_interval > 0), RL_ERR_BAD_CFG, "env step: intervals must be positive");
  RL_REQUIRE(cfg->num_train_envs >= 0 && cfg->num_train_envs <= cfg->num_envs, RL_ERR_BAD_CFG, "env step: num_train_envs=%d of %d envs",
             cfg->num_train_envs, cfg->num_envs);
  RL_REQUIRE(!has_eval_split(*cfg) || !cfg->eval_push_robots || cfg->eval_push_interval > 0, RL_ERR_BAD_CFG,
             "env step: the evaluation push interval must be positive");
  const int P = cfg->measure_heights ? cfg->num_height_points : 0;
  RL_REQUIRE(cfg->num_obs - P > 0, RL_ERR_BAD_CFG, "env step: num_obs=%d <= height points %d", cfg->num_obs, P);
  RL_REQUIRE(b->root_states && b->dof_state && b->contact_forces && b->actions_in && b->torques &&
             b->obs_buf && b->privileged_obs_buf && b->rew_buf && b->reset_buf && b->last_actions &&
             b->last_dof_vel && b->last_root_vel && b->joint_pos_target && b->base_lin_vel &&
             b->base_ang_vel && b->projected_gravity && b->Kp_factors && b->Kd_factors &&
             b->motor_strengths && b->friction_coeffs && b->restitutions && b->payloads &&
             b->com_displacements && b->feet_air_time && b->last_contacts && b->episode_length_buf &&
             b->commands && b->episode_sums && b->command_sums,
             RL_ERR_BAD_ARG, "env step: a required buffer pointer is null");
  RL_REQUIRE(((uintptr_t)b->commands & 15) == 0 && ((uintptr_t)b->last_contacts & 3) == 0, RL_ERR_BAD_ARG,
             "env step: commands must be 16B aligned, last_contacts 4B aligned");
  if (cfg->measure_heights) {
    RL_REQUIRE(b->measured_heights && b->height_points, RL_ERR_BAD_ARG, "env step: heights enabled without buffers");
    RL_REQUIRE(cfg->heights_plane || b->height_samples, RL_ERR_BAD_ARG, "env step: height_samples missing");
    RL_REQUIRE(cfg->heights_plane || (cfg->hf_rows >= 2 && cfg->hf_cols >= 2 && (long long)cfg->hf_rows * cfg->hf_cols < (1ll << 31)),
               RL_ERR_BAD_CFG, "env step: height table %d x %d (needs >= 2 x 2 and < 2^31 cells)", cfg->hf_rows, cfg->hf_cols);
  }
  if (cfg->timeout_resets) RL_REQUIRE(b->time_out_buf, RL_ERR_BAD_ARG, "env step: time_out_buf missing");
  (void)need_torques_in;
  return RL_OK;
}

template <bool FUSE, int TILE, bool STD>
static int launch_inst(const StepArgs& args, size_t smem, cudaStream_t st) {
  static size_t configured = 0;  // per instantiation
  if (smem > configured) {
    cudaError_t err = cudaFuncSetAttribute(env_step_kernel<FUSE, TILE, STD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    configured = smem;
  }
  const int grid = (args.cfg.num_envs + TILE - 1) / TILE;
  env_step_kernel<FUSE, TILE, STD><<<grid, TILE, smem, st>>>(args);
  return check_launch("env_step_kernel");
}

template <bool FUSE, int TILE>
static int launch_tile(const StepArgs& args, size_t smem, cudaStream_t st) {
  const RlEnvCfg& c = args.cfg;
  const int core = c.num_obs - (c.measure_heights ? c.num_height_points : 0);
  const bool std_obs = c.observe_command && !c.observe_vel && !c.observe_only_ang_vel && !c.observe_only_lin_vel &&
                       !c.observe_yaw && core == 42;
  return std_obs ? launch_inst<FUSE, TILE, true>(args, smem, st) : launch_inst<FUSE, TILE, false>(args, smem, st);
}

static int env_tile() {
  static int tile = 0;
  if (!tile) {
    const char* e = getenv("RL_ENV_TILE");
    tile = (e && atoi(e) == 128) ? 128 : 64;
  }
  return tile;
}

template <bool FUSE>
static int launch_step(const RlEnvCfg* cfg, const RlEnvBuffers* b, uint64_t seed, uint64_t step, void* stream) {
  int rc = validate(cfg, b, !FUSE);
  if (rc != RL_OK) return rc;
  StepArgs qargs;
  qargs.cfg = *cfg; qargs.b = *b; qargs.seed = seed; qargs.step = step;
  if (cfg->measure_heights && b->height_mean) {
    // terrain heights first, one warp per env over the whole GPU (heights.cu); the step kernel reads the means
    rc = launch_heights_prepass(qargs, (cudaStream_t)stream);
    if (rc != RL_OK) return rc;
  }
  {
    // standard observation layout -> one-warp-per-leg kernel (env_step_quad.cu); RL_ENV_MODE=thread forces
    // the one-thread-per-env kernel below (also the path of every other observe_* combination)
    static int force_thread = -1;
    if (force_thread < 0) { const char* m = getenv("RL_ENV_MODE"); force_thread = (m && m[0] == 't') ? 1 : 0; }
    const RlEnvCfg& c = *cfg;
    const int core = c.num_obs - (c.measure_heights ? c.num_height_points : 0);
    const bool std_obs = c.observe_command && !c.observe_vel && !c.observe_only_ang_vel && !c.observe_only_lin_vel &&
                         !c.observe_yaw && core == 42;
    // (a train / eval split picks per-env variants of the teleport / push / re-draw fields: the one-thread-per-env kernel)
    if (std_obs && !force_thread && !has_eval_split(c)) return launch_step_quad(qargs, FUSE, (cudaStream_t)stream);
  }
  const int tile = env_tile();
  const size_t smem = step_smem_bytes(*cfg, tile);
  RL_REQUIRE(smem <= 227 * 1024, RL_ERR_UNSUPPORTED, "env step: tile needs %zu B of shared memory", smem);
  StepArgs args;
  args.cfg = *cfg; args.b = *b; args.seed = seed; args.step = step;
  if (tile == 128) return launch_tile<FUSE, 128>(args, smem, (cudaStream_t)stream);
  return launch_tile<FUSE, 64>(args, smem, (cudaStream_t)stream);
}

}  // namespace rl

using namespace rl;

extern "C" int rl_env_step_fused(const RlEnvCfg* cfg, const RlEnvBuffers* b, uint64_t seed, uint64_t step, void* stream) {
  return launch_step<true>(cfg, b, seed, step, stream);
}

extern "C" int rl_env_post_physics(const RlEnvCfg* cfg, const RlEnvBuffers* b, uint64_t seed, uint64_t step, void* stream) {
  return launch_step<false>(cfg, b, seed, step, stream);
}

extern "C" int rl_env_torques(const RlEnvCfg* cfg, const RlEnvBuffers* b, void* stream) {
  RL_REQUIRE(cfg && b, RL_ERR_BAD_ARG, "rl_env_torques: null cfg/buffers");
  RL_REQUIRE(cfg->num_actions == ND, RL_ERR_UNSUPPORTED, "rl_env_torques: num_actions=%d", cfg->num_actions);
  RL_REQUIRE(b->actions_in && b->dof_state && b->torques && b->joint_pos_target && b->Kp_factors &&
             b->Kd_factors && b->motor_strengths && b->last_dof_vel, RL_ERR_BAD_ARG, "rl_env_torques: null buffer");
  RL_REQUIRE((((uintptr_t)b->actions_in | (uintptr_t)b->dof_state | (uintptr_t)b->torques) & 15) == 0,
             RL_ERR_BAD_ARG, "rl_env_torques: actions/dof_state/torques must be 16B aligned");
  StepArgs args;
  args.cfg = *cfg; args.b = *b; args.seed = 0; args.step = 0;
  const int grid = (cfg->num_envs + 127) / 128;
  env_torques_kernel<<<dim3(grid, 4), 128, 0, (cudaStream_t)stream>>>(args);
  return check_launch("env_torques_kernel");
}
