// Fused env step, shipped training configuration, ALL tile traffic on the async (TMA) proxy.
//
// Same arithmetic and warp roles as env_step_quad.cu (reference: mini_gym/envs/base/legged_robot.py
// :106-417, :653-688, :1506-1646) - what changes is how the bytes move.  The quad kernel stages only
// the simulator's AoS rows with cp.async.bulk and reads / writes the env-owned SoA state with ~60
// per-thread LDG / STG (each with 64-bit address arithmetic): ncu showed every CTA of the single wave
// spending 4.6 us issuing loads before the first one could start computing, then all of them computing
// with DRAM idle.  Here the env-owned state lives in three packed row blocks
//     RO [42][N]  Kp, Kd, motor strength factors (12 each), friction, restitution, payload, com (3)
//     RW [28][N]  last_actions (12), last_dof_vel (12), feet_air_time (4)
//     WO [27][N]  joint_pos_target (12), base_lin_vel, base_ang_vel, projected_gravity (3 each), last_root_vel (6)
// (the Python attributes are views of those blocks) and a CTA fetches its 32-env column of every block
// with ONE 2-D tensor copy each (box = rows x 32 envs, 128 B inner extent), plus the accumulator rows of
// episode_sums / command_sums it touches and the four AoS row spans: 11 copies issued by one thread in
// the first ~100 cycles of the CTA, all landing on one mbarrier.  Per SM the TMA unit serves its CTAs in
// issue order, so the first CTA computes while the rows of the later ones are still in flight, and the
// kernel body reads shared memory with immediate offsets (no address arithmetic, no LDG latency).  The
// results leave the same way: 5 tensor stores + 3 bulk stores issued by one thread.
// Measured per-CTA timeline at 32768 envs (profiles/r02_env_rows_trace.md): copies issued 0.9 us after entry, tiles land
// together 4.8 us later (25.6 MB at 5.3 TB/s - all 1024 CTAs are resident at once, so every tile is requested at t = 0),
// then 5 us of compute and stores.  Two things were tried on top and dropped (A/B in one gpurun call): programmatic
// dependent launch (4000 envs: 5.9 -> 8.5 us per launch, 32768: 13.2 -> 13.1) and a staggered start in which a CTA of
// group g issues its copies only when a CTA of group g - 1 has its tile (2 / 3 / 4 groups: 13.7 -> 15.4 / 18.8 / 22.6 us:
// a group's first byte costs ~1.1 us of latency whatever its size, so serialised groups lose more than overlap wins).
// Used when: standard configuration (see launch_step_quad's is_std), num_envs % 32 == 0, the state
// pointers form the packed blocks; otherwise the caller falls back to env_step_quad.cu.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <tuple>

#include "env_common.cuh"

namespace rl {

int make_tmap_f32_rows(CUtensorMap* out, const void* base, uint64_t rows, uint64_t n, uint32_t box_rows);
extern int g_rows_persist_mode;    // rl_debug_env_rows (env_step_quad.cu)

static bool rows_trace_on = false;
namespace rows {

constexpr int QT = 32, QTHREADS = 128;
enum RoRow { RO_KP = 0, RO_KD = 12, RO_MS = 24, RO_FRIC = 36, RO_REST = 37, RO_PAYLOAD = 38, RO_COM = 39, RO_ROWS = 42 };
enum RwRow { RW_LA = 0, RW_LDV = 12, RW_AIR = 24, RW_ROWS = 28 };
enum WoRow { WO_JPT = 0, WO_BLV = 12, WO_BAV = 15, WO_GRAV = 18, WO_LRV = 21, WO_ROWS = 27 };
constexpr int ES_ROWS = 13;       // term rows 0-11 | total
constexpr int CS_ROWS = 17;       // term rows 0-11 | lin_vel_raw, ang_vel_raw, lin_vel_residual, ang_vel_residual, ep_timesteps
enum Part { P_TQ2 = 0, P_ACC2, P_RATE2, P_LIM, NPART };
constexpr int ROWB = QT * 4;      // bytes of one 32-env row

// Everything the copy-issuing thread reads before its first request sits in ONE 64-byte block at the start of the kernel
// parameters: a parameter read that misses the constant cache costs a few hundred cycles, and the issuing thread used to
// touch five different lines of the 3.5 KB argument block (cfg, buffers, stagger fields) one after the other before the
// first byte was requested (entry -> copies issued: 0.51 us at 4000 envs).
struct IssueHdr {
  const float* dof_state; const float* actions; const float* root_states; const float* contact_forces; const float* torques_in;
  int num_bodies;
  int stagger_ns, stagger_from, stagger_group;  // CTAs with blockIdx >= stagger_from issue their copies stagger_ns later (0: off);
                                                // stagger_group > 0: another stagger_ns for every further `stagger_group` CTAs
  int pad[2];
};
static_assert(sizeof(IssueHdr) == 64, "IssueHdr is one 64-byte block");
struct RowsArgs {
  IssueHdr h;
  CUtensorMap m_ro, m_rw, m_wo, m_es12, m_es1, m_cs12, m_cs5;
  StepArgs a;
  unsigned long long* trace;     // profiling aid (rl_debug_env_rows_trace): 8 globaltimer stamps per CTA, or null
};
constexpr int TRACE_STAMPS = 8;
template <bool TRACE>
__device__ __forceinline__ void stamp(unsigned long long* trace, int i) {
  if (TRACE && trace && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[(size_t)blockIdx.x * TRACE_STAMPS + i] = t;
  }
}

__device__ __forceinline__ void tma_load_rows(void* smem_dst, const CUtensorMap* map, int env0, int row0, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(env0), "r"(row0), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_rows(const CUtensorMap* map, const void* smem_src, int env0, int row0) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(env0), "r"(row0) : "memory");
}

// shared-memory map (bytes from the 128 B aligned base); NB = bodies
struct Lay {
  int ro, rw, es, cs, wo, root, dof, con, act, tq_in, x, r, total;
  __host__ __device__ Lay(int NB, bool fuse) {
    ro = 0;
    rw = ro + RO_ROWS * ROWB;
    es = rw + RW_ROWS * ROWB;
    cs = es + ES_ROWS * ROWB;
    wo = cs + CS_ROWS * ROWB;
    root = wo + WO_ROWS * ROWB;
    dof = root + QT * 13 * 4;
    con = dof + QT * 24 * 4;
    act = con + QT * NB * 3 * 4;
    tq_in = act + QT * ND * 4;
    const int in_end = tq_in + (fuse ? 0 : QT * ND * 4);
    const int out_end = dof + QT * (42 + RL_PRIV_DIM + ND) * 4;
    x = ((in_end > out_end ? in_end : out_end) + 127) & ~127;
    total = x + (NPART * 4 + 2) * ROWB;           // partial sums | collision, air-time reward
    // r_i [12][32] (written in phase 2, read in phase 3) re-uses the action rows - dead after phase 1, exactly 12 x 32
    // floats - when the output rows that re-use the input tile end below them (17 bodies: 7 CTAs per SM fit only so)
    if (act >= out_end) r = act;
    else { r = total; total += 12 * ROWB; }
  }
};

// shared-memory pointers of one tile buffer + the per-thread constants every part of the step uses
#define RL_ROWS_TILE_SETUP(BUF)                                                                          \
  const RlEnvCfg& cfg = args.a.cfg;                                                                      \
  const RlEnvBuffers& b = args.a.b;                                                                      \
  const int N = cfg.num_envs, NB = cfg.num_bodies;                                                       \
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;                                            \
  const int e = tile0 + lane;                                                                            \
  constexpr int W = 42;                                                                                  \
  const float co = cfg.clip_obs;                                                                         \
  uint8_t* const smem_raw = (BUF);                                                                       \
  const Lay L(NB, FUSE);                                                                                 \
  float* s_ro = reinterpret_cast<float*>(smem_raw + L.ro);                                               \
  float* s_rw = reinterpret_cast<float*>(smem_raw + L.rw);                                               \
  float* s_es = reinterpret_cast<float*>(smem_raw + L.es);                                               \
  float* s_cs = reinterpret_cast<float*>(smem_raw + L.cs);                                               \
  float* s_wo = reinterpret_cast<float*>(smem_raw + L.wo);                                               \
  float* s_root = reinterpret_cast<float*>(smem_raw + L.root);                                           \
  float* s_dof = reinterpret_cast<float*>(smem_raw + L.dof);                                             \
  float* s_con = reinterpret_cast<float*>(smem_raw + L.con);                                             \
  float* s_act = reinterpret_cast<float*>(smem_raw + L.act);                                             \
  float* s_tq_in = reinterpret_cast<float*>(smem_raw + L.tq_in);                                         \
  float* s_obs = s_dof;                                   /* output rows re-use the input tile */        \
  float* s_priv = s_obs + QT * W;                                                                        \
  float* s_tq = s_priv + QT * RL_PRIV_DIM;                                                               \
  float* s_part = reinterpret_cast<float*>(smem_raw + L.x);        /* [NPART][4][32] */                  \
  float* s_coll = s_part + NPART * 4 * QT;                /* [32] */                                     \
  float* s_air = s_coll + QT;                             /* [32] */                                     \
  float* s_r = reinterpret_cast<float*>(smem_raw + L.r);    /* [12][32] per-term rewards */              \
  (void)N; (void)e; (void)co; (void)s_ro; (void)s_rw; (void)s_es; (void)s_cs; (void)s_wo; (void)s_root; \
  (void)s_con; (void)s_act; (void)s_tq_in; (void)s_obs; (void)s_priv; (void)s_tq; (void)s_part;          \
  (void)s_coll; (void)s_air; (void)s_r; (void)w; (void)lane; (void)b; (void)tid

// ---- one thread issues every copy of a tile: 11 requests landing on `bar` ------------------------------------
template <bool FUSE>
__device__ __forceinline__ void issue_tile_loads(const RowsArgs& args, uint8_t* buf, uint64_t* bar, int tile0) {
  const IssueHdr& h = args.h;
  const int NB = h.num_bodies;
  const Lay L(NB, FUSE);
  // two landing groups: bar[0] = what the per-leg work of phase 1 reads (DOF state, actions, RO / RW blocks), bar[1] = the
  // rest (root / contact rows for the per-env roles, the accumulator rows of phase 2).  The TMA unit serves a CTA's requests
  // in issue order, so the leg work starts while the second half of the tile is still in flight.
  uint64_t& bar_a = bar[0];
  uint64_t& bar_b = bar[1];
  mbar_expect_tx(&bar_a, (uint32_t)((RO_ROWS + RW_ROWS) * ROWB + QT * (24 + ND + (FUSE ? 0 : ND)) * 4));
  mbar_expect_tx(&bar_b, (uint32_t)((ES_ROWS + CS_ROWS) * ROWB + QT * (13 + NB * 3) * 4));
  bulk_g2s(buf + L.dof, h.dof_state + (size_t)tile0 * 24, QT * 24 * 4, &bar_a);
  bulk_g2s(buf + L.act, h.actions + (size_t)tile0 * ND, QT * ND * 4, &bar_a);
  tma_load_rows(buf + L.ro, &args.m_ro, tile0, 0, &bar_a);
  tma_load_rows(buf + L.rw, &args.m_rw, tile0, 0, &bar_a);
  if (!FUSE) bulk_g2s(buf + L.tq_in, h.torques_in + (size_t)tile0 * ND, QT * ND * 4, &bar_a);
  bulk_g2s(buf + L.root, h.root_states + (size_t)tile0 * 13, QT * 13 * 4, &bar_b);
  bulk_g2s(buf + L.con, h.contact_forces + (size_t)tile0 * NB * 3, (uint32_t)(QT * NB * 3 * 4), &bar_b);
  tma_load_rows(buf + L.es, &args.m_es12, tile0, 0, &bar_b);
  tma_load_rows(buf + L.es + 12 * ROWB, &args.m_es1, tile0, RL_ROW_TOTAL, &bar_b);
  tma_load_rows(buf + L.cs, &args.m_cs12, tile0, 0, &bar_b);
  tma_load_rows(buf + L.cs + 12 * ROWB, &args.m_cs5, tile0, RL_ROW_EXTRAS, &bar_b);
}

// ---- the step of one tile: waits for `bar` (phase `parity`), phases 1 - 3, stores issued (one bulk group) ----
template <bool FUSE, bool TRACE, bool BUMP_STEP>
__device__ __forceinline__ void tile_body(const RowsArgs& args, uint8_t* buf, uint64_t* bar, uint32_t parity, int tile0,
                                          int* s_root_dirty_p, unsigned long long* trace_p) {
  RL_ROWS_TILE_SETUP(buf);
  int& s_root_dirty = *s_root_dirty_p;
  struct { unsigned long long* trace; } targs{trace_p};
  // small per-env scalars with their own dtypes: plain coalesced loads
  const uint64_t rng_step = args.a.step + (b.step_state ? b.step_state[0] : 0ull);
  int ep = (int)b.episode_length_buf[e];
  const float4 cmd = *reinterpret_cast<const float4*>(b.commands + (size_t)e * 4);
  uint32_t last_contacts = 0;
  if (w == 1) last_contacts = *reinterpret_cast<const uint32_t*>(b.last_contacts + (size_t)e * 4);
  // ---- work that does not depend on the tile runs while it is in flight: the Philox blocks of this warp's noise columns
  // (a ~90-instruction dependent chain each) and the DOF-property re-draw test (an integer remainder) ----
  const float* const nu_row = b.noise_u ? b.noise_u + (size_t)e * cfg.num_obs : nullptr;
  uint32_t rq4[4] = {0u, 0u, 0u, 0u}, rg4[4] = {0u, 0u, 0u, 0u};
  if (!nu_row) {
    Philox::gen(args.a.seed, (uint32_t)e, (uint32_t)rng_step, (uint32_t)(rng_step >> 32), (RNG_NOISE << 16) | (uint32_t)w, rq4);
    if (w == 0) Philox::gen(args.a.seed, (uint32_t)e, (uint32_t)rng_step, (uint32_t)(rng_step >> 32), (RNG_NOISE << 16) | 4u, rg4);
  }
  const bool redraw = ((ep + 1) % cfg.rand_interval) == 0 &&
                      (cfg.randomize_motor_strength | cfg.randomize_Kp_factor | cfg.randomize_Kd_factor);
  __syncthreads();                      // mbarriers initialised
  if (BUMP_STEP && tid == 96 && b.step_state) {     // (a lane of warp 3, whose phase 1 is the shortest)
    // Device step counter (CUDA-graph replay): every thread of this CTA has read step_state[0] (above the barrier), so the
    // CTA can be counted NOW - the last CTA to have read the counter advances it for the next launch.  Counting at the end
    // of the kernel put the atomic's round trip (~0.5 us) after the last store of the last CTA, on the launch's tail; here
    // it overlaps with the tile's flight.
    const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(b.step_state + 1), 1ull);
    if (done == gridDim.x - 1) {
      b.step_state[1] = 0;
      atomicAdd(reinterpret_cast<unsigned long long*>(b.step_state), 1ull);
    }
  }
  // a landing group is complete (bounded: a byte-count bug must trap, not hang)
  auto wait_group = [&](int gidx) {
    uint32_t spins = 0, ok = 0;
    const uint32_t bar_addr = smem_u32(bar + gidx);
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
      if (!ok && ++spins > (1u << 22)) {
        if (tid == 0) printf("env_step_rows_kernel: tile %d never landed (group %d)\n", tile0 / QT, gidx);
        __trap();
      }
    }
  };
  wait_group(0);
  stamp<TRACE>(targs.trace, 2);
  ep += 1;                              // :152
  float* root = s_root + lane * 13;
  bool dirty = false;

  // =================================== phase 1 ===========================================================
  float tq[3], oq[3], oqd[3], oa[3], pm[3], jpt[3];
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  float sc6[6];
  {
    // ---- all warps: the three DOFs of leg w (:653-688 and the per-DOF reward sums) ----
    const float2 d0 = *reinterpret_cast<const float2*>(s_dof + lane * 24 + 6 * w);
    const float2 d1 = *reinterpret_cast<const float2*>(s_dof + lane * 24 + 6 * w + 2);
    const float2 d2 = *reinterpret_cast<const float2*>(s_dof + lane * 24 + 6 * w + 4);
    const float q[3] = {d0.x, d1.x, d2.x}, qd[3] = {d0.y, d1.y, d2.y};
    float part[NPART];
#pragma unroll
    for (int t = 0; t < NPART; ++t) part[t] = 0.f;
    float ms[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int j = 3 * w + k;
      const float kp = s_ro[(RO_KP + j) * QT + lane], kd = s_ro[(RO_KD + j) * QT + lane];
      ms[k] = s_ro[(RO_MS + j) * QT + lane];
      const float la = s_rw[(RW_LA + j) * QT + lane], ldv = s_rw[(RW_LDV + j) * QT + lane];
      const float a = clampf(s_act[lane * ND + j], -cfg.clip_actions, cfg.clip_actions);   // :112-113
      float t;
      if (FUSE) {
        float as = a * cfg.action_scale;
        if (k == 0) as *= cfg.hip_scale_reduction;             // dofs 0,3,6,9 (:666)
        jpt[k] = as + cfg.default_dof_pos[j];
        t = cfg.p_gains[j] * kp * (jpt[k] - q[k]) - cfg.d_gains[j] * kd * qd[k];
        t = t * ms[k];
        t = clampf(t, -cfg.torque_limits[j], cfg.torque_limits[j]);
      } else {
        t = s_tq_in[lane * ND + j];
      }
      tq[k] = t;
      part[P_TQ2] += sq(t);
      part[P_ACC2] += sq((ldv - qd[k]) / cfg.dt);
      part[P_RATE2] += sq(la - a);
      {
        float ov = -fminf(q[k] - cfg.dof_pos_lo[j], 0.f);
        ov += fmaxf(q[k] - cfg.dof_pos_hi[j], 0.f);
        part[P_LIM] += ov;
      }
      oq[k] = (q[k] - cfg.default_dof_pos[j]) * cfg.obs_scale_dof_pos;
      oqd[k] = qd[k] * cfg.obs_scale_dof_vel;
      oa[k] = a;
      s_rw[(RW_LA + j) * QT + lane] = a;              // :181-182 (row j, this lane: touched by this thread only)
      s_rw[(RW_LDV + j) * QT + lane] = qd[k];
      if (FUSE) s_wo[(WO_JPT + j) * QT + lane] = jpt[k];
    }
#pragma unroll
    for (int t = 0; t < NPART; ++t) s_part[(t * 4 + w) * QT + lane] = part[t];

    // DOF-property re-draw (:591-593, :544-560): rare - written straight to global memory
    if (redraw) {
      float u3[4];
      if (b.dr_u) { u3[0] = b.dr_u[e]; u3[1] = b.dr_u[N + e]; u3[2] = b.dr_u[2 * N + e]; }
      else rng4(args.a.seed, (uint32_t)e, rng_step, RNG_DR, 0, u3);
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
    {
      const float* nu = nu_row;
      const uint32_t (&r4)[4] = rq4;
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

    // ---- the second landing group: root / contact rows, accumulators ----
    wait_group(1);
    // teleport (:768-791) by warp 0
    if (w == 0 && cfg.teleport_robots) {
      float x = root[0], y = root[1];
      const float x0 = x, y0 = y;
      if (x < cfg.teleport_lo_x) x += cfg.teleport_shift_x;
      if (x > cfg.teleport_hi_x) x -= cfg.teleport_shift_x;
      if (y < cfg.teleport_lo_y) y += cfg.teleport_shift_y;
      if (y > cfg.teleport_hi_y) y -= cfg.teleport_shift_y;
      if (x != x0 || y != y0) { root[0] = x; root[1] = y; dirty = true; }
    }
    if (w == 0) {
      // ---- frames (:159-162) ----
      const float qx = root[3], qy = root[4], qz = root[5], qw = root[6];
      const V3 vw = {root[7], root[8], root[9]};
      const V3 ww = {root[10], root[11], root[12]};
      const V3 blv = quat_rotate_inverse(qx, qy, qz, qw, vw);
      const V3 bav = quat_rotate_inverse(qx, qy, qz, qw, ww);
      const V3 grav = quat_rotate_inverse(qx, qy, qz, qw, V3{0.f, 0.f, -1.f});
      s_wo[(WO_BLV + 0) * QT + lane] = blv.x; s_wo[(WO_BLV + 1) * QT + lane] = blv.y; s_wo[(WO_BLV + 2) * QT + lane] = blv.z;
      s_wo[(WO_BAV + 0) * QT + lane] = bav.x; s_wo[(WO_BAV + 1) * QT + lane] = bav.y; s_wo[(WO_BAV + 2) * QT + lane] = bav.z;
      s_wo[(WO_GRAV + 0) * QT + lane] = grav.x; s_wo[(WO_GRAV + 1) * QT + lane] = grav.y; s_wo[(WO_GRAV + 2) * QT + lane] = grav.z;
      s_wo[(WO_LRV + 0) * QT + lane] = vw.x; s_wo[(WO_LRV + 1) * QT + lane] = vw.y; s_wo[(WO_LRV + 2) * QT + lane] = vw.z;
      s_wo[(WO_LRV + 3) * QT + lane] = ww.x; s_wo[(WO_LRV + 4) * QT + lane] = ww.y; s_wo[(WO_LRV + 5) * QT + lane] = ww.z;
      // gravity observation columns 0-2 with noise: Philox block 4, lanes 0-2
      g0 = grav.x; g1 = grav.y; g2 = grav.z;
      const float* nu = nu_row;
      if (nu) {
        g0 += (2.0f * nu[0] - 1.0f) * cfg.noise_scale_core[0];
        g1 += (2.0f * nu[1] - 1.0f) * cfg.noise_scale_core[1];
        g2 += (2.0f * nu[2] - 1.0f) * cfg.noise_scale_core[2];
      } else {
        const uint32_t (&r4)[4] = rg4;
        g0 = __fmaf_rn(2.0f * centered_u16(r4, 0), cfg.noise_scale_core[0], g0);
        g1 = __fmaf_rn(2.0f * centered_u16(r4, 1), cfg.noise_scale_core[1], g1);
        g2 = __fmaf_rn(2.0f * centered_u16(r4, 2), cfg.noise_scale_core[2], g2);
      }
      g0 = clampf(g0, -co, co); g1 = clampf(g1, -co, co); g2 = clampf(g2, -co, co);
    } else if (w == 1) {
      // ---- contact terms: termination (:190-202), collision, feet air time (:1619-1631) ----
      const float* con = s_con + lane * NB * 3;
      bool reset = false;
#pragma unroll 1
      for (int k = 0; k < cfg.n_term_bodies; ++k) {
        const float* f = con + cfg.term_idx[k] * 3;
        reset |= sqrtf((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]) > 1.0f;
      }
      b.reset_buf[e] = reset ? 1 : 0;
      b.episode_length_buf[e] = (int64_t)ep;
      float coll = 0.f;
#pragma unroll 1
      for (int k = 0; k < cfg.n_pen_bodies; ++k) {
        const float* f = con + cfg.pen_idx[k] * 3;
        coll += (sqrtf((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]) > 0.1f) ? 1.f : 0.f;
      }
      float r_air = 0.f;
      uint32_t nc = 0;
#pragma unroll
      for (int k = 0; k < RL_NUM_FEET; ++k) {
        float air = s_rw[(RW_AIR + k) * QT + lane];
        const bool contact = con[cfg.feet_idx[k] * 3 + 2] > 1.0f;
        const bool filt = contact || ((last_contacts >> (8 * k)) & 0xffu);
        nc |= (contact ? 1u : 0u) << (8 * k);
        const bool first = (air > 0.f) && filt;
        air += cfg.dt;
        r_air += (air - 0.5f) * (first ? 1.f : 0.f);
        air *= filt ? 0.f : 1.f;
        s_rw[(RW_AIR + k) * QT + lane] = air;
      }
      *reinterpret_cast<uint32_t*>(b.last_contacts + (size_t)e * 4) = nc;
      s_coll[lane] = coll;
      s_air[lane] = r_air;
    } else if (w == 2) {
      // ---- privileged-observation scalars (:398-417) + clip (:136) ----
      sc6[0] = clampf((s_ro[RO_FRIC * QT + lane] - cfg.priv_shift[0]) * cfg.priv_scale[0], -co, co);
      sc6[1] = clampf((s_ro[RO_REST * QT + lane] - cfg.priv_shift[1]) * cfg.priv_scale[1], -co, co);
      sc6[2] = clampf((s_ro[RO_PAYLOAD * QT + lane] - cfg.priv_shift[2]) * cfg.priv_scale[2], -co, co);
#pragma unroll
      for (int k = 0; k < 3; ++k) sc6[3 + k] = clampf((s_ro[(RO_COM + k) * QT + lane] - cfg.priv_shift[3]) * cfg.priv_scale[3], -co, co);
    }
  }
  // the RW and WO blocks are final after phase 1 (phase 2 only reads them): they leave now, under phase 2, issued by a lane
  // of warp 3 (the shortest phase 2) as a bulk group of its own
  fence_async_smem();
  __syncthreads();
  if (tid == 96) {
    const int wo0 = FUSE ? 0 : WO_BLV;           // post_physics leaves joint_pos_target alone
    tma_store_rows(&args.m_rw, s_rw, tile0, 0);
    tma_store_rows(&args.m_wo, s_wo + wo0 * QT, tile0, wo0);
    bulk_commit();
  }
  stamp<TRACE>(targs.trace, 3);

  // =================================== phase 2 ===========================================================
  // warp w evaluates terms w, w + 4, w + 8 (reward_names order of the shipped configuration), adds them to its
  // accumulator rows (in shared memory) and publishes r_i; it also writes its slices of the output rows.
  {
    auto psum = [&](int t) {
      const float* p = s_part + t * 4 * QT + lane;
      return (p[0] + p[QT]) + (p[2 * QT] + p[3 * QT]);
    };
    auto BLV = [&](int i) { return s_wo[(WO_BLV + i) * QT + lane]; };
    auto BAV = [&](int i) { return s_wo[(WO_BAV + i) * QT + lane]; };
    auto GRAV = [&](int i) { return s_wo[(WO_GRAV + i) * QT + lane]; };
    float r0, r1, r2;
    if (w == 0) {
      r0 = expf(-(sq(cmd.x - BLV(0)) + sq(cmd.y - BLV(1))) / cfg.tracking_sigma);        // tracking_lin_vel
      r1 = sq(GRAV(0)) + sq(GRAV(1));                                                     // orientation
      const float cmd_xy_norm = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y);
      r2 = s_air[lane] * ((cmd_xy_norm > 0.1f) ? 1.f : 0.f);                              // feet_air_time
    } else if (w == 1) {
      r0 = expf(-sq(cmd.z - BAV(2)) / cfg.tracking_sigma_yaw);                            // tracking_ang_vel
      r1 = psum(P_TQ2);                                                                   // torques
      r2 = s_coll[lane];                                                                  // collision
    } else if (w == 2) {
      r0 = sq(BLV(2));                                                                    // lin_vel_z
      r1 = psum(P_ACC2);                                                                  // dof_acc
      r2 = psum(P_RATE2);                                                                 // action_rate
    } else {
      r0 = sq(BAV(0)) + sq(BAV(1));                                                       // ang_vel_xy
      r1 = sq(root[2] - cfg.base_height_target);                                          // base_height
      r2 = psum(P_LIM);                                                                   // dof_pos_limits
    }
    r0 *= cfg.term_scale[w]; r1 *= cfg.term_scale[w + 4]; r2 *= cfg.term_scale[w + 8];
    s_es[w * QT + lane] += r0; s_cs[w * QT + lane] += r0; s_r[w * QT + lane] = r0;
    s_es[(w + 4) * QT + lane] += r1; s_cs[(w + 4) * QT + lane] += r1; s_r[(w + 4) * QT + lane] = r1;
    s_es[(w + 8) * QT + lane] += r2; s_cs[(w + 8) * QT + lane] += r2; s_r[(w + 8) * QT + lane] = r2;
    if (w == 1) {
      const float bx = BLV(0), wz = BAV(2);
      float* x = s_cs + 12 * QT + lane;
      x[0 * QT] += bx;                      // lin_vel_raw (:336-340)
      x[1 * QT] += wz;                      // ang_vel_raw
      x[2 * QT] += sq(bx - cmd.x);          // lin_vel_residual
      x[3 * QT] += sq(wz - cmd.z);          // ang_vel_residual
      x[4 * QT] += 1.0f;                    // ep_timesteps
    }
    // output rows (re-using the input tile: every thread passed the barrier above)
    float* obs = s_obs + lane * W;
    float* priv = s_priv + lane * RL_PRIV_DIM;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      obs[6 + 3 * w + k] = oq[k];
      obs[18 + 3 * w + k] = oqd[k];
      obs[30 + 3 * w + k] = oa[k];
      priv[6 + 3 * w + k] = pm[k];
      if (FUSE) {
        s_tq[lane * ND + 3 * w + k] = tq[k];
      }
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
  fence_async_smem();
  __syncthreads();
  stamp<TRACE>(targs.trace, 4);

  // =================================== stores + phase 3 ==================================================
  if (tid == 0) {
    tma_store_rows(&args.m_es12, s_es, tile0, 0);
    tma_store_rows(&args.m_cs12, s_cs, tile0, 0);
    tma_store_rows(&args.m_cs5, s_cs + 12 * QT, tile0, RL_ROW_EXTRAS);
    bulk_s2g(b.obs_buf + (size_t)tile0 * W, s_obs, QT * W * 4);
    bulk_s2g(b.privileged_obs_buf + (size_t)tile0 * RL_PRIV_DIM, s_priv, QT * RL_PRIV_DIM * 4);
    if (FUSE) bulk_s2g(b.torques + (size_t)tile0 * ND, s_tq, QT * ND * 4);
    if (s_root_dirty) bulk_s2g(b.root_states + (size_t)tile0 * 13, s_root, QT * 13 * 4);
    bulk_commit();
  }
  if (w == 3) {
    // warp 3 closes compute_reward (:314-340): sum in reward_names order, positive clip, total row
    float rew = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) rew += s_r[i * QT + lane];
    if (b.rew_raw) b.rew_raw[e] = rew;
    rew = fmaxf(rew, 0.f);
    b.episode_sums[RL_ROW_TOTAL * N + e] = s_es[12 * QT + lane] + rew;
    b.rew_buf[e] = rew;
  }
}

template <bool FUSE, int MINB, bool TRACE>
__global__ void __launch_bounds__(QTHREADS, MINB)
env_step_rows_kernel(const __grid_constant__ RowsArgs args) {
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar[2];      // the two landing groups of the tile
  __shared__ int s_root_dirty;
  const int tile0 = blockIdx.x * QT;
  const int tid = threadIdx.x;
  const RlEnvBuffers& b = args.a.b;

  // Programmatic dependent launch: let the NEXT kernel of the stream start launching its CTAs now (they run their
  // prologue and park in griddepcontrol.wait until this grid has completed and flushed), and wait for the PREVIOUS
  // kernel before the first global access.  Hides the launch latency / CTA ramp between consecutive steps; both
  // instructions are no-ops when the launch does not carry the programmatic-serialization attribute.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  stamp<TRACE>(args.trace, 0);
  if (TRACE && args.trace && tid == 0) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    args.trace[(size_t)blockIdx.x * TRACE_STAMPS + 7] = smid;
  }
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
    s_root_dirty = 0;
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (tid == 0) {
    // time-staggered second group: its requests queue up BEHIND the first group's instead of competing with them, so
    // the first group's tiles land early and are computed while the second group's bytes stream in
    if (args.h.stagger_ns > 0 && (int)blockIdx.x >= args.h.stagger_from) {
      const int grp = args.h.stagger_group > 0 ? 1 + ((int)blockIdx.x - args.h.stagger_from) / args.h.stagger_group : 1;
      __nanosleep((unsigned)(args.h.stagger_ns * grp));
    }
    issue_tile_loads<FUSE>(args, smem_dyn, s_bar, tile0);
  }
  stamp<TRACE>(args.trace, 1);
  tile_body<FUSE, TRACE, true>(args, smem_dyn, s_bar, 0u, tile0, &s_root_dirty, args.trace);
  stamp<TRACE>(args.trace, 5);
  if (tid == 96) bulk_wait_read0();    // (the early group of the RW / WO blocks)
  if (tid == 0) {
    bulk_wait_read0();                 // the stores have read their shared-memory source
    stamp<TRACE>(args.trace, 6);
  }
}

// ---- wide variant for grids of at most two CTAs per SM: 16 warps per 32-env tile ---------------------------------------
// With one or two CTAs per SM (4000 envs = 125 CTAs) nothing hides a warp's latency: the four warps of env_step_rows_kernel
// run ~700 dependent instructions each after the tile has landed (2.0 - 2.3 us of the 5.0 us a launch takes, r02_env_step.md).
// Here the same arithmetic is dealt to 16 warps - warp j < 12 owns DOF j (torque, its four reward partials, its observation /
// privileged columns), warps 12 - 15 the per-env roles; in phase 2 warp i < 12 evaluates reward term i - so the dependent chain
// of a warp is about a third as long.  Every value is computed by the same operations in the same order (the per-leg partial
// sums are re-associated exactly as the four-warp kernel forms them: ((0 + x0) + x1) + x2 per leg, (l0 + l1) + (l2 + l3) over
// legs), so the results are bit-identical (tests/test_env_gpu.py::test_rows_kernel_matches_quad_kernel, mode 4).
constexpr int WIDE_THREADS = 512;
template <bool FUSE>
__global__ void __launch_bounds__(WIDE_THREADS, 1)
env_step_rows_wide_kernel(const __grid_constant__ RowsArgs args) {
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ int s_root_dirty;
  const int tile0 = blockIdx.x * QT;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // (one issuing thread: four threads issuing a part of the copies each onto four barriers measured SLOWER - 4000 envs 4.62 vs
  // 4.43 us per launch - as did two issuing threads in the four-warp kernel)
  if (threadIdx.x == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
    s_root_dirty = 0;
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) issue_tile_loads<FUSE>(args, smem_dyn, s_bar, tile0);

  RL_ROWS_TILE_SETUP(smem_dyn);
  float* s_pd = reinterpret_cast<float*>(smem_raw + L.total);          // [NPART][12][32]: per-DOF reward partials
  float* s_gn = s_pd + NPART * ND * QT;                                // [3][32]: noise of the gravity columns
  const uint64_t rng_step = args.a.step + (b.step_state ? b.step_state[0] : 0ull);
  int ep = (int)b.episode_length_buf[e];
  const float4 cmd = *reinterpret_cast<const float4*>(b.commands + (size_t)e * 4);
  uint32_t last_contacts = 0;
  if (w == 13) last_contacts = *reinterpret_cast<const uint32_t*>(b.last_contacts + (size_t)e * 4);
  // ---- work that does not need the tile: Philox blocks (DOF warps: the block of their leg; warp 12: the gravity block),
  // the DOF re-draw test ----
  const float* const nu_row = b.noise_u ? b.noise_u + (size_t)e * cfg.num_obs : nullptr;
  uint32_t r4[4] = {0u, 0u, 0u, 0u};
  if (!nu_row && w <= 12)
    Philox::gen(args.a.seed, (uint32_t)e, (uint32_t)rng_step, (uint32_t)(rng_step >> 32),
                (RNG_NOISE << 16) | (uint32_t)(w < 12 ? w / 3 : 4), r4);
  const bool redraw = w < 12 && ((ep + 1) % cfg.rand_interval) == 0 &&
                      (cfg.randomize_motor_strength | cfg.randomize_Kp_factor | cfg.randomize_Kd_factor);
  __syncthreads();                      // mbarriers initialised
  if (tid == 480 && b.step_state) {     // counted as soon as every thread has read the counter (see tile_body)
    const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(b.step_state + 1), 1ull);
    if (done == gridDim.x - 1) {
      b.step_state[1] = 0;
      atomicAdd(reinterpret_cast<unsigned long long*>(b.step_state), 1ull);
    }
  }
  auto wait_group = [&](int gidx) {
    uint32_t spins = 0, ok = 0;
    const uint32_t bar_addr = smem_u32(s_bar + gidx);
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar_addr), "r"(0u) : "memory");
      if (!ok && ++spins > (1u << 22)) {
        if (tid == 0) printf("env_step_rows_wide_kernel: tile %d never landed (group %d)\n", tile0 / QT, gidx);
        __trap();
      }
    }
  };
  ep += 1;                              // :152
  float* root = s_root + lane * 13;
  bool dirty = false;

  // =================================== phase 1 ===========================================================
  float tq = 0.f, oq = 0.f, oqd = 0.f, oa = 0.f, pm = 0.f;
  float sc6[6];
  if (w < 12) {
    // ---- DOF j (:653-688 and its reward partials) ----
    wait_group(0);
    const int j = w, k = j % 3;
    const float2 d = *reinterpret_cast<const float2*>(s_dof + lane * 24 + 2 * j);
    const float q = d.x, qd = d.y;
    const float kp = s_ro[(RO_KP + j) * QT + lane], kd = s_ro[(RO_KD + j) * QT + lane];
    float ms = s_ro[(RO_MS + j) * QT + lane];
    const float la = s_rw[(RW_LA + j) * QT + lane], ldv = s_rw[(RW_LDV + j) * QT + lane];
    const float a = clampf(s_act[lane * ND + j], -cfg.clip_actions, cfg.clip_actions);   // :112-113
    float t;
    if (FUSE) {
      float as = a * cfg.action_scale;
      if (k == 0) as *= cfg.hip_scale_reduction;             // dofs 0,3,6,9 (:666)
      const float jpt = as + cfg.default_dof_pos[j];
      t = cfg.p_gains[j] * kp * (jpt - q) - cfg.d_gains[j] * kd * qd;
      t = t * ms;
      t = clampf(t, -cfg.torque_limits[j], cfg.torque_limits[j]);
      s_wo[(WO_JPT + j) * QT + lane] = jpt;
    } else {
      t = s_tq_in[lane * ND + j];
    }
    tq = t;
    s_pd[(P_TQ2 * ND + j) * QT + lane] = sq(t);
    s_pd[(P_ACC2 * ND + j) * QT + lane] = sq((ldv - qd) / cfg.dt);
    s_pd[(P_RATE2 * ND + j) * QT + lane] = sq(la - a);
    {
      float ov = -fminf(q - cfg.dof_pos_lo[j], 0.f);
      ov += fmaxf(q - cfg.dof_pos_hi[j], 0.f);
      s_pd[(P_LIM * ND + j) * QT + lane] = ov;
    }
    oq = (q - cfg.default_dof_pos[j]) * cfg.obs_scale_dof_pos;
    oqd = qd * cfg.obs_scale_dof_vel;
    oa = a;
    s_rw[(RW_LA + j) * QT + lane] = a;              // :181-182
    s_rw[(RW_LDV + j) * QT + lane] = qd;
    // DOF-property re-draw (:591-593, :544-560): rare - written straight to global memory
    if (redraw) {
      float u3[4];
      if (b.dr_u) { u3[0] = b.dr_u[e]; u3[1] = b.dr_u[N + e]; u3[2] = b.dr_u[2 * N + e]; }
      else rng4(args.a.seed, (uint32_t)e, rng_step, RNG_DR, 0, u3);
      const int ix = j * N + e;
      if (cfg.randomize_motor_strength) {
        ms = u3[0] * cfg.motor_strength_lo_span[1] + cfg.motor_strength_lo_span[0];
        b.motor_strengths[ix] = ms;
      }
      if (cfg.randomize_Kp_factor) b.Kp_factors[ix] = u3[1] * cfg.Kp_factor_lo_span[1] + cfg.Kp_factor_lo_span[0];
      if (cfg.randomize_Kd_factor) b.Kd_factors[ix] = u3[2] * cfg.Kd_factor_lo_span[1] + cfg.Kd_factor_lo_span[0];
    }
    pm = clampf((ms - cfg.priv_shift[4]) * cfg.priv_scale[4], -co, co);
    // observation noise of this DOF's q / qd columns (:392): lanes k and 3 + k of the leg's Philox block
    {
      const int cq = 6 + j, cqd = 18 + j;
      if (nu_row) {
        oq += (2.0f * nu_row[cq] - 1.0f) * cfg.noise_scale_core[cq];
        oqd += (2.0f * nu_row[cqd] - 1.0f) * cfg.noise_scale_core[cqd];
      } else {
        oq = __fmaf_rn(2.0f * centered_u16(r4, k), cfg.noise_scale_core[cq], oq);
        oqd = __fmaf_rn(2.0f * centered_u16(r4, 3 + k), cfg.noise_scale_core[cqd], oqd);
      }
    }
    oq = clampf(oq, -co, co); oqd = clampf(oqd, -co, co); oa = clampf(oa, -co, co);
  } else {
    // ---- per-env roles ----
    wait_group(0);
    wait_group(1);
    if (w == 12) {
      // teleport (:768-791), base linear velocity + projected gravity (:159-162), the gravity columns' noise draw
      if (cfg.teleport_robots) {
        float x = root[0], y = root[1];
        const float x0 = x, y0 = y;
        if (x < cfg.teleport_lo_x) x += cfg.teleport_shift_x;
        if (x > cfg.teleport_hi_x) x -= cfg.teleport_shift_x;
        if (y < cfg.teleport_lo_y) y += cfg.teleport_shift_y;
        if (y > cfg.teleport_hi_y) y -= cfg.teleport_shift_y;
        if (x != x0 || y != y0) { root[0] = x; root[1] = y; dirty = true; }
      }
      const float qx = root[3], qy = root[4], qz = root[5], qw = root[6];
      const V3 vw = {root[7], root[8], root[9]};
      const V3 blv = quat_rotate_inverse(qx, qy, qz, qw, vw);
      const V3 grav = quat_rotate_inverse(qx, qy, qz, qw, V3{0.f, 0.f, -1.f});
      s_wo[(WO_BLV + 0) * QT + lane] = blv.x; s_wo[(WO_BLV + 1) * QT + lane] = blv.y; s_wo[(WO_BLV + 2) * QT + lane] = blv.z;
      s_wo[(WO_GRAV + 0) * QT + lane] = grav.x; s_wo[(WO_GRAV + 1) * QT + lane] = grav.y; s_wo[(WO_GRAV + 2) * QT + lane] = grav.z;
      s_wo[(WO_LRV + 0) * QT + lane] = vw.x; s_wo[(WO_LRV + 1) * QT + lane] = vw.y; s_wo[(WO_LRV + 2) * QT + lane] = vw.z;
      if (nu_row) {
        s_gn[0 * QT + lane] = 2.0f * nu_row[0] - 1.0f; s_gn[1 * QT + lane] = 2.0f * nu_row[1] - 1.0f;
        s_gn[2 * QT + lane] = 2.0f * nu_row[2] - 1.0f;
      } else {
        s_gn[0 * QT + lane] = 2.0f * centered_u16(r4, 0); s_gn[1 * QT + lane] = 2.0f * centered_u16(r4, 1);
        s_gn[2 * QT + lane] = 2.0f * centered_u16(r4, 2);
      }
    } else if (w == 13) {
      // termination (:190-202), feet air time (:1619-1631), episode length
      const float* con = s_con + lane * NB * 3;
      bool reset = false;
#pragma unroll 1
      for (int k = 0; k < cfg.n_term_bodies; ++k) {
        const float* f = con + cfg.term_idx[k] * 3;
        reset |= sqrtf((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]) > 1.0f;
      }
      b.reset_buf[e] = reset ? 1 : 0;
      b.episode_length_buf[e] = (int64_t)ep;
      float r_air = 0.f;
      uint32_t nc = 0;
#pragma unroll
      for (int k = 0; k < RL_NUM_FEET; ++k) {
        float air = s_rw[(RW_AIR + k) * QT + lane];
        const bool contact = con[cfg.feet_idx[k] * 3 + 2] > 1.0f;
        const bool filt = contact || ((last_contacts >> (8 * k)) & 0xffu);
        nc |= (contact ? 1u : 0u) << (8 * k);
        const bool first = (air > 0.f) && filt;
        air += cfg.dt;
        r_air += (air - 0.5f) * (first ? 1.f : 0.f);
        air *= filt ? 0.f : 1.f;
        s_rw[(RW_AIR + k) * QT + lane] = air;
      }
      *reinterpret_cast<uint32_t*>(b.last_contacts + (size_t)e * 4) = nc;
      s_air[lane] = r_air;
    } else if (w == 14) {
      // base angular velocity, privileged-observation scalars (:398-417)
      const float qx = root[3], qy = root[4], qz = root[5], qw = root[6];
      const V3 ww = {root[10], root[11], root[12]};
      const V3 bav = quat_rotate_inverse(qx, qy, qz, qw, ww);
      s_wo[(WO_BAV + 0) * QT + lane] = bav.x; s_wo[(WO_BAV + 1) * QT + lane] = bav.y; s_wo[(WO_BAV + 2) * QT + lane] = bav.z;
      s_wo[(WO_LRV + 3) * QT + lane] = ww.x; s_wo[(WO_LRV + 4) * QT + lane] = ww.y; s_wo[(WO_LRV + 5) * QT + lane] = ww.z;
      sc6[0] = clampf((s_ro[RO_FRIC * QT + lane] - cfg.priv_shift[0]) * cfg.priv_scale[0], -co, co);
      sc6[1] = clampf((s_ro[RO_REST * QT + lane] - cfg.priv_shift[1]) * cfg.priv_scale[1], -co, co);
      sc6[2] = clampf((s_ro[RO_PAYLOAD * QT + lane] - cfg.priv_shift[2]) * cfg.priv_scale[2], -co, co);
#pragma unroll
      for (int k = 0; k < 3; ++k) sc6[3 + k] = clampf((s_ro[(RO_COM + k) * QT + lane] - cfg.priv_shift[3]) * cfg.priv_scale[3], -co, co);
    } else {
      // collision (:1550-1553)
      const float* con = s_con + lane * NB * 3;
      float coll = 0.f;
#pragma unroll 1
      for (int k = 0; k < cfg.n_pen_bodies; ++k) {
        const float* f = con + cfg.pen_idx[k] * 3;
        coll += (sqrtf((f[0] * f[0] + f[1] * f[1]) + f[2] * f[2]) > 0.1f) ? 1.f : 0.f;
      }
      s_coll[lane] = coll;
    }
  }
  fence_async_smem();
  __syncthreads();
  if (tid == 480) {                      // the RW / WO blocks are final: they leave under phase 2
    const int wo0 = FUSE ? 0 : WO_BLV;
    tma_store_rows(&args.m_rw, s_rw, tile0, 0);
    tma_store_rows(&args.m_wo, s_wo + wo0 * QT, tile0, wo0);
    bulk_commit();
  }

  // =================================== phase 2 ===========================================================
  {
    // the four-warp kernel's partial sums: per leg ((0 + x0) + x1) + x2, over legs (l0 + l1) + (l2 + l3)
    auto psum = [&](int t) {
      const float* x = s_pd + t * ND * QT + lane;
      float l[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) l[g] = ((0.f + x[(3 * g) * QT]) + x[(3 * g + 1) * QT]) + x[(3 * g + 2) * QT];
      return (l[0] + l[1]) + (l[2] + l[3]);
    };
    auto BLV = [&](int i) { return s_wo[(WO_BLV + i) * QT + lane]; };
    auto BAV = [&](int i) { return s_wo[(WO_BAV + i) * QT + lane]; };
    auto GRAV = [&](int i) { return s_wo[(WO_GRAV + i) * QT + lane]; };
    float* obs = s_obs + lane * W;
    float* priv = s_priv + lane * RL_PRIV_DIM;
    if (w < 12) {
      wait_group(1);                     // (root rows for base_height, accumulator rows: landed long ago - this thread's own observation)
      float r;
      switch (w) {                       // reward_names order of the shipped configuration
        case 0: r = expf(-(sq(cmd.x - BLV(0)) + sq(cmd.y - BLV(1))) / cfg.tracking_sigma); break;    // tracking_lin_vel
        case 1: r = expf(-sq(cmd.z - BAV(2)) / cfg.tracking_sigma_yaw); break;                       // tracking_ang_vel
        case 2: r = sq(BLV(2)); break;                                                               // lin_vel_z
        case 3: r = sq(BAV(0)) + sq(BAV(1)); break;                                                  // ang_vel_xy
        case 4: r = sq(GRAV(0)) + sq(GRAV(1)); break;                                                // orientation
        case 5: r = psum(P_TQ2); break;                                                              // torques
        case 6: r = psum(P_ACC2); break;                                                             // dof_acc
        case 7: r = sq(root[2] - cfg.base_height_target); break;                                     // base_height
        case 8: {                                                                                    // feet_air_time
          const float cmd_xy_norm = sqrtf(cmd.x * cmd.x + cmd.y * cmd.y);
          r = s_air[lane] * ((cmd_xy_norm > 0.1f) ? 1.f : 0.f);
          break;
        }
        case 9: r = s_coll[lane]; break;                                                             // collision
        case 10: r = psum(P_RATE2); break;                                                           // action_rate
        default: r = psum(P_LIM); break;                                                             // dof_pos_limits
      }
      r *= cfg.term_scale[w];
      s_es[w * QT + lane] += r; s_cs[w * QT + lane] += r; s_r[w * QT + lane] = r;
      // this DOF's columns of the output rows (re-using the input tile: every thread passed the barrier above)
      obs[6 + w] = oq; obs[18 + w] = oqd; obs[30 + w] = oa;
      priv[6 + w] = pm;
      if (FUSE) s_tq[lane * ND + w] = tq;
    } else if (w == 12) {
      float g0, g1, g2;
      const float n0 = s_gn[0 * QT + lane], n1 = s_gn[1 * QT + lane], n2 = s_gn[2 * QT + lane];
      if (nu_row) {
        g0 = GRAV(0) + n0 * cfg.noise_scale_core[0];
        g1 = GRAV(1) + n1 * cfg.noise_scale_core[1];
        g2 = GRAV(2) + n2 * cfg.noise_scale_core[2];
      } else {
        g0 = __fmaf_rn(n0, cfg.noise_scale_core[0], GRAV(0));
        g1 = __fmaf_rn(n1, cfg.noise_scale_core[1], GRAV(1));
        g2 = __fmaf_rn(n2, cfg.noise_scale_core[2], GRAV(2));
      }
      obs[0] = clampf(g0, -co, co); obs[1] = clampf(g1, -co, co); obs[2] = clampf(g2, -co, co);
      obs[3] = clampf(cmd.x * cfg.commands_scale[0], -co, co);
      obs[4] = clampf(cmd.y * cfg.commands_scale[1], -co, co);
      obs[5] = clampf(cmd.z * cfg.commands_scale[2], -co, co);
      if (dirty) s_root_dirty = 1;
    } else if (w == 13) {
      const float bx = BLV(0), wz = BAV(2);
      float* x = s_cs + 12 * QT + lane;
      x[0 * QT] += bx;                      // lin_vel_raw (:336-340)
      x[1 * QT] += wz;                      // ang_vel_raw
      x[2 * QT] += sq(bx - cmd.x);          // lin_vel_residual
      x[3 * QT] += sq(wz - cmd.z);          // ang_vel_residual
      x[4 * QT] += 1.0f;                    // ep_timesteps
    } else if (w == 14) {
#pragma unroll
      for (int k = 0; k < 6; ++k) priv[k] = sc6[k];
    }
  }
  fence_async_smem();
  __syncthreads();

  // =================================== stores + phase 3 ==================================================
  if (tid == 0) {
    tma_store_rows(&args.m_es12, s_es, tile0, 0);
    tma_store_rows(&args.m_cs12, s_cs, tile0, 0);
    tma_store_rows(&args.m_cs5, s_cs + 12 * QT, tile0, RL_ROW_EXTRAS);
    bulk_s2g(b.obs_buf + (size_t)tile0 * W, s_obs, QT * W * 4);
    bulk_s2g(b.privileged_obs_buf + (size_t)tile0 * RL_PRIV_DIM, s_priv, QT * RL_PRIV_DIM * 4);
    if (FUSE) bulk_s2g(b.torques + (size_t)tile0 * ND, s_tq, QT * ND * 4);
    if (s_root_dirty) bulk_s2g(b.root_states + (size_t)tile0 * 13, s_root, QT * 13 * 4);
    bulk_commit();
  }
  if (w == 15) {
    // compute_reward's tail (:314-340): sum in reward_names order, positive clip, total row
    float rew = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) rew += s_r[i * QT + lane];
    if (b.rew_raw) b.rew_raw[e] = rew;
    rew = fmaxf(rew, 0.f);
    b.episode_sums[RL_ROW_TOTAL * N + e] = s_es[12 * QT + lane] + rew;
    b.rew_buf[e] = rew;
  }
  if (tid == 480 || tid == 0) bulk_wait_read0();     // each issuing thread: its stores have read their shared-memory source
}

// ---- persistent variant: a few CTAs per SM, each with TWO tile buffers ---------------------------------------------
// With one tile per CTA and the whole grid resident (32768 envs = 1024 CTAs, 7 per SM) every tile is requested in the
// first microsecond, all of them land together ~5 us later, and only then does anybody compute: load and compute phases
// are serial (profiles/r02_env_step.md).  Here a CTA keeps one tile loading while it computes the other: while it works
// on buffer k & 1, the copies of its next tile fly into the other buffer, and the results of the tile before leave
// through bulk stores.  Tiles are handed out by a queue in global memory (RlEnvBuffers.tile_queue; without it CTA c takes
// tiles c, c + grid, ...): the CTAs of an SM drift apart, so at any moment some are loading and some computing.
template <bool FUSE>
__global__ void __launch_bounds__(QTHREADS, 3)
env_step_rows_persistent_kernel(const __grid_constant__ RowsArgs args, int buf_bytes) {
  extern __shared__ __align__(128) uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar[4];      // two buffers x two landing groups
  __shared__ int s_root_dirty;
  __shared__ int s_tile[2];
  const int tid = threadIdx.x;
  const RlEnvBuffers& b = args.a.b;
  const int n_tiles = args.a.cfg.num_envs / QT;
  unsigned int* q = b.tile_queue;
  auto grab = [&](int cur) -> int {            // thread 0 only
    const int nx = q ? (int)(gridDim.x + atomicAdd(q, 1u)) : cur + (int)gridDim.x;
    return nx < n_tiles ? nx : -1;
  };
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&s_bar[i], 1);
    mbar_fence_init();
    const int t0 = blockIdx.x;
    s_tile[0] = t0;
    issue_tile_loads<FUSE>(args, smem_dyn, &s_bar[0], t0 * QT);
    const int t1 = grab(t0);
    s_tile[1] = t1;
    if (t1 >= 0) issue_tile_loads<FUSE>(args, smem_dyn + buf_bytes, &s_bar[2], t1 * QT);
  }
  __syncthreads();
  for (int k = 0;; ++k) {
    const int bi = k & 1;
    const int tile = s_tile[bi];
    if (tile < 0) break;                       // uniform: the queue is empty
    uint8_t* buf = smem_dyn + bi * buf_bytes;
    if (tid == 0) s_root_dirty = 0;            // (tile_body's first barrier publishes it)
    tile_body<FUSE, false, false>(args, buf, &s_bar[2 * bi], (uint32_t)((k >> 1) & 1), tile * QT, &s_root_dirty, nullptr);
    if (tid == 96) bulk_wait_read0();          // (the early store group of the RW / WO blocks)
    __syncthreads();                           // every thread is done with this buffer (phase 3 read it)
    if (tid == 0) {
      bulk_wait_read0();                       // ... and so are its bulk stores: the buffer may be refilled
      const int tn = grab(tile);
      s_tile[bi] = tn;                         // read two iterations from now (a barrier lies in between)
      if (tn >= 0) issue_tile_loads<FUSE>(args, buf, &s_bar[2 * bi], tn * QT);
    }
  }
  if (tid == 0) {
    bulk_wait_read0();
    if (q) {
      const unsigned int done = atomicAdd(q + 1, 1u);
      if (done == gridDim.x - 1) { q[0] = 0u; q[1] = 0u; __threadfence(); }
    }
    if (b.step_state) {
      const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(b.step_state + 1), 1ull);
      if (done == gridDim.x - 1) {
        b.step_state[1] = 0;
        atomicAdd(reinterpret_cast<unsigned long long*>(b.step_state), 1ull);
      }
    }
  }
}

// ---- host side --------------------------------------------------------------------------------------------
struct MapKey {
  const void* base; uint64_t rows, n; uint32_t box;
  bool operator<(const MapKey& o) const { return std::tie(base, rows, n, box) < std::tie(o.base, o.rows, o.n, o.box); }
};

static int cached_map(CUtensorMap* out, const void* base, uint64_t rows, uint64_t n, uint32_t box) {
  // descriptors are pure functions of (base, rows, n, box): build once per env instance
  static thread_local std::map<MapKey, CUtensorMap> cache;
  const MapKey k{base, rows, n, box};
  auto it = cache.find(k);
  if (it == cache.end()) {
    CUtensorMap m;
    const int rc = make_tmap_f32_rows(&m, base, rows, n, box);
    if (rc != RL_OK) return rc;
    if (cache.size() > 4096) cache.clear();
    it = cache.emplace(k, m).first;
  }
  *out = it->second;
  return RL_OK;
}

template <bool FUSE, int MINB, bool TRACE>
static int launch_inst_t(const RowsArgs& ra, size_t smem, cudaStream_t st) {
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t err = cudaFuncSetAttribute(env_step_rows_kernel<FUSE, MINB, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    configured = smem;
  }
  // RL_ENV_PDL=1: programmatic dependent launch.  Measured inside the K-step CUDA graph (us per launch, PDL off / on):
  // 4000 envs 5.94 / 8.51, 32768 envs 13.23 / 13.05, 262144 envs 66.05 / 65.89 - programmatic edges cost more than
  // they hide for small grids, so it stays off by default
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("RL_ENV_PDL"); pdl = (e && atoi(e) == 1) ? 1 : 0; }
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(ra.a.cfg.num_envs / QT); lc.blockDim = dim3(QTHREADS); lc.dynamicSmemBytes = smem; lc.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at; lc.numAttrs = pdl ? 1 : 0;
  cudaError_t err = cudaLaunchKernelEx(&lc, env_step_rows_kernel<FUSE, MINB, TRACE>, ra);
  RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "env_step_rows_kernel launch: %s", cudaGetErrorString(err));
  return check_launch("env_step_rows_kernel");
}

// (the profiling stamps are a template parameter: the shipped instantiation carries none of their instructions)
template <bool FUSE, int MINB>
static int launch_inst(const RowsArgs& ra, size_t smem, cudaStream_t st) {
  return ra.trace ? launch_inst_t<FUSE, MINB, true>(ra, smem, st) : launch_inst_t<FUSE, MINB, false>(ra, smem, st);
}

template <bool FUSE>
static int launch_wide(const RowsArgs& ra, size_t smem, cudaStream_t st) {
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t err = cudaFuncSetAttribute(env_step_rows_wide_kernel<FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    configured = smem;
  }
  static int pdl = -1;          // RL_ENV_PDL=1: programmatic dependent launch (A/B knob, see launch_inst_t)
  if (pdl < 0) { const char* e = getenv("RL_ENV_PDL"); pdl = (e && atoi(e) == 1) ? 1 : 0; }
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(ra.a.cfg.num_envs / QT); lc.blockDim = dim3(WIDE_THREADS); lc.dynamicSmemBytes = smem; lc.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at; lc.numAttrs = pdl ? 1 : 0;
  cudaError_t err = cudaLaunchKernelEx(&lc, env_step_rows_wide_kernel<FUSE>, ra);
  RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "env_step_rows_wide_kernel launch: %s", cudaGetErrorString(err));
  return check_launch("env_step_rows_wide_kernel");
}

template <bool FUSE>
static int launch_persistent(const RowsArgs& ra, int buf_bytes, int grid, cudaStream_t st) {
  static size_t configured = 0;
  const size_t smem = 2 * (size_t)buf_bytes;
  if (smem > configured) {
    cudaError_t err = cudaFuncSetAttribute(env_step_rows_persistent_kernel<FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(err));
    configured = smem;
  }
  env_step_rows_persistent_kernel<FUSE><<<grid, QTHREADS, smem, st>>>(ra, buf_bytes);
  return check_launch("env_step_rows_persistent_kernel");
}

}  // namespace rows

static unsigned long long* g_trace_buf = nullptr;
static int g_trace_ctas = 0;
}  // namespace rl
// profiling aid: per-CTA globaltimer stamps of the next env_step_rows launches {entry, loads issued, tile landed, phase 1
// done, phase 2 done, stores issued, stores read, smid}.  enable > 0: allocate for `enable` CTAs and switch on; 0: off.
// out_host (optional): receives capacity_ctas x 8 values of the last traced launch.
extern "C" int rl_debug_env_rows_trace(int32_t enable, uint64_t* out_host, int32_t capacity_ctas) {
  using namespace rl;
  if (out_host && g_trace_buf) {
    const int n = capacity_ctas < g_trace_ctas ? capacity_ctas : g_trace_ctas;
    cudaError_t err = cudaMemcpy(out_host, g_trace_buf, (size_t)n * rows::TRACE_STAMPS * 8, cudaMemcpyDeviceToHost);
    RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "rl_debug_env_rows_trace: %s", cudaGetErrorString(err));
  }
  if (enable > 0 && enable > g_trace_ctas) {
    if (g_trace_buf) cudaFree(g_trace_buf);
    cudaError_t err = cudaMalloc(&g_trace_buf, (size_t)enable * rows::TRACE_STAMPS * 8);
    RL_REQUIRE(err == cudaSuccess, RL_ERR_CUDA, "rl_debug_env_rows_trace: %s", cudaGetErrorString(err));
    g_trace_ctas = enable;
  }
  if (enable == 0 && !out_host) g_trace_ctas = g_trace_ctas;   // buffer kept for later read-back
  rows_trace_on = enable > 0;
  return RL_OK;
}
namespace rl {
// true when the env-owned state pointers form the packed RO / RW / WO blocks and every tile is full and aligned
bool rows_layout_ok(const StepArgs& a) {
  const RlEnvCfg& c = a.cfg;
  const RlEnvBuffers& b = a.b;
  const size_t N = (size_t)c.num_envs;
  if (c.num_envs % rows::QT != 0 || c.num_obs != 42 || c.num_actions != ND) return false;
  if (b.Kd_factors != b.Kp_factors + 12 * N || b.motor_strengths != b.Kp_factors + 24 * N ||
      b.friction_coeffs != b.Kp_factors + 36 * N || b.restitutions != b.Kp_factors + 37 * N ||
      b.payloads != b.Kp_factors + 38 * N || b.com_displacements != b.Kp_factors + 39 * N)
    return false;
  if (b.last_dof_vel != b.last_actions + 12 * N || b.feet_air_time != b.last_actions + 24 * N) return false;
  if (b.base_lin_vel != b.joint_pos_target + 12 * N || b.base_ang_vel != b.joint_pos_target + 15 * N ||
      b.projected_gravity != b.joint_pos_target + 18 * N || b.last_root_vel != b.joint_pos_target + 21 * N)
    return false;
  const uintptr_t al = (uintptr_t)b.Kp_factors | (uintptr_t)b.last_actions | (uintptr_t)b.joint_pos_target |
                       (uintptr_t)b.episode_sums | (uintptr_t)b.command_sums | (uintptr_t)b.root_states | (uintptr_t)b.dof_state |
                       (uintptr_t)b.contact_forces | (uintptr_t)b.actions_in | (uintptr_t)b.torques | (uintptr_t)b.obs_buf |
                       (uintptr_t)b.privileged_obs_buf;
  return (al & 15) == 0;
}

int launch_step_rows(const StepArgs& a, bool fuse, cudaStream_t st) {
  using namespace rows;
  RowsArgs ra;
  ra.a = a;
  ra.trace = (rows_trace_on && g_trace_buf && a.cfg.num_envs / QT <= g_trace_ctas) ? g_trace_buf : nullptr;
  ra.h.stagger_ns = 0; ra.h.stagger_from = 0; ra.h.stagger_group = 0;      // set below, once the grid shape is known
  ra.h.dof_state = a.b.dof_state; ra.h.actions = a.b.actions_in; ra.h.root_states = a.b.root_states;
  ra.h.contact_forces = a.b.contact_forces; ra.h.torques_in = a.b.torques; ra.h.num_bodies = a.cfg.num_bodies;
  ra.h.pad[0] = ra.h.pad[1] = 0;
  const RlEnvBuffers& b = a.b;
  const uint64_t N = (uint64_t)a.cfg.num_envs;
  int rc;
  if ((rc = cached_map(&ra.m_ro, b.Kp_factors, RO_ROWS, N, RO_ROWS)) != RL_OK) return rc;
  if ((rc = cached_map(&ra.m_rw, b.last_actions, RW_ROWS, N, RW_ROWS)) != RL_OK) return rc;
  if ((rc = cached_map(&ra.m_wo, b.joint_pos_target, WO_ROWS, N, fuse ? WO_ROWS : WO_ROWS - WO_BLV)) != RL_OK) return rc;
  if ((rc = cached_map(&ra.m_es12, b.episode_sums, RL_EPISODE_ROWS, N, 12)) != RL_OK) return rc;
  if ((rc = cached_map(&ra.m_es1, b.episode_sums, RL_EPISODE_ROWS, N, 1)) != RL_OK) return rc;
  if ((rc = cached_map(&ra.m_cs12, b.command_sums, RL_COMMAND_ROWS, N, 12)) != RL_OK) return rc;
  if ((rc = cached_map(&ra.m_cs5, b.command_sums, RL_COMMAND_ROWS, N, 5)) != RL_OK) return rc;
  const Lay L(a.cfg.num_bodies, fuse);
  const size_t smem = (size_t)L.total;
  {
    // persistent variant (two tile buffers per CTA, 3 CTAs per SM): opt-in (RL_ENV_PERSIST=1 or rl_debug_env_rows(2)).
    // Measured against one tile per CTA (us per launch): 32768 envs 18.8 vs 13.6, 262144 envs 79.2 vs 66.7 - a tile takes
    // ~4.3 us through the 4 warps of one CTA (long dependent chains: Philox, sqrt, exp), so 12 resident warps per SM
    // compute slower than 28 even though their loads are hidden; kept because it is bit-identical and shows the bound
    static int persist = -2, sms = 0;
    if (persist == -2) {
      const char* e = getenv("RL_ENV_PERSIST");
      persist = e ? atoi(e) : -1;
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (sms <= 0) sms = 148;
    }
    const int n_tiles = a.cfg.num_envs / QT;
    const int mode = g_rows_persist_mode >= 0 ? g_rows_persist_mode : persist;
    // wide variant (16 warps per tile) for grids of at most RL_ENV_WIDE_MAX CTAs per SM (default below; 0: never);
    // rl_debug_env_rows(4) forces it, (3) forces the four-warp kernel
    static int wide_max = -1;
    if (wide_max < 0) { const char* e = getenv("RL_ENV_WIDE_MAX"); wide_max = e ? atoi(e) : 2; }
    if (mode == 2 || (mode < 0 && n_tiles <= wide_max * sms)) {
      const size_t wsmem = (size_t)L.total + (NPART * ND + 3) * ROWB;
      return fuse ? rows::launch_wide<true>(ra, wsmem, st) : rows::launch_wide<false>(ra, wsmem, st);
    }
    if (mode == 1) {
      const int buf_bytes = (L.total + 127) & ~127;
      const int grid = n_tiles < 3 * sms ? n_tiles : 3 * sms;
      return fuse ? rows::launch_persistent<true>(ra, buf_bytes, grid, st) : rows::launch_persistent<false>(ra, buf_bytes, grid, st);
    }
  }
  // 7 CTAs / SM when the tile fits (13 bodies), else 6
  static int force = -1;
  if (force < 0) { const char* m = getenv("RL_ROWS_MINB"); force = m ? atoi(m) : 0; }
  const bool seven = force == 7 || (force == 0 && 7 * (smem + 1024 + 128) <= 233472);
  {
    // Time-staggered copies for a grid that is resident all at once (one wave, several CTAs per SM): with every tile
    // requested at t = 0 they all land together ~5 us later and nobody computes before that.  The CTAs of the later
    // "rows" of the wave (blockIdx >= 2 x SMs, then every further 2 x SMs) sleep 1 / 2 / .. us before they issue, so their
    // requests queue up BEHIND the first ones: the early tiles land sooner and are computed while the rest streams in.
    // Measured per launch (us, off -> on): 16384 envs 9.95 -> 9.27, 32768 envs 13.63 -> 12.29, Go1 32768 14.79 -> 12.76;
    // deeper grids (waves start as CTAs retire) and grids of <= 2 CTAs per SM are left alone.  RL_ENV_STAGGER=
    // "ns,first_cta,ctas_per_group" overrides ("0": off).
    static int sms = 0, env_ns = -1, env_from = 0, env_group = 0;
    if (!sms) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (sms <= 0) sms = 148;
      const char* e = getenv("RL_ENV_STAGGER");
      if (e) sscanf(e, "%d,%d,%d", &env_ns, &env_from, &env_group);
    }
    const int n_tiles = a.cfg.num_envs / QT;
    if (env_ns >= 0) { ra.h.stagger_ns = env_ns; ra.h.stagger_from = env_from; ra.h.stagger_group = env_group; }
    else if (n_tiles > 2 * sms && n_tiles <= (seven ? 7 : 6) * sms) {
      // (re-tuned after the compute phase got shorter: with more than four CTAs per SM the first group is three CTA rows and
      // the delay 1.1 us - 32768 envs: 11.13 -> 10.79 us per launch; sweep in profiles/r02_env_step.md)
      // (last sweep of round 2, profiles/jobs/r2_job60.sh / r2_job61.sh: with more than six CTAs per SM a first group of
      // 3.25 - 4 CTA rows is a plateau - 32768 envs 10.18 -> 9.98 us per launch with 3.5 rows; Go1 and 24576 envs unchanged)
      const bool deep = n_tiles > 4 * sms;
      ra.h.stagger_ns = deep ? 1100 : 1000;
      ra.h.stagger_from = n_tiles > 6 * sms ? (7 * sms) / 2 : (deep ? 3 : 2) * sms;
      ra.h.stagger_group = 2 * sms;
    }
  }
  if (seven) return fuse ? launch_inst<true, 7>(ra, smem, st) : launch_inst<false, 7>(ra, smem, st);
  return fuse ? launch_inst<true, 6>(ra, smem, st) : launch_inst<false, 6>(ra, smem, st);
}

}  // namespace rl
