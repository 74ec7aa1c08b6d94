/*
 * rl_b200.h - C-ABI of the B200-native hot path for rapid-locomotion-rl.
 *
 * The reference (dhruvmetha/rapid-locomotion-rl) is pure Python/PyTorch and has no
 * FFI of its own; the drop-in boundary is its Python object API.
 * This header is the thin C boundary that the Python mirror of that API binds
 * with ctypes.  Each entry point cites the reference code it replaces
 * (paths relative to the reference root).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - every pointer is a DEVICE pointer borrowed for the duration of the call
 *     unless the name ends in _host; the library keeps no reference to it.
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises the host,
 *     so every call is CUDA-graph capturable.
 *   - return value: 0 = RL_OK, negative = error (rl_last_error() has the text).
 *   - persistent env state owned by the caller is stored SoA: a `[K][N]` array
 *     holds field k of env n at `p[k*N + n]`.  Simulator-owned tensors keep the
 *     PhysX AoS layout (`[N,13]` root rows, `[N,12,2]` dof rows, `[N,NB,3]`
 *     contact rows).
 */
#ifndef RL_B200_H
#define RL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RL_OK 0
#define RL_ERR_BAD_ARG (-1)
#define RL_ERR_BAD_CFG (-2)
#define RL_ERR_CUDA (-3)
#define RL_ERR_NCCL (-4)
#define RL_ERR_UNSUPPORTED (-5)

#define RL_NUM_DOF 12
#define RL_NUM_FEET 4
#define RL_MAX_BODIES 24
#define RL_MAX_TERMS 21
#define RL_PRIV_DIM 18
#define RL_MAX_CORE_OBS 64                       /* observation columns before the height samples */
/* Fixed row layout of the per-env reward accumulators (rows of unused terms are never touched):
 *   episode_sums [RL_EPISODE_ROWS][N]: row i = enabled term i, RL_ROW_TERMINATION, RL_ROW_TOTAL
 *   command_sums [RL_COMMAND_ROWS][N]: row i = enabled term i, RL_ROW_TERMINATION, then
 *     RL_ROW_EXTRAS + {0 lin_vel_raw, 1 ang_vel_raw, 2 lin_vel_residual, 3 ang_vel_residual, 4 ep_timesteps} */
#define RL_ROW_TERMINATION RL_MAX_TERMS
#define RL_ROW_TOTAL (RL_MAX_TERMS + 1)
#define RL_ROW_EXTRAS (RL_MAX_TERMS + 1)
#define RL_EPISODE_ROWS (RL_MAX_TERMS + 2)
#define RL_COMMAND_ROWS (RL_MAX_TERMS + 6)

/* Reward term ids; formulas: legged_robot.py:1506-1646 (_reward_<name>). */
enum RlRewardTerm {
  RL_REW_TRACKING_LIN_VEL = 0, /* :1578 */
  RL_REW_TRACKING_ANG_VEL = 1, /* :1614 */
  RL_REW_LIN_VEL_Z = 2,        /* :1506 */
  RL_REW_ANG_VEL_XY = 3,       /* :1510 */
  RL_REW_ORIENTATION = 4,      /* :1514 */
  RL_REW_TORQUES = 5,          /* :1523 */
  RL_REW_DOF_ACC = 6,          /* :1539 */
  RL_REW_BASE_HEIGHT = 7,      /* :1518 */
  RL_REW_FEET_AIR_TIME = 8,    /* :1619 */
  RL_REW_COLLISION = 9,        /* :1547 */
  RL_REW_ACTION_RATE = 10,     /* :1543 */
  RL_REW_DOF_POS_LIMITS = 11,  /* :1560 */
  RL_REW_ENERGY = 12,          /* :1527 */
  RL_REW_ENERGY_EXPENDITURE = 13, /* :1531 */
  RL_REW_DOF_VEL = 14,         /* :1535 */
  RL_REW_SURVIVAL = 15,        /* :1556 */
  RL_REW_DOF_VEL_LIMITS = 16,  /* :1566 */
  RL_REW_TORQUE_LIMITS = 17,   /* :1573 */
  RL_REW_STUMBLE = 18,         /* :1633 */
  RL_REW_STAND_STILL = 19,     /* :1638 */
  RL_REW_FEET_CONTACT_FORCES = 20, /* :1643 */
  RL_REW_TERMINATION = 21,     /* :1552, added after the positive clip (:330-334) */
  RL_REW_COUNT = 22
};

/* Frozen view of the reference's `Cfg` namespace (legged_robot_config.py:6-256)
 * plus robot constants, resolved once at env construction (legged_robot.py
 * _parse_cfg :1417-1429, _init_buffers :935-1028, _prepare_reward_function
 * :1074-1110, _get_noise_scale_vec :882-932).  Plain data, passed by value. */
typedef struct RlEnvCfg {
  int32_t num_envs;
  int32_t num_bodies;         /* NB: 13 Mini Cheetah, 17 Go1 */
  int32_t num_actions;        /* >= 12; only the first 12 drive joints (:665) */
  int32_t num_obs;            /* env.num_observations */
  int32_t num_height_points;  /* 187 when measure_heights, else 0 */
  int32_t control_type;       /* 0 'P', 1 'V', 2 'T' (:667-685) */
  int32_t decimation;         /* control.decimation (:116) */
  float sim_dt;               /* float32(sim.dt) */
  float dt;                   /* float32(decimation * sim_dt) (:1418) */
  float action_scale;         /* control.action_scale */
  float hip_scale_reduction;  /* control.hip_scale_reduction, dofs 0,3,6,9 (:666) */
  float clip_actions;         /* normalization.clip_actions (:112) */
  float clip_obs;             /* normalization.clip_observations (:133) */
  float p_gains[RL_NUM_DOF];
  float d_gains[RL_NUM_DOF];
  float default_dof_pos[RL_NUM_DOF];
  float torque_limits[RL_NUM_DOF];
  float dof_pos_lo[RL_NUM_DOF]; /* soft limits (:512-515) */
  float dof_pos_hi[RL_NUM_DOF];
  float dof_vel_limits[RL_NUM_DOF];
  int32_t feet_idx[RL_NUM_FEET];
  int32_t n_term_bodies;
  int32_t term_idx[RL_MAX_BODIES]; /* termination_contact_indices (:1295) */
  int32_t n_pen_bodies;
  int32_t pen_idx[RL_MAX_BODIES];  /* penalised_contact_indices (:1288) */
  /* rewards: enabled terms in reward_names order; scale already multiplied by dt
   * in double then rounded to float (:1084, :322).  Row i of episode_sums /
   * command_sums belongs to enabled term i (fixed row layout above). */
  int32_t n_terms;
  int32_t term_id[RL_MAX_TERMS];
  float term_scale[RL_MAX_TERMS];
  uint32_t term_mask;              /* bit id set for every enabled term */
  int32_t has_termination;         /* scales.termination != 0 (:330-334) */
  float termination_scale;
  int32_t only_positive_rewards;
  float tracking_sigma;
  float tracking_sigma_yaw;
  float base_height_target;
  float soft_dof_vel_limit;
  float soft_torque_limit;
  float max_contact_force;
  int32_t use_terminal_body_height;
  float terminal_body_height;
  int32_t global_reference;        /* commands.global_reference (:1580) */
  /* observations (:342-417) */
  int32_t observe_command;
  int32_t observe_vel;
  int32_t observe_only_ang_vel;
  int32_t observe_only_lin_vel;
  int32_t observe_yaw;
  int32_t measure_heights;
  int32_t add_noise;
  float obs_scale_lin_vel;
  float obs_scale_ang_vel;
  float obs_scale_dof_pos;
  float obs_scale_dof_vel;
  float obs_scale_height;
  float commands_scale[3];
  /* noise amplitude per observation column (:882-932): the non-height columns sit here so that
   * the kernel reads them as constant-bank operands; every height column uses noise_scale_height */
  float noise_scale_core[RL_MAX_CORE_OBS];
  float noise_scale_height;
  /* privileged obs: (x - shift) * scale for friction, restitution, payload,
   * com_displacement, motor_strength (:398-417); scale 0 when not observed */
  float priv_scale[5];
  float priv_shift[5];
  /* teleport (:768-791) */
  int32_t teleport_robots;
  float teleport_lo_x;       /* thresh + int(x_offset * horizontal_scale)          (:776) */
  float teleport_hi_x;       /* terrain_length * num_rows - thresh + x_offset      (:780) */
  float teleport_shift_x;    /* terrain_length * (num_rows - 1)                    (:777) */
  float teleport_lo_y;       /* thresh                                             (:783) */
  float teleport_hi_y;       /* terrain_width * num_cols - thresh                  (:787) */
  float teleport_shift_y;    /* terrain_width * (num_cols - 1)                     (:784) */
  /* heights (:1469-1503) */
  int32_t heights_plane;     /* mesh_type == 'plane' => zeros */
  float border_size;
  float horizontal_scale;
  float vertical_scale;
  int32_t hf_rows;
  int32_t hf_cols;
  /* domain randomisation re-draw inside the step (:591-593, :544-560) and push (:757-766) */
  int32_t rand_interval;
  int32_t randomize_motor_strength;
  int32_t randomize_Kp_factor;
  int32_t randomize_Kd_factor;
  /* draws are u * span + lo with span = float32(hi - lo) evaluated in double first,
   * exactly as the eager expression `rand * (hi - lo) + lo` rounds */
  float motor_strength_lo_span[2];
  float Kp_factor_lo_span[2];
  float Kd_factor_lo_span[2];
  int32_t push_robots;
  int32_t push_interval;
  float push_lo_span[2];     /* (-max_push_vel_xy, 2*max_push_vel_xy) (:764) */
  /* upstream call order switch: 0 = as written in this fork */
  int32_t timeout_resets;    /* 1: time_out_buf = ep_len > max_episode_length, OR into reset (:197-198) */
  int32_t max_episode_length;
  /* train / eval env split (legged_robot.py:456-469, base_task.py:43-49): envs [num_train_envs, num_envs) are the
   * evaluation envs.  The reference hands the EVALUATION Cfg to the functions it calls through _call_train_eval -
   * inside the step these are _teleport_robots :576, _push_robots :588 and _randomize_dof_props :593 - and its own
   * Cfg to everything else (rewards, observations, noise, terminations, rand_interval :591).  The eval_* fields are the
   * evaluation configuration's values of exactly those fields.  num_train_envs == num_envs (or 0): no split. */
  int32_t num_train_envs;
  int32_t eval_teleport_robots;
  float eval_teleport_lo_x;
  float eval_teleport_hi_x;
  float eval_teleport_shift_x;
  float eval_teleport_lo_y;
  float eval_teleport_hi_y;
  float eval_teleport_shift_y;
  int32_t eval_randomize_motor_strength;
  int32_t eval_randomize_Kp_factor;
  int32_t eval_randomize_Kd_factor;
  float eval_motor_strength_lo_span[2];
  float eval_Kp_factor_lo_span[2];
  float eval_Kd_factor_lo_span[2];
  int32_t eval_push_robots;
  int32_t eval_push_interval;
  float eval_push_lo_span[2];
} RlEnvCfg;

/* Buffers of one vectorised env (all device pointers, caller-owned). */
typedef struct RlEnvBuffers {
  /* simulator-owned AoS tensors (legged_robot.py:950-971) */
  float* root_states;           /* [N,13] in/out: teleport and push write back */
  const float* dof_state;       /* [N,12,2] */
  const float* contact_forces;  /* [N,NB,3] */
  const float* actions_in;      /* [N,num_actions] raw policy output */
  float* torques;               /* [N,12] out (fused) or in (post_physics) */
  /* step outputs */
  float* obs_buf;               /* [N,num_obs] */
  float* privileged_obs_buf;    /* [N,18] */
  float* rew_buf;               /* [N] */
  uint8_t* reset_buf;           /* [N] bool */
  uint8_t* time_out_buf;        /* [N] bool (only written when timeout_resets) */
  float* measured_heights;      /* [N,P] or NULL */
  /* persistent state, SoA [K][N] */
  float* last_actions;          /* [12][N]; equals the clipped actions after step */
  float* last_dof_vel;          /* [12][N] */
  float* last_root_vel;         /* [6][N] */
  float* joint_pos_target;      /* [12][N] */
  float* base_lin_vel;          /* [3][N] */
  float* base_ang_vel;          /* [3][N] */
  float* projected_gravity;     /* [3][N] */
  float* Kp_factors;            /* [12][N] */
  float* Kd_factors;            /* [12][N] */
  float* motor_strengths;       /* [12][N] */
  const float* friction_coeffs; /* [N] */
  const float* restitutions;    /* [N] */
  const float* payloads;        /* [N] */
  const float* com_displacements; /* [3][N] */
  float* feet_air_time;         /* [4][N] */
  uint8_t* last_contacts;       /* [N,4] bool */
  int64_t* episode_length_buf;  /* [N] */
  const float* commands;        /* [N,4] */
  float* episode_sums;          /* [RL_EPISODE_ROWS][N] */
  float* command_sums;          /* [RL_COMMAND_ROWS][N] */
  /* read-only tables */
  const float* height_points;   /* [P,2] base-frame xy (:1453-1467) */
  const int16_t* height_samples;/* [hf_rows,hf_cols] (:1141) */
  /* optional injected uniforms in [0,1) for parity tests; NULL => Philox4x32-10
   * keyed by (seed, env, step, stream) */
  const float* noise_u;         /* [N,num_obs]  replaces torch.rand_like (:392) */
  const float* dr_u;            /* [3][N] motor, Kp, Kd draws (:547-558) */
  const float* push_u;          /* [2][N] (:764) */
  /* optional device-side step counter for CUDA-graph replay (kernel arguments are frozen in a
   * graph): [0] is added to the `step` argument to key the RNG and is incremented once per
   * launch by the last CTA to finish; [1] is that CTA ticket.  NULL => host `step` only. */
  uint64_t* step_state;         /* [2] */
  /* optional [N]: the sum of the enabled terms BEFORE the positive clip and the termination term (:328-334).  The
   * Python plugin layer needs it to add user-defined `_reward_<name>` terms exactly where the reference adds them. */
  float* rew_raw;
  /* optional [N] scratch.  When set (and measure_heights is on) the terrain heights (:1469-1503: measured_heights, the
   * height suffix of obs_buf, mean(z - heights) for the base_height term) are sampled by a launch of their own in front
   * of the step kernel - one warp per env over the whole GPU - instead of by the four warps of each 32-env CTA.  Same
   * arithmetic and summation order: identical bits. */
  float* height_mean;
  /* optional [2] (zero-initialised, owned by the library between launches): tile queue of the persistent step kernel
   * (env_step_rows.cu) - {next tile, CTAs done}; without it the tiles are dealt round-robin. */
  uint32_t* tile_queue;
} RlEnvBuffers;

/* profiling aid: enable / read the globaltimer phase stamps of one CTA of the fused env-step kernel
 * (entry, SoA loads issued, staged rows landed, phase 1 start / end, barrier, outputs written, barrier,
 * stores issued); out_host16 may be NULL */
int rl_debug_env_trace(int32_t enable, uint64_t* out_host16);
/* testing aid: selects the kernel behind rl_env_step_fused / rl_env_post_physics for the shipped configuration on
 * packed state blocks: 1 = env_step_rows.cu (all tile traffic on TMA; the default), 0 = env_step_quad.cu, -1 = back to
 * the default (environment variable RL_ENV_ROWS); 2 / 3 = env_step_rows.cu with its persistent two-buffer variant forced
 * on / off (default: off - it measured slower; RL_ENV_PERSIST=1).  Returns the previous setting.  All of
 * them produce identical bits. */
int rl_debug_env_rows(int32_t mode);
/* profiling aid: per-CTA globaltimer stamps of the following env_step_rows launches {entry, loads issued, tile landed,
 * phase 1 done, phase 2 done, stores issued, stores read, smid}.  enable > 0: number of CTAs to record (switches tracing
 * on), 0: off.  out_host (optional): receives min(capacity_ctas, recorded) x 8 values of the last traced launch. */
int rl_debug_env_rows_trace(int32_t enable, uint64_t* out_host, int32_t capacity_ctas);
const char* rl_last_error(void);
const char* rl_version(void);
/* sizeof(struct <name>) as compiled into the library (-1 for an unknown name): lets a
 * binding generated from this header check its layout at load time. */
int64_t rl_sizeof(const char* name);

/* legged_robot.py:653-688 _compute_torques.  Reads actions_in, dof_state,
 * Kp/Kd/motor factors; writes torques [N,12] and joint_pos_target.  This is the
 * entry a real gymapi integration calls `decimation` times per step (:116-126). */
int rl_env_torques(const RlEnvCfg* cfg_host, const RlEnvBuffers* bufs_host, void* stream);

/* legged_robot.py:139-188 post_physics_step + :133-136 observation clip, with the
 * torques taken from bufs->torques (already applied to the simulator). */
int rl_env_post_physics(const RlEnvCfg* cfg_host, const RlEnvBuffers* bufs_host, uint64_t seed,
                        uint64_t step, void* stream);

/* legged_robot.py:106-137 step() on synthetic simulator state: torques (once;
 * without physics the `decimation` evaluations are identical) + post-physics
 * pipeline in ONE kernel.  The benchmark entry (SURVEY 8d metric 1). */
int rl_env_step_fused(const RlEnvCfg* cfg_host, const RlEnvBuffers* bufs_host, uint64_t seed,
                      uint64_t step, void* stream);

/* legged_robot.py:1469-1503 _get_heights as a launch of its own: out[i, p] = height under measured point p of env
 * env_ids[i] (NULL: envs 0..n_ids-1), from the current root_states [N,13], the base-frame points [P,2] and the int16
 * table.  The fused step samples the same values in-kernel; this serves direct callers of the method. */
int rl_env_heights(const RlEnvCfg* cfg_host, const float* root_states, const float* height_points,
                   const int16_t* height_samples, const int64_t* env_ids, int32_t n_ids, float* out, void* stream);

/* Reset (legged_robot.py:227-290 reset_idx and the helpers it calls). */
typedef struct RlResetCfg {
  int32_t num_envs;
  int32_t n_terms;                  /* rows 0..n_terms-1 (+ termination, total) are reduced and zeroed */
  int32_t has_termination;
  int32_t custom_origins;           /* mesh_type in heightfield/trimesh (:1389-1404) */
  int32_t terrain_curriculum;       /* cfg.terrain.curriculum and init_done (:800-803) */
  int32_t max_terrain_level;        /* cfg.terrain.num_rows (:1401) */
  int32_t num_terrain_cols;
  float env_length_half;            /* terrain.env_length / 2 (:806) */
  float episode_length_s_half;      /* env.episode_length_s * 0.5 (:809) */
  float base_init_state[13];        /* :1209-1210 */
  float x_init_range, y_init_range; /* passed as (lower, upper) - quirk (:727-729) */
  float x_init_offset, y_init_offset;
  float default_dof_pos[RL_NUM_DOF];
  int32_t randomize_motor_strength, randomize_Kp_factor, randomize_Kd_factor;
  float motor_strength_lo_span[2], Kp_factor_lo_span[2], Kd_factor_lo_span[2];
} RlResetCfg;

typedef struct RlResetBuffers {
  const uint8_t* mask;       /* [N] nonzero = reset this env; or NULL when ids given */
  const int64_t* ids;        /* [n_ids] env ids; or NULL when mask given */
  int32_t n_ids;
  float* root_states;        /* [N,13] */
  float* dof_state;          /* [N,12,2] */
  float* env_origins;        /* [N,3] */
  int64_t* terrain_levels;   /* [N] */
  const int64_t* terrain_types; /* [N] */
  const float* terrain_origins; /* [rows, cols, 3] */
  const float* commands;     /* [N,4] */
  float* last_actions;       /* [12][N] */
  float* last_dof_vel;       /* [12][N] */
  float* feet_air_time;      /* [4][N] */
  int64_t* episode_length_buf; /* [N] */
  uint8_t* reset_buf;        /* [N] */
  float* Kp_factors, *Kd_factors, *motor_strengths; /* [12][N] */
  float* episode_sums;       /* [RL_EPISODE_ROWS][N]: live rows summed over the reset envs then zeroed */
  double* episode_sum_out;   /* [RL_EPISODE_ROWS+1]: per-row sum over the reset envs; last = count */
  float* obs_history;        /* [N,H] or NULL: rows zeroed (history_wrapper.py:34) */
  int32_t obs_history_len;
  /* injected uniforms (parity) or NULL => Philox */
  const float* dr_u;         /* [3][N] */
  const float* init_u;       /* [2][N] xy init draw (:727) */
  const float* level_u;      /* [N] randint draw for solved-last-level envs (:813) */
} RlResetBuffers;

int rl_env_reset(const RlResetCfg* cfg_host, const RlResetBuffers* bufs_host, uint64_t seed,
                 uint64_t step, void* stream);

/* Grid Adaptive Curriculum (legged_robot.py:595-626 _resample_commands;
 * curriculum.py:55-68 sample, :102-119 RewardThresholdCurriculum.update). */
typedef struct RlGacCfg {
  int32_t num_envs;
  int32_t n_bins;               /* nx*ny*nz = 5202 */
  int32_t dims[3];              /* 51, 2, 51 */
  double bin_size[3];           /* arr[1]-arr[0] per axis (curriculum.py:30) */
  float lin_threshold;          /* float32(forward_curriculum_threshold * scale_lin) (:606) */
  float ang_threshold;
  float ep_len;                 /* min(max_episode_length, int(resampling_time/dt)) (:602-603) */
  int32_t lin_slot, ang_slot;   /* command_sums rows of tracking_lin_vel / tracking_ang_vel */
  int32_t n_command_sums;       /* rows of command_sums zeroed for resampled envs (:625): RL_COMMAND_ROWS */
  int32_t num_train_envs;       /* only train envs feed the update (:612) */
} RlGacCfg;

typedef struct RlGacBuffers {
  const uint8_t* mask;          /* [N] or NULL */
  const int64_t* ids;           /* [n_ids] or NULL */
  int32_t n_ids;
  double* weights;              /* [n_bins] float64 (curriculum.py:49) */
  const double* centers;        /* [dims0+dims1+dims2] np.linspace values per axis (curriculum.py:28) */
  const int32_t* nbr_lo;        /* [dims0+dims1+dims2] inclusive neighbour index range per axis index, */
  const int32_t* nbr_hi;        /*   precomputed on the host in float64 exactly as get_local_bins (:102-108) */
  int32_t* hit_count;           /* [n_bins] workspace: neighbourhood incidence count */
  int32_t* own_flag;            /* [n_bins] workspace: bin is the own bin of a successful env */
  double* cdf;                  /* [n_bins] workspace: normalised inclusive prefix sum */
  int64_t* env_command_bins;    /* [N] */
  float* commands;              /* [N,4] */
  float* command_sums;          /* [RL_COMMAND_ROWS][N] */
  const double* u_bin;          /* [N] injected uniform for the categorical draw, or NULL */
  const double* u_cell;         /* [N,3] injected uniforms for the in-cell draw, or NULL */
} RlGacBuffers;

/* phase 1: success test + incidence scatter (int32, order independent) */
int rl_gac_scatter(const RlGacCfg* cfg_host, const RlGacBuffers* bufs_host, void* stream);
/* phase 2: saturating weight update (k times w<-min(1,w+0.2)), cdf, sampling, command write.
 * Between the two phases a multi-GPU caller all-reduces hit_count/own_flag. */
int rl_gac_update_sample(const RlGacCfg* cfg_host, const RlGacBuffers* bufs_host, uint64_t seed,
                         uint64_t step, void* stream);

/* rollout_storage.py:76-90 RolloutStorage.compute_returns.
 *  rewards, values, returns, advantages: [T,N] fp32; dones: [T,N] uint8;
 *  last_values: [N].  workspace: >= rl_gae_workspace_bytes(N) bytes.
 *  Phase A scans time in reverse per env and accumulates sum / sum-of-squares of
 *  the raw advantages in double; phase B normalises with the UNBIASED std (:90).
 *  `stats_out` (3 doubles: sum, sumsq, count) is filled between the phases so a
 *  multi-GPU caller can all-reduce it (call rl_gae_scan, reduce, rl_gae_normalize). */
int64_t rl_gae_workspace_bytes(int32_t num_envs);
int rl_gae_scan(const float* rewards, const float* values, const uint8_t* dones,
                const float* last_values, float* returns, float* advantages, int32_t T,
                int32_t N, float gamma, float lam, void* workspace, double* stats_out,
                void* stream);
int rl_gae_normalize(float* advantages, int32_t T, int32_t N, const double* stats, void* stream);
/* both phases back to back (single GPU) */
int rl_gae(const float* rewards, const float* values, const uint8_t* dones,
           const float* last_values, float* returns, float* advantages, int32_t T, int32_t N,
           float gamma, float lam, void* workspace, void* stream);

/* Dense contraction of the learner on the tcgen05 tensor cores (actor_critic.py:38-100 nn.Linear
 * layers; forward and the autograd backward of ppo.py:102-168).  bf16 operands, fp32 accumulation
 * in tensor memory, fused epilogue:
 *   transposed = 0:  C[M,N] = A[M,K] * B[N,K]^T   (forward: B = weight [out,in]; dgrad: B = weight^T)
 *   transposed = 1:  C[M,N] = A[K,M]^T * B[K,N]   (wgrad: A = dY [batch,out], B = X [batch,in])
 * epilogue: 0 fp32 store, 1 fp32 atomic add (split_k > 1), 2 bf16 elu(acc + bias), 3 fp32 acc + bias,
 *           4 bf16 acc * elu'(aux) with aux the bf16 ELU OUTPUT of the layer being differentiated,
 *           5 bf16 store, 6 bf16 acc + bias.
 * db (transposed form only, may be NULL): db[m] += sum_k A[k,m], the bias gradient, from an extra
 * ones-vector MMA.  All pitches in elements; bf16 operands need 16 B aligned bases and pitches that
 * are multiples of 8. */
/* one-time per-device set-up of the GEMM kernels (shared-memory limits, tensor-map encoder); required
 * before GEMM launches are captured into a CUDA graph, harmless otherwise */
int rl_gemm_init(void);
int rl_gemm_bf16(const void* A, const void* B, void* C, const float* bias, const void* aux, float* db,
                 int32_t M, int32_t N, int32_t K, int32_t lda, int32_t ldb, int32_t ldc, int32_t ld_aux,
                 int32_t transposed, int32_t epilogue, int32_t split_k, void* stream);

/* Weight / bias gradients of several layers in one launch (the autograd wgrad behind ppo.py:146-148,166):
 * dW[M,N] += dY[K,M]^T X[K,N] (fp32 atomics, split-K over the K batch rows), db[m] += sum_k dY[k,m].
 * dY, X: row-major bf16 with pitches ld_dy / ld_x (multiples of 8, 16 B aligned bases); dW fp32 pitch ld_dw. */
typedef struct RlWgradProblem {
  const void* dY;
  const void* X;
  float* dW;
  float* db;                        /* may be NULL */
  int32_t M, N, K;
  int32_t ld_dy, ld_x, ld_dw;
  int32_t split_k;                  /* k ranges per output tile; 0 = let the library choose */
  int32_t reserved;
} RlWgradProblem;
int rl_wgrad_grouped(const RlWgradProblem* problems_host, int32_t n, void* stream);

/* ---- Fused MLP chains on the tensor cores ----------------------------------------------------
 * The learner's networks (actor_critic.py:38-100: encoder 18-256-128-18, actor / critic 60-512-256-128-{12,1},
 * adaptation module 630-256-32-18) are chains of Linear(+ELU) layers.  rl_gemm_bf16 runs one layer per
 * launch and round-trips every activation through HBM; a CHAIN runs a whole forward (or the dgrad half of
 * the backward) of one or more networks for a 128-row tile inside one persistent CTA:
 *   - activations live in shared memory as 16 KB "boxes" ([128 rows x 64 bf16 columns], the 128 B-swizzled
 *     K-major layout TMA writes and tcgen05.mma reads), weights stream from L2 through a ring of 16 KB
 *     stages, accumulators live in the 512 tensor-memory columns;
 *   - three warp roles execute three host-built op lists in order, synchronised only by mbarriers:
 *       LOAD ops (1 thread each of two issuers): one TMA box each (input rows, weight blocks, saved activations for ELU');
 *       MMA ops  (1 thread each of two issuers): up to four tcgen05.mma K16 steps of [128 x n] += A box * B box^T;
 *       EPI ops  (2..4 x 4 warps): tcgen05.ld of <= 64 accumulator columns -> bias / ELU / ELU' -> bf16 box for
 *                           the next layer (and a TMA store of the box for the backward / wgrad) or fp32 rows;
 *                           two to four workers (the launch sizes the CTA by the highest worker index the program uses),
 *                           each runs the ops tagged with its index, in order.
 *   The schedule (which layer's k-block meets which box when) is data: the op lists.  They are built and
 *   checked (deadlock freedom, buffer hazards, numerics on an emulator) on the host:
 *   rapid_locomotion_rl_b200/ppo/chain.py.
 * A wait is encoded in 16 bits: bits 0-7 barrier id (0xFF = none), bit 8 = parity to wait for in the
 * CTA's first tile, bit 9 = 1 if that parity flips with every further tile of the persistent loop. */
#define RL_CHAIN_MAX_TENSORS 32
#define RL_CHAIN_MAX_BARRIERS 80
#define RL_CHAIN_MAX_UNITS 14                    /* 16 KB shared-memory units (boxes + ring stages) */
#define RL_CHAIN_MAX_OUTPUTS 4
#define RL_CHAIN_NONE 255

typedef struct RlChainTensor {      /* row-major bf16 [rows, cols], `ld` elements between rows */
  void* base;
  int64_t rows;
  int32_t cols, ld;
  int32_t box_rows;                 /* TMA box = [box_rows, 64 columns]; out-of-range elements read as 0 */
  int32_t reserved;
} RlChainTensor;

typedef struct RlChainLoadOp {
  uint16_t wait;                    /* destination unit free */
  uint8_t full_bar;                 /* completes (with the byte count) when the box has landed */
  uint8_t tensor;
  uint32_t smem_off;                /* destination, multiple of 1024 */
  int32_t col0, row0;               /* box origin (elements); row0 is relative to the tile when tile_rows */
  uint32_t expect_bytes;            /* box bytes (box_rows * 128) */
  uint8_t tile_rows;
  uint8_t issuer;                   /* 0 / 1: which of the two LOAD warps issues it (the list must be sorted by issuer) */
  uint8_t pad1, pad2;
  uint32_t pad3, pad4;
} RlChainLoadOp;

typedef struct RlChainMmaOp {
  uint32_t a_off, b_off;            /* A: [128 x 64] box, B: [n x 64] box (both K-major, 128 B swizzle) */
  uint16_t n;                       /* 16..256, multiple of 16 */
  uint16_t tmem_col;
  uint8_t k_steps;                  /* 1..4 K16 steps of the 64-wide k-block */
  uint8_t accumulate;               /* 0: the first step overwrites the accumulator */
  uint16_t wait0, wait1, wait2;
  uint8_t commit0, commit1, commit2;   /* mbarriers that tcgen05.commit arrives on after these MMAs */
  uint8_t issuer;                   /* 0 / 1: which of the two MMA warps issues it (sorted list; MMAs into one accumulator
                                     * must share an issuer: tcgen05.mma is ordered only within a thread) */
  uint16_t wait3;                   /* a fourth wait (0xFF in the low byte = none); zero-initialised structs must set it */
  uint16_t pad1;
  uint32_t pad2;
} RlChainMmaOp;

enum RlChainEpiMode {
  RL_CHAIN_EPI_BIAS_ELU = 0,        /* box = bf16(elu(acc + bias)) */
  RL_CHAIN_EPI_BIAS = 1,            /* box = bf16(acc + bias) */
  RL_CHAIN_EPI_BIAS_F32 = 2,        /* out[row, 0:ncols] = acc + bias (fp32 rows in global memory) */
  RL_CHAIN_EPI_DELU = 3,            /* box = bf16(acc * elu'(aux)), aux = saved ELU output box */
  RL_CHAIN_EPI_PLAIN = 4,           /* box = bf16(acc) */
  /* the tanh networks of the reference's high_level_policy learner (high_level_policy/ppo/actor_critic.py:15, :196-213) */
  RL_CHAIN_EPI_BIAS_TANH = 5,       /* box = bf16(tanh(acc + bias)) */
  RL_CHAIN_EPI_DTANH = 6            /* box = bf16(acc * (1 - aux^2)), aux = saved tanh output box */
};

typedef struct RlChainEpiOp {
  uint16_t wait_acc, wait_dst, wait_aux;
  uint8_t arrive_acc_free;          /* after the accumulator columns are in registers (count 4: one per warp) */
  uint8_t arrive_dst_ready;         /* after the box is written (count 4) */
  uint8_t release_aux;              /* after every thread has read the aux box (count 1: the last of the worker's warps arrives) */
  uint8_t mode;
  uint8_t ncols;                    /* accumulator columns handled: 1..64 */
  uint8_t dst_col0;                 /* first box column written (0 except when merging into a loaded box) */
  uint16_t tmem_col;
  uint8_t store_tensor;             /* TMA store of the whole destination box after writing, or NONE */
  int8_t store_wait_pending;        /* >= 0: cp.async.bulk.wait_group.read <n> before writing (box reuse) */
  uint8_t release_after_store;      /* barrier to arrive on once the store has read the box, or NONE */
  uint8_t out_id;                   /* RL_CHAIN_EPI_BIAS_F32: index into outputs[] */
  uint16_t out_ld;
  uint32_t bias_off;                /* float offset into `params` of this op's first column bias */
  uint32_t dst_off, aux_off;
  int32_t store_col0;
  uint8_t worker;                   /* which of the (up to four) epilogue warp groups executes this op */
  uint8_t padb0, padb1, padb2;
  uint32_t delay_ns;                /* sleep this long before the op (0: none): de-phases workers that would otherwise run
                                     * their MUFU-bound ELU phases in lock step */
  uint32_t pad2;
} RlChainEpiOp;

typedef struct RlChainDesc {
  RlChainTensor tensors[RL_CHAIN_MAX_TENSORS];
  int32_t n_tensors;
  int32_t n_units;                  /* 16 KB shared-memory units used (boxes + stages) */
  int32_t n_barriers;
  int32_t n_loads, n_mmas, n_epis;
  uint8_t barrier_count[RL_CHAIN_MAX_BARRIERS];
  const RlChainLoadOp* loads_host;  /* HOST arrays; copied to the device by rl_chain_create */
  const RlChainMmaOp* mmas_host;
  const RlChainEpiOp* epis_host;
  const float* params;              /* DEVICE: flat fp32 parameter buffer the bias offsets index */
  float* outputs[RL_CHAIN_MAX_OUTPUTS];   /* DEVICE fp32 outputs */
} RlChainDesc;

/* Compiles a chain: encodes the tensor maps, uploads the op lists.  `*handle` is owned by the library
 * until rl_chain_destroy.  The tensors' base pointers are baked in (workspace buffers must stay put). */
int rl_chain_create(const RlChainDesc* desc_host, void** handle);
/* Runs the chain over rows [0, rows) (tiles of 128; rows may be any value up to the tensors' extent). */
int rl_chain_run(void* handle, int32_t rows, void* stream);
/* the same for the row tiles [tile_begin, tile_end) of the batch only (128 rows per tile): lets the caller pipeline
 * chunks of a batch through dependent passes on different streams */
int rl_chain_run_tiles(void* handle, int32_t rows, int32_t tile_begin, int32_t tile_end, void* stream);
int rl_chain_destroy(void* handle);
/* PPO loss fused into a teacher-forward chain (ppo.py:110-144): the epilogue thread that writes row r of the critic output
 * (outputs[value_out]; the same thread wrote row r of outputs[mean_out] earlier in its op list - rl_chain_set_ppo_loss
 * checks that) evaluates the row's clipped surrogate / clipped value loss / KL and writes the gradients w.r.t. the network
 * outputs exactly as rl_ppo_loss does: dmean [rows,16] and dvalue [rows,8] bf16, dstd[12] / stats[0..2] / kl_slot
 * accumulated atomically (one set of atomics per warp).  Removes the loss launch between the forward and the backward chain
 * of an update.  Row indices are those of the whole batch (pointers are NOT offset by the first tile of a
 * rl_chain_run_tiles call).  loss_host = NULL switches it off; the setting applies to the launches that follow. */
typedef struct RlChainPpoLoss {
  const float* Lrow;                /* [rows, 40] per-row loss inputs (rl_ppo_gather) */
  const float* std;                 /* [12] */
  void* dmean;                      /* bf16 [rows, 16] */
  void* dvalue;                     /* bf16 [rows, 8] */
  float* dstd;                      /* [12] */
  double* stats;                    /* [4] */
  float* kl_slot;                   /* or NULL */
  float clip, value_coef, entropy_coef, inv_global_B;
  int32_t use_clipped_value;
  int32_t mean_out, value_out;      /* indices into the chain's outputs[] */
  int32_t pad;
} RlChainPpoLoss;
int rl_chain_set_ppo_loss(void* handle, const RlChainPpoLoss* loss_host);
/* Profiling aid: record per-op clock64 stamps of CTA 0 during its tile iteration `tile_iteration` (< 0: off):
 * 1 stamp per LOAD op, 2 per MMA op (waits passed, commits issued), 5 per EPI op (start, accumulator ready,
 * registers loaded, elementwise done, end), in that order.  rl_chain_read_trace copies them to the host
 * (synchronising) and returns how many stamps it wrote. */
int rl_chain_trace(void* handle, int32_t tile_iteration);
int64_t rl_chain_read_trace(void* handle, uint64_t* out_host, int64_t capacity);

/* ---- PPO update support (mini_gym_learn/ppo/ppo.py:94-178) ---------------------------------- */

/* rollout_storage.py:121-137: the twelve per-minibatch advanced-index gathers, fused with the bf16
 * staging of the GEMM inputs.  Flat storage rows are [T*N, dim] fp32; idx [B] int64.
 *   Xp  [B, ldp]  bf16: privileged obs (zero padded)          encoder input
 *   Xac [B, ldac] bf16: obs in columns [0, obs_dim); columns [obs_dim, obs_dim+18) are left for the
 *                       encoder latent; the rest zero          actor / critic input
 *   Xh  [B, ldh]  bf16: observation history (may be NULL)      adaptation-module input
 *   Lrow [B, 40] fp32: actions 12 | old mu 12 | old sigma 12 | old log-prob | advantage | return | old value */
int rl_ppo_gather(const float* obs, const float* priv, const float* hist, const float* actions,
                  const float* values, const float* returns, const float* logp, const float* adv,
                  const float* mu, const float* sigma, const int64_t* idx, int32_t B, int32_t obs_dim,
                  int32_t priv_dim, int32_t hist_dim, void* Xp, int32_t ldp, void* Xac, int32_t ldac, void* Xh,
                  int32_t ldh, float* Lrow, void* stream);
/* rollout_storage.py:124 (`obs_history_batch = observation_histories[batch_idx]`) alone: Xh [B, ldh] bf16.  The
 * history is 86 % of a minibatch's gathered bytes and only the adaptation module reads it, so PPO.update issues it
 * on the stream that runs the adaptation forward, off the policy path (rl_ppo_gather is then called with Xh = NULL). */
int rl_ppo_gather_history(const float* hist, const int64_t* idx, int32_t B, int32_t hist_dim, void* Xh, int32_t ldh,
                          void* stream);
/* fp32 [rows, cols] -> bf16 dst[:, dst_col0 : dst_col0 + pad_to], zero padded beyond cols */
int rl_cast_bf16(const float* src, int32_t ld_src, void* dst, int32_t ld_dst, int32_t rows, int32_t cols,
                 int32_t dst_col0, int32_t pad_to, void* stream);
/* ppo.py:110-144 (KL, clipped surrogate, clipped value loss, entropy) and :157-164 (adaptation MSE)
 * with the analytic gradients w.r.t. the network outputs written as bf16 GEMM operands:
 * dmean [B,16], dvalue [B,8], dpred [B,24]; dstd[12] and stats[4] (sum surrogate, sum value loss,
 * sum kl, sum adaptation squared error; double) are accumulated atomically.  inv_global_B =
 * 1 / (B * world_size).  kl_slot (optional): an fp32 word that also receives the KL sum - with several GPUs it is
 * the word behind the flat gradient, so the KL that drives the learning rate travels in the gradient all-reduce. */
int rl_ppo_loss(const float* mean, const float* value, const float* pred, const void* Xac, int32_t ldac,
                int32_t lat_off, const float* Lrow, const float* std, int32_t B, float clip, float value_coef,
                float entropy_coef, int32_t use_clipped_value, float inv_global_B, void* dmean, void* dvalue,
                void* dpred, float* dstd, double* stats, float* kl_slot, void* stream);
/* ppo.py:157-164 alone: mse(adaptation_module(hist), encoder(priv).detach()) and its gradient dpred
 * [B,24] bf16; adds the squared error to stats[3].  Called after the policy optimiser step, as the
 * reference computes the target with the already updated encoder. */
int rl_adapt_loss(const float* pred, const void* Xac, int32_t ldac, int32_t lat_off, int32_t B,
                  float inv_global_B, void* dpred, double* stats, void* stream);
/* ppo.py:116-124 + :149: gradient 2-norm over `n` floats -> clip coefficient; kl mean -> adaptive
 * learning rate, all on the device.  ctrl[0] = lr (in/out), ctrl[1] = clip coefficient, ctrl[2] = kl.
 * workspace: 16 zeroed bytes.  kl_slot (optional): the KL sum is read from this fp32 word (the all-reduced one)
 * instead of stats[2].  loss_acc (optional, double[4]): stats[0..2] are added to loss_acc[0..2] and zeroed
 * (kl_slot too) - the per-update loss bookkeeping of ppo.py:152-153 without extra launches. */
int rl_grad_finalize(const float* grad, int64_t n, double* stats, float* ctrl, void* workspace,
                     double global_B, float desired_kl, float max_grad_norm, int32_t adaptive, double* loss_acc,
                     float* kl_slot, void* stream);
/* same, from an already reduced squared gradient norm (device double) */
int rl_grad_finalize_from_norm(const double* norm2, double* stats, float* ctrl, double global_B,
                               float desired_kl, float max_grad_norm, int32_t adaptive, double* loss_acc, float* kl_slot,
                               void* stream);

/* ---- multi-GPU: gradient all-reduce over NVLink peer memory, fused with the gradient-norm reduction and
 * zero_grad (the sum of the env shards' gradients that precedes clip_grad_norm_ / optimizer.step,
 * ppo.py:146-150).  Every rank owns a gradient buffer, a staging buffer of the same size and a 256 B flag block,
 * all mapped by every other rank (CUDA IPC); `local_ws` is 16 zeroed bytes of ordinary device memory
 * (a double accumulator and the call counter).  One launch:
 * wait until every rank's gradient is complete -> rank r sums slice r of all gradients (rank order, bit-identical
 * everywhere) into its staging buffer and accumulates its squared norm -> wait -> gather all slices into `out`,
 * total norm into `norm2_out`, zero the own gradient segment.  `step` is the call ordinal (> 0, equal on all
 * ranks, strictly increasing), or 0 to use the device-side counter kept at local_ws + 8 (CUDA-graph replay).  offset and n are multiples of 4 floats; the norm covers the first norm_n floats. */
#define RL_PEER_MAX_RANKS 16
typedef struct RlPeerComm {
  void* grad[RL_PEER_MAX_RANKS];    /* rank p's gradient buffer as mapped in THIS process */
  void* stage[RL_PEER_MAX_RANKS];   /* rank p's staging buffer */
  void* flags[RL_PEER_MAX_RANKS];   /* rank p's flag block (zero initialised) */
  void* local_ws;
  int32_t world, rank;
} RlPeerComm;
int rl_enable_peer_access(int32_t peer_device);   /* current device -> peer_device loads / stores */
int rl_peer_allreduce(const RlPeerComm* comm_host, int64_t offset, int64_t n, int64_t norm_n, float* out,
                      double* norm2_out, uint32_t step, void* stream);

/* torch.optim.Adam step (ppo.py:44-46,150,168) fused with gradient scaling and zero_grad.
 * use_ctrl: lr = ctrl[0], grad scaled by ctrl[1]; else lr_fixed.  grad_scale: extra factor. */
int rl_adam(float* p, float* g, float* m, float* v, int64_t n, const float* ctrl, float lr_fixed,
            int32_t use_ctrl, float beta1, float beta2, float eps, int32_t step, float grad_scale,
            int32_t* step_dev, void* stream);
/* step_dev (may be NULL): device int32[2] {optimiser step count, CTA ticket}.  When given, the bias
 * corrections use step_dev[0] + 1 and the kernel advances the counter itself, so the launch can be captured
 * in a CUDA graph and replayed; `step` is then ignored. */
/* rl_adam and rl_refresh_shadows in one launch: the parameters [p, p + n) are stepped and every element that lies
 * inside one of the `n_layers` weight matrices (w_start[i]: its first element relative to p, [out_dim[i], in_dim[i]]
 * row-major) is also written to that layer's bf16 operands wb[i] ([out, ld_wb]) and wbt[i] ([in, ld_wbt], transposed).
 * Pad columns of the operands are left as they are (zero since the first rl_refresh_shadows). */
int rl_adam_shadows(float* p, float* g, float* m, float* v, int64_t n, const float* ctrl, float lr_fixed,
                    int32_t use_ctrl, float beta1, float beta2, float eps, int32_t step, float grad_scale,
                    int32_t* step_dev, const int64_t* w_start, void* const* wb, void* const* wbt,
                    const int32_t* out_dim, const int32_t* in_dim, const int32_t* ld_wb, const int32_t* ld_wbt,
                    int32_t n_layers, void* stream);
/* bf16 shadow copies of the fp32 master weights: wb [out, ld_wb] and its transpose wbt [in, ld_wbt]
 * (host arrays of n_layers device pointers / dims) */
int rl_refresh_shadows(const void* const* w, void* const* wb, void* const* wbt, const int32_t* out_dim,
                       const int32_t* in_dim, const int32_t* ld_wb, const int32_t* ld_wbt, int32_t n_layers,
                       void* stream);
/* actor_critic.py:137-147: a = mu + std * N(0,1) (Philox + Box-Muller, or injected normals [N,12]),
 * summed Normal log-prob, and the mu / sigma rows PPO.act stores */
int rl_policy_sample(const float* mean, const float* std, int32_t N, uint64_t seed, uint64_t step,
                     const float* inj_normal, float* actions, float* logp, float* mu_out, float* sigma_out,
                     void* stream);

/* rollout_storage.py:54-71 RolloutStorage.add_transitions: one launch writes the transition of every env into
 * the [t] slices of the storage (the reference: eleven copy_ kernels per step).  Sources are [N, dim] rows with
 * the given pitches (the observation history is a strided view of the ring buffer); destinations are dense.
 * obs / priv / hist may be NULL (with their destinations): those rows were stored when the action was taken. */
typedef struct RlStorageAdd {
  const float* obs;
  const float* priv;
  const float* hist;
  const float* actions;
  const float* mu;
  const float* sigma;
  const float* rewards;
  const float* values;
  const float* logp;
  const float* bins;
  const uint8_t* dones;
  float* dst_obs;
  float* dst_priv;
  float* dst_hist;
  float* dst_actions;
  float* dst_mu;
  float* dst_sigma;
  float* dst_rewards;
  float* dst_values;
  float* dst_logp;
  float* dst_bins;
  uint8_t* dst_dones;
  int64_t ld_obs, ld_priv, ld_hist;
  int32_t N, obs_dim, priv_dim, hist_dim, act_dim;
  int32_t reserved;
} RlStorageAdd;
int rl_storage_add(const RlStorageAdd* q_host, void* stream);

/* history_wrapper.py:23 - append obs to a 2H-slot ring so that the last H steps are
 * always one contiguous row span: hist [N, 2*H*num_obs], writes slots k and k+H. */
int rl_history_push(float* hist, const float* obs, int32_t N, int32_t num_obs, int32_t H,
                    int32_t slot, void* stream);

/* ---- rollout-step glue of Runner.learn (mini_gym_learn/ppo/__init__.py:126-141), two launches per step ----
 * rl_rollout_boundary runs BETWEEN env step t and the policy pass of step t+1, one warp per env:
 *   post part (do_post): closes transition t - dst_rewards[n] = rew[n] (+ gamma * values_prev[n] where time_outs[n],
 *     ppo.py:81-83), dst_dones, dst_bins (rollout_storage.py:63-70) - and pushes the new observation into the history
 *     ring at `push_slot` (history_wrapper.py:23; ring layout as rl_history_push);
 *   pre part (do_pre): opens transition t+1 - dst_obs / dst_priv / dst_hist (rollout_storage.py:57-60; the history row is
 *     the H-slot span starting at ring slot `hist_slot`, whose last slot is the observation just pushed) and the bf16
 *     staging of the policy inputs: Xac[n, 0:obs_dim] = obs, Xp[n, 0:ld_xp] = priv zero padded. */
typedef struct RlRolloutBoundary {
  const float* obs;          /* [N, obs_dim] observation after env step t */
  const float* priv;         /* [N, priv_dim] */
  float* ring;               /* [N, 2*H*obs_dim] */
  const float* rew;          /* [N] */
  const uint8_t* dones;      /* [N] bool */
  const uint8_t* time_outs;  /* [N] bool or NULL */
  const float* values_prev;  /* [N] value estimates of transition t (needed with time_outs) */
  const float* bins;         /* [N] or NULL (zeros) */
  float* dst_rewards;        /* storage.rewards[t] [N] */
  uint8_t* dst_dones;        /* storage.dones[t] [N] */
  float* dst_bins;           /* storage.env_bins[t] [N] */
  float* dst_obs;            /* storage.observations[t+1] [N, obs_dim] */
  float* dst_priv;           /* storage.privileged_observations[t+1] [N, priv_dim] */
  float* dst_hist;           /* storage.observation_histories[t+1] [N, H*obs_dim] */
  void* Xac;                 /* bf16 [N, ld_xac] */
  void* Xp;                  /* bf16 [N, ld_xp] */
  float gamma;
  int32_t N, obs_dim, priv_dim, H;
  int32_t push_slot, hist_slot;
  int32_t ld_xac, ld_xp;
  int32_t do_post, do_pre;
  int32_t reserved;
} RlRolloutBoundary;
int rl_rollout_boundary(const RlRolloutBoundary* q_host, void* stream);

/* rl_rollout_act runs AFTER the policy pass, one thread per env: a = mean + std * N(0,1) (actor_critic.py:142-147;
 * Philox keyed by (seed, env, step + step_state[0]) or the injected normals), its log-probability, and the
 * transition's slices: actions / mu / sigma / log-prob / value (rollout_storage.py:61-69) next to `actions_out`, the
 * buffer the env step reads.  step_state (optional, uint64[2]): device-side step counter advanced once per launch, so
 * the launch replays inside a CUDA graph.  dst_* may all be NULL (plain ActorCritic.act). */
typedef struct RlRolloutAct {
  const float* mean;         /* [N, 12] */
  const float* value;        /* [N] */
  const float* std;          /* [12] */
  const float* inj_normal;   /* [N, 12] or NULL */
  float* actions_out;        /* [N, 12] */
  float* logp_out;           /* [N] or NULL */
  float* dst_actions;        /* storage.actions[t] or NULL */
  float* dst_mu;
  float* dst_sigma;
  float* dst_logp;
  float* dst_values;
  uint64_t* step_state;      /* [2] or NULL */
  uint64_t seed, step;
  int32_t N;
  int32_t reserved;
} RlRolloutAct;
int rl_rollout_act(const RlRolloutAct* q_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RL_B200_H */
