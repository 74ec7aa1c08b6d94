"""LeggedRobot env core on the fused sm_100a kernels.

Drop-in for mini_gym/envs/base/legged_robot.py (`LeggedRobot` :21, BaseTask buffers
base_task.py:57-63): same constructor arguments, `step(actions) -> (obs, privileged_obs, rew,
reset, extras)` (:106-137), `reset()` (base_task.py:103-108), `reset_idx(env_ids)` (:227-290),
`_resample_commands(env_ids)` (:595-626) and the public attributes the learner / scripts read
(SURVEY.md 8b).  Everything between the simulator tensors and the returned buffers is one CUDA
launch per call; there is no eager fallback - without the native library construction fails.

Layout notes: state this class owns is stored SoA `[K, N]` for coalescing; attributes such as
`last_actions`, `motor_strengths`, `base_lin_vel` are exposed as `[N, K]` transposed views of that
storage, so indexing and in-place writes behave as in the reference.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from ..config import (COMMAND_SUM_EXTRAS, TerrainInfo, freeze_env_cfg, lo_span, public_vars, f32)
from ..robots import robot_for_asset
from ..sim import SyntheticSim
from .curriculum import RewardThresholdCurriculum


class LeggedRobot:
    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True,
                 eval_cfg=None, initial_dynamics_dict=None, sim=None, terrain=None, seed=0,
                 gac_rng="philox", upstream_order=False):
        # train / eval split (legged_robot.py:37-46, base_task.py:43-49): `eval_cfg.env.num_envs` evaluation envs follow the
        # `cfg.env.num_envs` training envs; the functions the reference calls through _call_train_eval :456-469 see the
        # evaluation Cfg for that range (teleport / push / DOF re-draw inside the step; curricula, DOF / root reset in reset_idx;
        # env origins and rigid-body properties at construction), everything else the training Cfg
        self.cfg = cfg
        self.eval_cfg = eval_cfg
        self.sim_params = sim_params
        self.physics_engine = physics_engine
        self.sim_device = sim_device
        self.headless = headless
        self.device = torch.device(sim_device)
        if self.device.type != "cuda":
            raise _lib.RlError("LeggedRobot needs a CUDA device (got %r): the hot path has no CPU fallback" % (sim_device,))
        self._lib = _lib.lib()
        self.seed = int(seed)
        self.gac_rng = gac_rng
        self.initial_dynamics_dict = initial_dynamics_dict
        self.init_done = False

        # ---- _parse_cfg (:1417-1429) + asset facts + terrain ----
        robot = robot_for_asset(cfg.asset.file)
        if terrain is None:
            terrain = TerrainInfo(cfg.terrain, eval_terrain=None if eval_cfg is None else eval_cfg.terrain)
        self.terrain = terrain
        sim_dt = None if sim_params is None else getattr(sim_params, "dt", None)
        # reward plugin surface (legged_robot.py:1074-1093): a scale whose name is not one of the fused terms - or whose
        # `_reward_<name>` a subclass overrides - is served by that Python method after the fused launch
        custom = [n for n, sc in public_vars(cfg.rewards.scales).items() if sc != 0 and n != "termination" and
                  self._is_custom_reward(n)]
        # upstream_order: this fork commented the reset / time-out / resampling calls out of the step
        # (legged_robot.py:177, :197-198, :246, :581); True restores the upstream legged_gym order: resample commands every
        # resampling_time -> terminations (+ time-outs) -> rewards -> reset_idx of the terminated envs (+ their commands).
        # Default False = the fork as written = what the goldens pin.
        self.upstream_order = bool(upstream_order)
        p = self.params = freeze_env_cfg(cfg, robot, terrain, sim_dt, custom_reward_names=custom,
                                         upstream_order=self.upstream_order, eval_cfg=eval_cfg)
        if cfg.terrain.mesh_type not in ("heightfield", "trimesh"):
            cfg.terrain.curriculum = False
        self.dt = p.dt_double
        self.obs_scales = cfg.normalization.obs_scales
        self.reward_scales = dict(p.reward_scales)
        self.reward_names = list(p.reward_names)
        self.reward_functions = [getattr(self, "_reward_" + n) for n in self.reward_names]
        self._custom_terms = list(p.custom_terms)
        base = LeggedRobot
        self._hook_overrides = [h for h in ("check_termination", "compute_reward", "compute_observations")
                                if getattr(type(self), h) is not getattr(base, h)]
        if type(self)._get_heights is not base._get_heights:
            raise NotImplementedError("_get_heights is overridden: the height samples feed the fused observation / reward "
                                      "code inside the kernel, a Python replacement cannot reach them")
        if "check_termination" in self._hook_overrides and (p.has_termination or "survival" in self.reward_names):
            raise NotImplementedError("check_termination is overridden while the `termination` / `survival` reward terms "
                                      "are enabled: those terms are evaluated inside the fused launch from the built-in "
                                      "termination rule")
        cfg.command_ranges = public_vars(cfg.commands)
        cfg.env.max_episode_length = p.max_episode_length_f
        self.max_episode_length = p.max_episode_length_f
        cfg.domain_rand.push_interval = float(p.push_interval)
        cfg.domain_rand.rand_interval = float(p.rand_interval)
        cfg.env.num_height_points = len(p.height_points)
        if eval_cfg is not None:
            # _parse_cfg(eval_cfg) :1417-1429 (called after the training Cfg's: its episode length is the one left in
            # self.max_episode_length, :43 / :1425)
            if eval_cfg.terrain.mesh_type not in ("heightfield", "trimesh"):
                eval_cfg.terrain.curriculum = False
            eval_cfg.command_ranges = public_vars(eval_cfg.commands)
            eval_cfg.env.max_episode_length = float(np.ceil(eval_cfg.env.episode_length_s / self.dt))
            self.max_episode_length = eval_cfg.env.max_episode_length
            eval_cfg.domain_rand.push_interval = float(np.ceil(eval_cfg.domain_rand.push_interval_s / self.dt))
            eval_cfg.domain_rand.rand_interval = float(np.ceil(eval_cfg.domain_rand.rand_interval_s / self.dt))

        N = self.num_envs = p.num_envs
        self.num_train_envs, self.num_eval_envs = p.num_train_envs, p.num_eval_envs
        self.num_obs, self.num_privileged_obs, self.num_actions = p.num_obs, p.num_privileged_obs, p.num_actions
        self.num_dof = self.num_dofs = robot.num_dof
        self.num_bodies = robot.num_bodies
        self.dof_names = list(robot.dof_names)
        dev = self.device

        def z(*shape, dtype=torch.float):
            return torch.zeros(*shape, dtype=dtype, device=dev)

        # ---- simulator tensors (:939-971); identity actor index tables (one actor per env) ----
        self.sim = sim if sim is not None else SyntheticSim(robot, N, dev)
        self.all_root_states = self.root_states = self.sim.root_states
        self.all_dof_state = self.dof_state = self.sim.dof_state
        self.all_contact_forces = self.sim.contact_forces
        self.all_rigid_body_state = self.rigid_body_state = self.sim.rigid_body_state
        self.dof_pos = self.dof_state.view(N, self.num_dof, 2)[..., 0]
        self.dof_vel = self.dof_state.view(N, self.num_dof, 2)[..., 1]
        self.base_quat = self.root_states[:, 3:7]
        self.contact_forces = self.all_contact_forces.view(N, -1, 3)

        # ---- BaseTask buffers (base_task.py:57-63) ----
        self.obs_buf = z(N, self.num_obs)
        self.privileged_obs_buf = z(N, self.num_privileged_obs)
        self.rew_buf = z(N)
        self._reset_u8 = torch.ones(N, dtype=torch.uint8, device=dev)
        self.reset_buf = self._reset_u8.view(torch.bool)
        self.episode_length_buf = z(N, dtype=torch.long)
        self._time_out_u8 = z(N, dtype=torch.uint8)
        self.time_out_buf = self._time_out_u8.view(torch.bool)
        self.extras = {}

        # ---- persistent state, SoA storage with [N,K] views ----
        # Three packed row blocks (csrc/env_step_rows.cu): a CTA of the fused step fetches its 32-env column of each
        # block with one 2-D TMA copy.  RO = read by the step, RW = read and written, WO = written.
        nd = self.num_dof
        self._ro = z(42, N)       # Kp 0-11 | Kd 12-23 | motor 24-35 | friction 36 | restitution 37 | payload 38 | com 39-41
        self._rw = z(28, N)       # last_actions 0-11 | last_dof_vel 12-23 | feet_air_time 24-27
        self._wo = z(27, N)       # joint_pos_target 0-11 | base_lin_vel | base_ang_vel | projected_gravity | last_root_vel 21-26
        self._last_actions = self._rw[0:12]; self.last_actions = self._last_actions.t()
        self.actions = self.last_actions  # equal after every step (:181)
        self._last_dof_vel = self._rw[12:24]; self.last_dof_vel = self._last_dof_vel.t()
        self._feet_air_time = self._rw[24:28]; self.feet_air_time = self._feet_air_time.t()
        self._joint_pos_target = self._wo[0:12]; self.joint_pos_target = self._joint_pos_target.t()
        self._base_lin_vel = self._wo[12:15]; self.base_lin_vel = self._base_lin_vel.t()
        self._base_ang_vel = self._wo[15:18]; self.base_ang_vel = self._base_ang_vel.t()
        self._projected_gravity = self._wo[18:21]; self.projected_gravity = self._projected_gravity.t()
        self._last_root_vel = self._wo[21:27]; self.last_root_vel = self._last_root_vel.t()
        self._ro[0:36].fill_(1.0)
        self._Kp = self._ro[0:12]; self.Kp_factors = self._Kp.t()
        self._Kd = self._ro[12:24]; self.Kd_factors = self._Kd.t()
        self._motor = self._ro[24:36]; self.motor_strengths = self._motor.t()
        self.default_friction, self.default_restitution = robot.default_friction, robot.default_restitution
        self.friction_coeffs = self._ro[36]; self.friction_coeffs.fill_(self.default_friction)
        self.restitutions = self._ro[37]; self.restitutions.fill_(self.default_restitution)
        self.payloads = self._ro[38]
        self._com = self._ro[39:42]; self.com_displacements = self._com.t()
        self._last_contacts_u8 = z(N, 4, dtype=torch.uint8)
        self.last_contacts = self._last_contacts_u8.view(torch.bool)
        self.torques = z(N, nd)
        self.commands_value = z(N, cfg.commands.num_commands)
        self.commands = torch.zeros_like(self.commands_value)
        self.commands_scale = torch.tensor(p.commands_scale, device=dev)
        self.measured_heights = z(N, p.num_height_points) if p.measure_heights else 0
        D = _lib.DEFINES
        self._episode_sums = z(D["RL_MAX_TERMS"] + 2, N)      # fixed row layout, see rl_b200.h
        self._command_sums = z(D["RL_MAX_TERMS"] + 6, N)
        self._episode_rows = dict(p.sum_rows, total=D["RL_MAX_TERMS"] + 1)
        self._command_rows = dict(p.sum_rows, **{n: D["RL_MAX_TERMS"] + 1 + i for i, n in enumerate(COMMAND_SUM_EXTRAS)})
        self.episode_sums = {n: self._episode_sums[r] for n, r in self._episode_rows.items()}
        self.command_sums = {n: self._command_sums[r] for n, r in self._command_rows.items()}
        for name, _ in self._custom_terms:          # plugin terms keep their accumulators outside the packed rows
            self.episode_sums[name] = z(N)
            self.command_sums[name] = z(N)
        # (:1101-1105: -1 marks "no finished evaluation episode recorded yet"; the total row starts at 0 in the reference)
        self.episode_sums_eval = {n: -1 * torch.ones(N, device=dev) for n in self.episode_sums if n != "total"}
        self.episode_sums_eval["total"] = z(N)
        self._rew_raw = z(N) if self._custom_terms else None
        self._episode_sum_out = z(D["RL_MAX_TERMS"] + 3, dtype=torch.float64)
        self.common_step_counter = 0

        # ---- constants as tensors (callers read them) ----
        self.p_gains = torch.tensor(p.p_gains, device=dev)
        self.d_gains = torch.tensor(p.d_gains, device=dev)
        self.default_dof_pos = torch.tensor(p.default_dof_pos, device=dev).unsqueeze(0)
        self.torque_limits = torch.tensor(p.torque_limits, device=dev)
        self.dof_vel_limits = torch.tensor(p.dof_vel_limits, device=dev)
        self.dof_pos_limits = torch.tensor(list(zip(p.dof_pos_lo, p.dof_pos_hi)), device=dev)
        self.feet_indices = torch.tensor(p.feet_idx, dtype=torch.long, device=dev)
        self.termination_contact_indices = torch.tensor(p.term_idx[:p.n_term_bodies], dtype=torch.long, device=dev)
        self.penalised_contact_indices = torch.tensor(p.pen_idx[:p.n_pen_bodies], dtype=torch.long, device=dev)
        self.noise_scale_vec = torch.from_numpy(p.noise_scale_vec).to(dev)
        self.add_noise = bool(p.add_noise)
        self._height_points_xy = torch.from_numpy(p.height_points).to(dev).contiguous()
        if p.measure_heights:
            pts = torch.zeros(N, len(p.height_points), 3, device=dev)
            pts[:, :, :2] = self._height_points_xy
            self.height_points = pts
        self.height_samples = None
        if terrain.custom:
            self.height_samples = torch.from_numpy(terrain.heightsamples).to(dev).contiguous()
        self.gravity_vec = torch.tensor([0.0, 0.0, -1.0], device=dev).repeat(N, 1)
        self.forward_vec = torch.tensor([1.0, 0.0, 0.0], device=dev).repeat(N, 1)
        base_init = list(cfg.init_state.pos) + list(cfg.init_state.rot) + list(cfg.init_state.lin_vel) + list(cfg.init_state.ang_vel)
        self.base_init_state = torch.tensor(base_init, dtype=torch.float, device=dev)

        # ---- env origins (:1385-1415) ----
        self.env_origins = z(N, 3)
        self.terrain_levels = z(N, dtype=torch.long)
        self.terrain_types = z(N, dtype=torch.long)
        self.terrain_origins_t = None
        self._init_env_origins()

        # ---- domain randomisation initial draw (:1231, :519-541) ----
        if initial_dynamics_dict is not None:
            for k, v in initial_dynamics_dict.items():
                self._set_dynamics(k, v)
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(self.seed)
        self._randomize_rigid_body_props(torch.arange(self.num_train_envs, device=dev), cfg)          # :1231 per range
        if self.num_eval_envs:
            self._randomize_rigid_body_props(torch.arange(self.num_train_envs, N, device=dev), eval_cfg)

        # ---- command curriculum (:1056-1072) ----
        self._init_command_distribution()

        self._inject = {}     # optional injected uniforms for parity tests
        self._bufs = self._make_buffers()
        self._cfg_struct = p.to_struct()
        self._reset_cfg = self._make_reset_cfg(cfg)
        self._reset_cfg_eval = self._make_reset_cfg(eval_cfg) if eval_cfg is not None else None
        self.init_done = True
        self.record_now = self.record_eval_now = False

    # ------------------------------------------------------------------------------------------
    # construction helpers
    # ------------------------------------------------------------------------------------------
    def _set_dynamics(self, key, value):
        value = value.to(self.device)
        target = {"friction_coeffs": self.friction_coeffs, "restitutions": self.restitutions,
                  "payloads": self.payloads, "com_displacements": self.com_displacements,
                  "motor_strengths": self.motor_strengths, "Kp_factors": self.Kp_factors,
                  "Kd_factors": self.Kd_factors}.get(key)
        if target is not None:
            target.copy_(value)

    def _init_env_origins(self):
        """:1220 `_call_train_eval(self._get_env_origins, arange(num_envs))`: one pass per env range with that range's Cfg."""
        N, dev = self.num_envs, self.device
        self.terrain_origins_t = self.terrain_origins_eval_t = None
        self._origin_gen = torch.Generator(device="cpu"); self._origin_gen.manual_seed(self.seed)
        self._get_env_origins(torch.arange(self.num_train_envs, device=dev), self.cfg, False)
        if self.num_eval_envs:
            self._get_env_origins(torch.arange(self.num_train_envs, N, device=dev), self.eval_cfg, True)

    def _get_env_origins(self, env_ids, cfg, is_eval):
        """legged_robot.py:1385-1415 for one env range."""
        dev, n = self.device, len(env_ids)
        if cfg.terrain.mesh_type in ("heightfield", "trimesh"):
            self.custom_origins = True
            t = cfg.terrain
            max_lvl, min_lvl = t.max_init_terrain_level, t.min_init_terrain_level
            if not t.curriculum:
                max_lvl, min_lvl = t.num_rows - 1, 0
            self.terrain_levels[env_ids] = torch.randint(min_lvl, max_lvl + 1, (n,), generator=self._origin_gen).to(dev)
            self.terrain_types[env_ids] = torch.div(torch.arange(n), (n / t.num_cols), rounding_mode="floor").to(torch.long).to(dev)
            t.max_terrain_level = t.num_rows
            table = self.terrain.eval_env_origins if is_eval else self.terrain.env_origins
            table_t = torch.from_numpy(table).to(dev).to(torch.float).contiguous()
            if is_eval:
                self.terrain_origins_eval_t = table_t
            else:
                self.terrain_origins_t = table_t
            t.terrain_origins = table_t
            self.env_origins[env_ids] = table_t[self.terrain_levels[env_ids], self.terrain_types[env_ids]]
        else:
            self.custom_origins = False
            num_cols = np.floor(np.sqrt(n))
            num_rows = np.ceil(self.num_envs / num_cols)           # (:1408: the TOTAL env count, as written)
            xx, yy = torch.meshgrid(torch.arange(num_rows), torch.arange(num_cols), indexing="ij")
            sp = cfg.env.env_spacing
            self.env_origins[env_ids, 0] = (sp * xx.flatten()[:n]).to(dev)
            self.env_origins[env_ids, 1] = (sp * yy.flatten()[:n]).to(dev)
            self.env_origins[env_ids, 2] = 0.

    def _init_command_distribution(self):
        c = self.cfg.commands
        self.curriculum = RewardThresholdCurriculum(
            seed=c.curriculum_seed, device=self.device,
            x_vel=(c.limit_vel_x[0], c.limit_vel_x[1], 51), y_vel=(c.limit_vel_y[0], c.limit_vel_y[1], 2),
            yaw_vel=(c.limit_vel_yaw[0], c.limit_vel_yaw[1], 51))
        self._env_command_bins = torch.zeros(self.num_envs, dtype=torch.long, device=self.device)
        low = np.array([c.lin_vel_x[0], c.lin_vel_y[0], c.ang_vel_yaw[0]])
        high = np.array([c.lin_vel_x[1], c.lin_vel_y[1], c.ang_vel_yaw[1]])
        self.curriculum.set_to(low=low, high=high)

    @property
    def env_command_bins(self):
        """Device tensor [N] int64 (the reference keeps a host numpy array, :1065)."""
        return self._env_command_bins

    def _make_buffers(self):
        b = _lib.RlEnvBuffers()
        P = _lib.ptr
        b.root_states = P(self.root_states); b.dof_state = P(self.dof_state)
        b.contact_forces = P(self.all_contact_forces); b.actions_in = None
        b.torques = P(self.torques); b.obs_buf = P(self.obs_buf)
        b.privileged_obs_buf = P(self.privileged_obs_buf); b.rew_buf = P(self.rew_buf)
        b.reset_buf = P(self._reset_u8); b.time_out_buf = P(self._time_out_u8)
        b.measured_heights = P(self.measured_heights) if self.params.measure_heights else None
        b.last_actions = P(self._last_actions); b.last_dof_vel = P(self._last_dof_vel)
        b.last_root_vel = P(self._last_root_vel); b.joint_pos_target = P(self._joint_pos_target)
        b.base_lin_vel = P(self._base_lin_vel); b.base_ang_vel = P(self._base_ang_vel)
        b.projected_gravity = P(self._projected_gravity)
        b.Kp_factors = P(self._Kp); b.Kd_factors = P(self._Kd); b.motor_strengths = P(self._motor)
        b.friction_coeffs = P(self.friction_coeffs); b.restitutions = P(self.restitutions)
        b.payloads = P(self.payloads); b.com_displacements = P(self._com)
        b.feet_air_time = P(self._feet_air_time); b.last_contacts = P(self._last_contacts_u8)
        b.episode_length_buf = P(self.episode_length_buf); b.commands = P(self.commands)
        b.episode_sums = P(self._episode_sums); b.command_sums = P(self._command_sums)
        b.height_points = P(self._height_points_xy)
        b.height_samples = P(self.height_samples) if self.height_samples is not None else None
        b.noise_u = b.dr_u = b.push_u = None
        b.step_state = None
        b.rew_raw = P(self._rew_raw)
        # terrain heights in a launch of their own, one warp per env (csrc/heights.cu); RL_ENV_HEIGHTS_PREPASS=0: inside
        # the step kernel (same bits)
        import os
        if self.params.measure_heights and os.environ.get("RL_ENV_HEIGHTS_PREPASS", "1") != "0":
            self._height_mean = torch.zeros(self.num_envs, device=self.device)
            b.height_mean = P(self._height_mean)
        else:
            b.height_mean = None
        self._tile_queue = torch.zeros(2, dtype=torch.int32, device=self.device)
        b.tile_queue = P(self._tile_queue)
        return b

    def use_device_step_counter(self, enable=True):
        """Key the in-kernel RNG with a device-resident step counter so that `step()` can be captured
        in a CUDA graph and replayed (kernel arguments, including the host step number, are frozen
        inside a graph).  The counter continues from `common_step_counter`: the next step() keys its draws
        with common_step_counter + 1, exactly as the host-counter path does."""
        if enable:
            self._step_state = torch.tensor([self.common_step_counter + 1, 0], dtype=torch.int64, device=self.device)
            self._bufs.step_state = self._step_state.data_ptr()
        else:
            self._bufs.step_state = None
        self._device_steps = bool(enable)

    def _make_reset_cfg(self, cfg):
        """The constants of reset_idx for one env range (the functions behind _call_train_eval :242-251 read them from the
        range's Cfg)."""
        p = self.params
        r = _lib.RlResetCfg()
        r.num_envs, r.n_terms, r.has_termination = p.num_envs, p.n_terms, p.has_termination
        r.custom_origins = int(self.custom_origins)
        r.terrain_curriculum = int(bool(cfg.terrain.curriculum))
        r.max_terrain_level = int(getattr(cfg.terrain, "max_terrain_level", cfg.terrain.num_rows))
        r.num_terrain_cols = int(cfg.terrain.num_cols)
        r.env_length_half = f32(cfg.terrain.terrain_length / 2)
        r.episode_length_s_half = f32(cfg.env.episode_length_s * 0.5)
        for i, v in enumerate(self.base_init_state.tolist()):
            r.base_init_state[i] = v
        r.x_init_range, r.y_init_range = f32(cfg.terrain.x_init_range), f32(cfg.terrain.y_init_range)
        r.x_init_offset, r.y_init_offset = f32(cfg.terrain.x_init_offset), f32(cfg.terrain.y_init_offset)
        for i, v in enumerate(p.default_dof_pos):
            r.default_dof_pos[i] = v
        dr = cfg.domain_rand
        r.randomize_motor_strength, r.randomize_Kp_factor, r.randomize_Kd_factor = \
            int(bool(dr.randomize_motor_strength)), int(bool(dr.randomize_Kp_factor)), int(bool(dr.randomize_Kd_factor))
        for name, rng in (("motor_strength_lo_span", dr.motor_strength_range), ("Kp_factor_lo_span", dr.Kp_factor_range),
                          ("Kd_factor_lo_span", dr.Kd_factor_range)):
            arr = getattr(r, name)
            arr[0], arr[1] = lo_span(rng)
        return r

    # ------------------------------------------------------------------------------------------
    # hot path
    # ------------------------------------------------------------------------------------------
    def bind_host_io(self, root_states, dof_state, contact_forces, obs, priv, rew, reset_u8):
        """Zero-copy host boundary (the reference's CPU-pipeline case, `sim.use_gpu_pipeline = False`: the
        simulator's state tensors live in host memory).  The arguments are PINNED host tensors of the simulator
        shapes; under unified addressing the kernels read / write them in place over PCIe (TMA bulk loads of
        the AoS rows, bulk stores of the observation rows), so a step is ONE launch with no staging copies:
        reads and writes of different CTAs overlap on the two PCIe directions by themselves.  All persistent
        env state stays in device memory.  Use `step_host` afterwards."""
        want = {"root_states": (root_states, self.root_states), "dof_state": (dof_state, self.dof_state),
                "contact_forces": (contact_forces, self.all_contact_forces), "obs": (obs, self.obs_buf),
                "priv": (priv, self.privileged_obs_buf), "rew": (rew, self.rew_buf), "reset": (reset_u8, self._reset_u8)}
        for name, (h, d) in want.items():
            if not (h.is_pinned() and h.is_contiguous() and h.dtype == d.dtype and h.numel() == d.numel()):
                raise ValueError("bind_host_io: %s must be a pinned, contiguous host tensor of dtype %s with %d elements"
                                 % (name, d.dtype, d.numel()))
        self._host_io = dict(root_states=root_states, dof_state=dof_state, contact_forces=contact_forces, obs=obs, priv=priv,
                             rew=rew, reset=reset_u8)
        b, P = self._bufs, _lib.ptr
        b.root_states, b.dof_state, b.contact_forces = P(root_states), P(dof_state), P(contact_forces)
        b.obs_buf, b.privileged_obs_buf, b.rew_buf, b.reset_buf = P(obs), P(priv), P(rew), P(reset_u8)

    def pack_io(self):
        """Packed host boundary: ONE contiguous device block for what a host-side simulator sends per step
        (`[root_states | dof_state | contact_forces | actions]`, 352 B per env with 13 bodies) and one for what it reads back
        (`[obs | privileged_obs | rew | reset]`), so that a step costs one H2D and one D2H copy instead of four each - over PCIe
        the eight small copies of a 32768-env step take 278 us, the two packed ones 244 us (profiles/jobs/pcie_probe.py).
        Rebinds the env's simulator tensors and output buffers to views of the blocks (call it right after construction) and
        returns `(in_block, out_block, layout_in, layout_out)`: uint8 device tensors and `{name: (byte offset, shape, dtype)}`;
        pinned host blocks of the same layouts, filled / read through `host_views`, are the other end of the two copies."""
        N, nd, nb, dev = self.num_envs, self.num_dof, self.num_bodies, self.device
        if N % 4:
            raise ValueError("pack_io needs num_envs % 4 == 0 (16-byte aligned sections)")
        lay_in, off = {}, 0
        for name, shape in (("root_states", (N, 13)), ("dof_state", (N * nd, 2)), ("contact_forces", (N * nb, 3)),
                            ("actions", (N, self.num_actions))):
            lay_in[name] = (off, shape, torch.float32)
            off += 4 * shape[0] * shape[1]
        in_block = torch.zeros(off, dtype=torch.uint8, device=dev)
        lay_out, off = {}, 0
        for name, shape, dt in (("obs", (N, self.num_obs), torch.float32), ("priv", (N, self.num_privileged_obs), torch.float32),
                                ("rew", (N,), torch.float32), ("reset", (N,), torch.uint8)):
            lay_out[name] = (off, shape, dt)
            off += int(np.prod(shape)) * (4 if dt == torch.float32 else 1)
        out_block = torch.zeros((off + 15) // 16 * 16, dtype=torch.uint8, device=dev)

        def view(block, spec):
            o, shape, dt = spec
            n = int(np.prod(shape)) * (4 if dt == torch.float32 else 1)
            return block[o:o + n].view(dt).view(*shape)
        sim = self.sim
        for name, attr in (("root_states", "root_states"), ("dof_state", "dof_state"), ("contact_forces", "contact_forces")):
            v = view(in_block, lay_in[name])
            v.copy_(getattr(sim, attr))
            setattr(sim, attr, v)
        self.all_root_states = self.root_states = sim.root_states
        self.all_dof_state = self.dof_state = sim.dof_state
        self.all_contact_forces = sim.contact_forces
        self.dof_pos = self.dof_state.view(N, nd, 2)[..., 0]
        self.dof_vel = self.dof_state.view(N, nd, 2)[..., 1]
        self.base_quat = self.root_states[:, 3:7]
        self.contact_forces = self.all_contact_forces.view(N, -1, 3)
        self.packed_actions = view(in_block, lay_in["actions"])
        self.obs_buf = view(out_block, lay_out["obs"]); self.privileged_obs_buf = view(out_block, lay_out["priv"])
        self.rew_buf = view(out_block, lay_out["rew"])
        self._reset_u8 = view(out_block, lay_out["reset"]); self._reset_u8.fill_(1)
        self.reset_buf = self._reset_u8.view(torch.bool)
        b, P = self._bufs, _lib.ptr
        b.root_states, b.dof_state, b.contact_forces = P(self.root_states), P(self.dof_state), P(self.all_contact_forces)
        b.obs_buf, b.privileged_obs_buf, b.rew_buf, b.reset_buf = P(self.obs_buf), P(self.privileged_obs_buf), P(self.rew_buf), P(self._reset_u8)
        self._packed = (in_block, out_block, lay_in, lay_out)
        return self._packed

    @staticmethod
    def host_views(block, layout):
        """Typed views of a (pinned) host block laid out like `pack_io`'s device blocks."""
        out = {}
        for name, (o, shape, dt) in layout.items():
            n = int(np.prod(shape)) * (4 if dt == torch.float32 else 1)
            out[name] = block[o:o + n].view(dt).view(*shape)
        return out

    def step_host(self, actions_host):
        """`step` for an env bound with `bind_host_io`: `actions_host` is a pinned host tensor [N, num_actions].
        Returns the bound host tensors (obs, priv, rew, reset_u8); they are valid once the stream has been
        synchronised."""
        io = self._host_io
        if not (actions_host.is_pinned() and actions_host.is_contiguous() and actions_host.dtype == torch.float32 and
                tuple(actions_host.shape) == (self.num_envs, self.num_actions)):
            raise ValueError("step_host: actions must be a pinned contiguous float32 host tensor [%d, %d]"
                             % (self.num_envs, self.num_actions))
        self.common_step_counter += 1
        b = self._bufs
        b.actions_in = actions_host.data_ptr()
        b.noise_u = b.dr_u = b.push_u = None
        host_step = 0 if b.step_state else self.common_step_counter
        _lib.check(self._lib.rl_env_step_fused(C.byref(self._cfg_struct), C.byref(b), self.seed, host_step,
                                               _lib.current_stream()))
        return io["obs"], io["priv"], io["rew"], io["reset"]

    def step(self, actions):
        """legged_robot.py:106-137.  Synthetic simulator state: one fused launch (torques + post-physics pipeline).
        Live simulator (`GymApiSim`): the reference's loop - `decimation` x [torque kernel -> physics sub-step] ->
        refresh -> one post-physics launch."""
        actions = actions.to(self.device, torch.float)
        if not actions.is_contiguous():
            actions = actions.contiguous()
        if actions.shape != (self.num_envs, self.num_actions):
            raise ValueError("actions must be [%d, %d], got %s" % (self.num_envs, self.num_actions, tuple(actions.shape)))
        if self.upstream_order:
            self._upstream_resample()
        if getattr(self.sim, "live", False):
            for _ in range(self.cfg.control.decimation):                       # :116-126
                self.sim.apply_torques_and_step(self._compute_torques(actions))
            self.sim.refresh()                                                 # :143-146, :156, :165-170
            self.post_physics_step(actions)
            if self._moved_roots():                                            # teleport / push wrote root rows (:789, :765)
                self.sim.push_root_state(torch.arange(self.num_envs, device=self.device))
            if self.upstream_order:
                self._upstream_reset()
            return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras
        self._actions_in = actions
        self.common_step_counter += 1
        b = self._bufs
        b.actions_in = actions.data_ptr()
        inj = self._inject
        b.noise_u = _lib.ptr(inj.get("noise_u")); b.dr_u = _lib.ptr(inj.get("dr_u")); b.push_u = _lib.ptr(inj.get("push_u"))
        host_step = 0 if b.step_state else self.common_step_counter
        _lib.check(self._lib.rl_env_step_fused(C.byref(self._cfg_struct), C.byref(b), self.seed,
                                               host_step, _lib.current_stream()))
        if self._custom_terms or self._hook_overrides:
            self._run_plugins()
        if self.upstream_order:
            self._upstream_reset()
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def _upstream_resample(self):
        """:578-581 in upstream order: envs whose episode length (after this step's increment) hits the resampling
        interval draw new commands BEFORE the step's rewards and observations are computed."""
        interval = self.params.resample_interval
        ids = ((self.episode_length_buf + 1) % interval == 0).nonzero(as_tuple=False).flatten()
        if len(ids):
            self._resample_commands(ids)

    def _upstream_reset(self):
        """:175-177 in upstream order: reset_idx of the envs that terminated (or timed out) in this step, commands
        included (:246).  Deviation from upstream, stated: upstream builds the observations AFTER this reset; here the
        fused launch has already written them, so an env that resets returns its last pre-reset observation (the next
        step's observation is the first of the new episode)."""
        ids = self.reset_buf.nonzero(as_tuple=False).flatten()
        if len(ids):
            self.reset_idx(ids)
            self._resample_commands(ids)

    def _moved_roots(self):
        """True when the step may have rewritten root rows that a live simulator must be told about."""
        p = self.params
        return bool(p.teleport_robots or p.push_robots)

    def _compute_torques(self, actions):
        """legged_robot.py:653-688 as a standalone launch (the entry a real simulator loop calls
        `decimation` times, :116-126).  Returns the [N,12] torque tensor."""
        actions = actions.to(self.device, torch.float).contiguous()
        b = self._bufs
        b.actions_in = actions.data_ptr()
        _lib.check(self._lib.rl_env_torques(C.byref(self._cfg_struct), C.byref(b), _lib.current_stream()))
        return self.torques

    def post_physics_step(self, actions=None):
        """legged_robot.py:139-188 with the torques already in `self.torques`."""
        if actions is not None:
            self._actions_in = actions.to(self.device, torch.float).contiguous()
        self.common_step_counter += 1
        b = self._bufs
        b.actions_in = self._actions_in.data_ptr()
        inj = self._inject
        b.noise_u = _lib.ptr(inj.get("noise_u")); b.dr_u = _lib.ptr(inj.get("dr_u")); b.push_u = _lib.ptr(inj.get("push_u"))
        host_step = 0 if b.step_state else self.common_step_counter
        _lib.check(self._lib.rl_env_post_physics(C.byref(self._cfg_struct), C.byref(b), self.seed,
                                                 host_step, _lib.current_stream()))
        if self._custom_terms or self._hook_overrides:
            self._run_plugins()

    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self, horizon=0):
        if horizon != 0:
            raise NotImplementedError("privileged_future_horizon > 0 reads next_privileged_obs_buf, which the "
                                      "reference never defines (base_task.py:89-97)")
        return self.privileged_obs_buf

    def reset(self):
        """base_task.py:103-108."""
        self.reset_idx(torch.arange(self.num_envs, device=self.device))
        obs, privileged_obs, _, _, _ = self.step(torch.zeros(self.num_envs, self.num_actions, device=self.device))
        return obs, privileged_obs

    def reset_idx(self, env_ids, obs_history=None):
        """legged_robot.py:227-290: one launch per env range (+ the host-side uniform command curriculum)."""
        if len(env_ids) == 0:
            return
        env_ids = env_ids.to(self.device, torch.long).contiguous()
        cfg = self.cfg
        if self.num_eval_envs == 0:
            self._update_command_curriculum_uniform(env_ids, cfg)
            self._reset_launch(env_ids, cfg, self._reset_cfg, self.terrain_origins_t, obs_history)
            self._fill_train_extras(env_ids)
        else:
            # _call_train_eval :456-469: every per-range function runs on the training ids with the training Cfg, then on
            # the evaluation ids with the evaluation Cfg
            train_ids = env_ids[env_ids < self.num_train_envs].contiguous()
            eval_ids = env_ids[env_ids >= self.num_train_envs].contiguous()
            if len(train_ids):
                self._update_command_curriculum_uniform(train_ids, cfg)
            if len(eval_ids):
                self._update_command_curriculum_uniform(eval_ids, self.eval_cfg)
                # :268-276 - the evaluation rollout result is saved once per env before its accumulators are cleared (the
                # launch below clears them)
                self.extras["eval/episode"] = {}
                for key in self.episode_sums:
                    saved = self.episode_sums_eval[key]
                    unset = eval_ids[saved[eval_ids] == -1]
                    saved[unset] = self.episode_sums[key][unset]
            if len(train_ids):
                self._reset_launch(train_ids, cfg, self._reset_cfg, self.terrain_origins_t, obs_history)
                self._fill_train_extras(train_ids)
            if len(eval_ids):
                self._reset_launch(eval_ids, self.eval_cfg, self._reset_cfg_eval, self.terrain_origins_eval_t, obs_history)
                for name, _ in self._custom_terms:
                    self.episode_sums[name][eval_ids] = 0.
        # (:279-290: curriculum info and time-outs are written whatever the ids were)
        if cfg.terrain.curriculum:
            self.extras.setdefault("train/episode", {})["terrain_level"] = torch.mean(self.terrain_levels[:self.num_train_envs].float())
        if cfg.commands.command_curriculum:
            # (a persistent buffer rewritten in place: captured rollout graphs keep reading this address)
            if getattr(self, "_env_bins_f", None) is None:
                self._env_bins_f = torch.zeros(self.num_envs, device=self.device)
            self._env_bins_f.copy_(self._env_command_bins)
            self.extras["env_bins"] = self._env_bins_f[:self.num_train_envs]
            self.extras.setdefault("train/episode", {})["command_area"] = (self.curriculum.weights_device.sum() / len(self.curriculum))
        if cfg.commands.yaw_command_curriculum:
            self.extras.setdefault("train/episode", {})["max_command_yaw"] = cfg.command_ranges["ang_vel_yaw"][1]
            if self.eval_cfg is not None:
                self.extras.setdefault("eval/episode", {})["max_command_yaw"] = self.eval_cfg.command_ranges["ang_vel_yaw"][1]
        if cfg.env.send_timeouts:
            self.extras["time_outs"] = self.time_out_buf[:self.num_train_envs]

    def _reset_launch(self, env_ids, cfg, rc, terrain_origins, obs_history):
        """The reset kernel on `env_ids` with the constants of their range (`rc`) and that range's terrain-origin table."""
        rc.terrain_curriculum = int(bool(cfg.terrain.curriculum) and self.init_done)
        b = _lib.RlResetBuffers()
        P = _lib.ptr
        b.mask = None; b.ids = P(env_ids); b.n_ids = int(env_ids.numel())
        b.root_states = P(self.root_states); b.dof_state = P(self.dof_state); b.env_origins = P(self.env_origins)
        b.terrain_levels = P(self.terrain_levels); b.terrain_types = P(self.terrain_types)
        b.terrain_origins = P(terrain_origins); b.commands = P(self.commands)
        b.last_actions = P(self._last_actions); b.last_dof_vel = P(self._last_dof_vel)
        b.feet_air_time = P(self._feet_air_time); b.episode_length_buf = P(self.episode_length_buf)
        b.reset_buf = P(self._reset_u8)
        b.Kp_factors = P(self._Kp); b.Kd_factors = P(self._Kd); b.motor_strengths = P(self._motor)
        b.episode_sums = P(self._episode_sums)
        self._episode_sum_out.zero_()
        b.episode_sum_out = P(self._episode_sum_out)
        b.obs_history = P(obs_history); b.obs_history_len = 0 if obs_history is None else int(obs_history.shape[1])
        inj = self._inject
        b.dr_u = P(inj.get("reset_dr_u")); b.init_u = P(inj.get("init_u")); b.level_u = P(inj.get("level_u"))
        _lib.check(self._lib.rl_env_reset(C.byref(rc), C.byref(b), self.seed, self.common_step_counter,
                                          _lib.current_stream()))
        if getattr(self.sim, "live", False):
            # hand the rewritten rows to the simulator: int32 actor ids of the reset envs (:713-717, :739-741)
            self.sim.push_dof_state(env_ids)
            self.sim.push_root_state(env_ids)

    def _fill_train_extras(self, train_ids):
        """:261-267: mean episode sums of the training envs that were just reset (device-side, no host sync)."""
        sums = self._episode_sum_out
        means = (sums[:-1] / sums[-1]).to(torch.float)
        self.extras["train/episode"] = {"rew_" + n: means[r] for n, r in self._episode_rows.items()}
        for name, _ in self._custom_terms:          # plugin accumulators (:264-267)
            self.extras["train/episode"]["rew_" + name] = torch.mean(self.episode_sums[name][train_ids])
            self.episode_sums[name][train_ids] = 0.

    def reset_evaluation_envs(self):
        """legged_robot.py:204-225: log the finished evaluation batch, widen the evaluation command ranges, reset every
        evaluation env and clear the saved results."""
        if self.eval_cfg is None:
            return
        env_ids_eval = torch.arange(self.num_train_envs, self.num_envs, device=self.device)
        # (as in the reference, extras["eval/episode"] must exist - it does once an evaluation env went through reset_idx)
        for key, saved in self.episode_sums_eval.items():
            unset = env_ids_eval[saved[env_ids_eval] == -1]
            saved[unset] = self.episode_sums[key][unset]
            self.extras["eval/episode"]["rew_" + key] = torch.mean(saved[saved != -1])
        self._update_command_curriculum_uniform(env_ids_eval, self.eval_cfg)       # :218
        self.reset_idx(env_ids_eval)
        for key in self.episode_sums_eval:
            self.episode_sums_eval[key] = -1 * torch.ones(self.num_envs, device=self.device)

    def _update_command_curriculum_uniform(self, env_ids, cfg=None):
        """legged_robot.py:851-880: rare (every max_episode_length steps) host-side range widening.  `cfg` is the Cfg of the
        env range (its command_ranges are widened; its curriculum switches and clip limits apply); the episode length and the
        thresholds come from the training Cfg, the accumulators are always `self.episode_sums` (:849)."""
        cfg = self.cfg if cfg is None else cfg
        max_len = self.cfg.env.max_episode_length
        if self.common_step_counter % max_len != 0:
            return
        rs = self.reward_scales
        tc = self.cfg.commands

        def widen(key, sum_name, thr, lo_clip, hi_clip):
            mean = torch.mean(self.episode_sums[sum_name][env_ids]) / max_len
            if mean > thr * rs[sum_name]:
                r = cfg.command_ranges[key]
                r[0] = np.clip(r[0] - 0.2, -lo_clip, 0.0)
                r[1] = np.clip(r[1] + 0.2, 0.0, hi_clip)
        c = cfg.commands
        if c.command_curriculum and rs.get("tracking_lin_vel", 0) > 0:
            widen("lin_vel_x", "tracking_lin_vel", tc.forward_curriculum_threshold, c.max_reverse_curriculum,
                  c.max_forward_curriculum)
        if c.yaw_command_curriculum and rs.get("tracking_ang_vel", 0) > 0:
            widen("ang_vel_yaw", "tracking_ang_vel", tc.yaw_curriculum_threshold, c.max_yaw_curriculum,
                  c.max_yaw_curriculum)

    def _randomize_rigid_body_props(self, env_ids, cfg):
        """legged_robot.py:519-541 (init-time draw; rigid-body props live in the simulator)."""
        dr, n = cfg.domain_rand, len(env_ids)

        def draw(k=None):
            shape = (n,) if k is None else (n, k)
            return torch.rand(*shape, device=self.device, generator=self._gen)
        if dr.randomize_base_mass:
            lo, hi = dr.added_mass_range
            self.payloads[env_ids] = draw() * (hi - lo) + lo
        if dr.randomize_com_displacement:
            lo, hi = dr.com_displacement_range
            self.com_displacements[env_ids, :] = draw(3) * (hi - lo) + lo
        if dr.randomize_friction:
            lo, hi = dr.friction_range
            self.friction_coeffs[env_ids] = draw() * (hi - lo) + lo
        if dr.randomize_restitution:
            lo, hi = dr.restitution_range
            self.restitutions[env_ids] = draw() * (hi - lo) + lo

    def _resample_commands(self, env_ids):
        """legged_robot.py:595-626 + curriculum.py:110-119,55-68 on the device."""
        if len(env_ids) == 0:
            return
        env_ids = env_ids.to(self.device, torch.long).contiguous()
        p, cur = self.params, self.curriculum
        g = _lib.RlGacCfg()
        g.num_envs, g.n_bins = self.num_envs, len(cur)
        for i in range(3):
            g.dims[i] = cur.dims[i]
            g.bin_size[i] = float(list(cur.bin_sizes.values())[i])
        timesteps = int(self.cfg.commands.resampling_time / self.dt)
        g.ep_len = f32(min(self.cfg.env.max_episode_length, timesteps))
        g.lin_threshold = f32(self.cfg.commands.forward_curriculum_threshold * self.reward_scales["tracking_lin_vel"])
        g.ang_threshold = f32(self.cfg.commands.yaw_curriculum_threshold * self.reward_scales["tracking_ang_vel"])
        g.lin_slot, g.ang_slot = p.sum_rows["tracking_lin_vel"], p.sum_rows["tracking_ang_vel"]
        g.n_command_sums = self._command_sums.shape[0]
        g.num_train_envs = self.num_train_envs
        nlo, nhi = cur.neighbour_ranges(0.5)
        b = _lib.RlGacBuffers()
        P = _lib.ptr
        b.mask = None; b.ids = P(env_ids); b.n_ids = int(env_ids.numel())
        b.weights = P(cur.weights_device); b.centers = P(cur.centers_device)
        b.nbr_lo, b.nbr_hi = P(nlo), P(nhi)
        b.hit_count, b.own_flag, b.cdf = P(cur.hit_count), P(cur.own_flag), P(cur.cdf)
        b.env_command_bins = P(self._env_command_bins); b.commands = P(self.commands)
        b.command_sums = P(self._command_sums)
        u_bin = self._inject.get("gac_u_bin"); u_cell = self._inject.get("gac_u_cell")
        if u_bin is None and self.gac_rng == "numpy":
            # replay of the reference's host generator (curriculum.py:57,64): choice() consumes one
            # uniform per env, then each env draws three - bit-identical command streams
            n = int(env_ids.numel())
            ub = cur.rng.random_sample(n)
            uc = cur.rng.random_sample((n, 3))
            u_bin = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
            u_cell = torch.zeros(self.num_envs, 3, dtype=torch.float64, device=self.device)
            u_bin[env_ids] = torch.from_numpy(ub).to(self.device)
            u_cell[env_ids] = torch.from_numpy(uc).to(self.device)
        self._gac_keepalive = (u_bin, u_cell, env_ids)
        b.u_bin, b.u_cell = P(u_bin), P(u_cell)
        st = _lib.current_stream()
        _lib.check(self._lib.rl_gac_scatter(C.byref(g), C.byref(b), st))
        self._gac_allreduce()
        _lib.check(self._lib.rl_gac_update_sample(C.byref(g), C.byref(b), self.seed, self.common_step_counter, st))

    def _gac_allreduce(self):
        """Multi-GPU hook: sum the int32 incidence counters over ranks so every rank applies the
        identical saturating update."""
        from ..sharding import all_reduce_sum_
        all_reduce_sum_(self.curriculum.hit_count, self.curriculum.own_flag)

    # ------------------------------------------------------------------------------------------
    # plugin surface of the reference class (legged_robot.py:190, :314, :342, :1074-1093, :1469)
    # ------------------------------------------------------------------------------------------
    # The fused launch evaluates the reference's default pipeline.  What a subclass can still do, as with the
    # reference class:
    #   * define `_reward_<name>(self)` and give `Cfg.rewards.scales.<name>` a non-zero value: the method runs after
    #     the launch and its term enters rew_buf / episode_sums / command_sums where compute_reward :320-327 adds it
    #     (before the positive clip; the kernel hands out the unclipped sum for that);
    #   * override `check_termination`, `compute_reward`, `compute_observations`: the base methods are the results of
    #     the launch (already in reset_buf / rew_buf / obs_buf), so the usual `super().check_termination();
    #     self.reset_buf |= ...` idiom works; overrides run after the launch in the reference's order (:164-180).
    #   * call `_get_heights()` for the terrain heights under any env (a kernel of its own).
    def _is_custom_reward(self, name):
        """True when `_reward_<name>` must run as a Python method: not a fused term, or overridden by a subclass."""
        meth = getattr(type(self), "_reward_" + name, None)
        if name not in _lib.REWARD_TERM_IDS:
            if meth is None:
                raise AttributeError("'%s' object has no attribute '_reward_%s'" % (type(self).__name__, name))   # :1093
            return True
        return meth is not None

    def __getattr__(self, name):
        if name.startswith("_reward_"):
            term = name[len("_reward_"):]
            if term in _lib.REWARD_TERM_IDS:
                return lambda: self._fused_term(term)
        raise AttributeError("'%s' object has no attribute '%s'" % (type(self).__name__, name))

    def _fused_term(self, term):
        raise NotImplementedError("reward term %r is evaluated inside the fused kernel; per-term values are "
                                  "available through episode_sums / command_sums" % term)

    def check_termination(self):
        """legged_robot.py:190-202: `reset_buf` (and `time_out_buf`) of the last fused launch."""
        return self.reset_buf

    def compute_reward(self):
        """legged_robot.py:314-340: `rew_buf` and the accumulators of the last fused launch (plugin terms included)."""
        return self.rew_buf

    def compute_observations(self):
        """legged_robot.py:342-417: `obs_buf` / `privileged_obs_buf` of the last fused launch."""
        return self.obs_buf

    def _get_heights(self, env_ids=None, cfg=None):
        """legged_robot.py:1469-1503: terrain heights [n, num_height_points] under the robots (min of three truncated
        samples per point), from the current root states.  The fused step computes the same values in-kernel."""
        p = self.params
        if self.cfg.terrain.mesh_type == "plane" or not p.measure_heights:
            n = self.num_envs if env_ids is None else len(env_ids)
            return torch.zeros(n, max(1, p.num_height_points), device=self.device)
        if self.cfg.terrain.mesh_type == "none":
            raise NameError("Can't measure height with terrain mesh type 'none'")
        ids = torch.arange(self.num_envs, device=self.device) if env_ids is None else env_ids.to(self.device, torch.long)
        out = torch.empty(len(ids), p.num_height_points, device=self.device)
        _lib.check(self._lib.rl_env_heights(C.byref(self._cfg_struct), _lib.ptr(self.root_states), _lib.ptr(self._height_points_xy),
                                            _lib.ptr(self.height_samples), _lib.ptr(ids.contiguous()), int(len(ids)),
                                            _lib.ptr(out), _lib.current_stream()))
        return out

    def _run_plugins(self):
        """After the fused launch, in the reference's order (:164-180): termination override, reward plugins, reward
        override, observation override."""
        p = self.params
        if "check_termination" in self._hook_overrides:
            self.check_termination()
        if self._custom_terms:
            raw = self._rew_raw                                   # sum of the fused terms before clip / termination
            clipped = torch.clip(raw, min=0.) if p.only_positive_rewards else raw
            after_clip = self.rew_buf - clipped                   # the termination term (:330-334), added after the clip
            total = raw.clone()
            for name, scale in self._custom_terms:
                rew = getattr(self, "_reward_" + name)() * scale
                total += rew
                self.episode_sums[name] += rew
                self.command_sums[name] += rew
            new = torch.clip(total, min=0.) if p.only_positive_rewards else total
            self.episode_sums["total"] += new - clipped
            self.rew_buf.copy_(new + after_clip)
        if "compute_reward" in self._hook_overrides:
            self.compute_reward()
        if "compute_observations" in self._hook_overrides:
            self.compute_observations()

    def close(self):
        pass
