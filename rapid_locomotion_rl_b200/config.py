"""`Cfg` namespace tree + robot presets + the freeze step that turns it into kernel constants.

Mirrors the reference's configuration surface (mini_gym/envs/base/legged_robot_config.py:6-256,
mini_gym/envs/mini_cheetah/mini_cheetah_config.py:8-105, mini_gym/envs/go1/go1_config.py:8-106):
class-attribute namespaces that callers mutate (`Cfg.env.num_envs = 4000`) before building an
env.  Any object with the same attribute tree (including the reference's own params_proto
`Cfg`) is accepted by `freeze_env_cfg`.

The defaults are held as one data table (`DEFAULTS`) and the class tree is generated from it;
`new_cfg()` returns a fresh, independent tree so tests and multi-env processes do not share a
process-global config the way the reference does.
"""
import copy
import math

import numpy as np

from . import _lib
from .robots import RobotSpec, robot_for_asset

DEFAULTS = {
    "env": dict(
        num_envs=4096, num_observations=235, num_privileged_obs=18, privileged_future_horizon=1,
        num_actions=12, num_observation_history=15, env_spacing=3.0, send_timeouts=True,
        episode_length_s=20, observe_vel=True, observe_only_ang_vel=False, observe_only_lin_vel=False,
        observe_yaw=False, observe_command=True, record_video=True,
        priv_observe_friction=True, priv_observe_restitution=True, priv_observe_base_mass=True,
        priv_observe_com_displacement=True, priv_observe_motor_strength=True,
        priv_observe_Kp_factor=True, priv_observe_Kd_factor=True),
    "terrain": dict(
        mesh_type="trimesh", horizontal_scale=0.1, vertical_scale=0.005, border_size=0, curriculum=True,
        static_friction=1.0, dynamic_friction=1.0, restitution=0.0, terrain_noise_magnitude=0.1,
        terrain_smoothness=0.005, measure_heights=True,
        measured_points_x=[round(-0.8 + 0.1 * i, 1) for i in range(17)],
        measured_points_y=[round(-0.5 + 0.1 * i, 1) for i in range(11)],
        selected=False, terrain_kwargs=None, min_init_terrain_level=0, max_init_terrain_level=5,
        terrain_length=8.0, terrain_width=8.0, num_rows=10, num_cols=20,
        terrain_proportions=[0.1, 0.1, 0.35, 0.25, 0.2], slope_treshold=0.75, difficulty_scale=1.0,
        x_init_range=1.0, y_init_range=1.0, x_init_offset=0.0, y_init_offset=0.0,
        teleport_robots=True, teleport_thresh=2.0, max_platform_height=0.2),
    "commands": dict(
        command_curriculum=False, max_reverse_curriculum=1.0, max_forward_curriculum=1.0,
        forward_curriculum_threshold=0.8, yaw_command_curriculum=False, max_yaw_curriculum=1.0,
        yaw_curriculum_threshold=0.5, num_commands=4, resampling_time=10.0, heading_command=True,
        global_reference=False, num_lin_vel_bins=20, lin_vel_step=0.3, num_ang_vel_bins=20,
        ang_vel_step=0.3, distribution_update_extension_distance=1, curriculum_seed=100,
        lin_vel_x=[-1.0, 1.0], lin_vel_y=[-1.0, 1.0], ang_vel_yaw=[-1, 1],
        body_height_cmd=[-0.05, 0.05], impulse_height_commands=False,
        limit_vel_x=[-10.0, 10.0], limit_vel_y=[-0.6, 0.6], limit_vel_yaw=[-10.0, 10.0],
        heading=[-3.14, 3.14]),
    "init_state": dict(
        pos=[0.0, 0.0, 1.0], rot=[0.0, 0.0, 0.0, 1.0], lin_vel=[0.0, 0.0, 0.0], ang_vel=[0.0, 0.0, 0.0],
        default_joint_angles={"joint_a": 0.0, "joint_b": 0.0}),
    "control": dict(
        control_type="P", stiffness={"joint_a": 10.0, "joint_b": 15.0},
        damping={"joint_a": 1.0, "joint_b": 1.5}, action_scale=0.5, hip_scale_reduction=1.0, decimation=4),
    "asset": dict(
        file="", foot_name="None", penalize_contacts_on=[], terminate_after_contacts_on=[],
        disable_gravity=False, collapse_fixed_joints=True, fix_base_link=False, default_dof_drive_mode=3,
        self_collisions=0, replace_cylinder_with_capsule=True, flip_visual_attachments=True,
        density=0.001, angular_damping=0.0, linear_damping=0.0, max_angular_velocity=1000.0,
        max_linear_velocity=1000.0, armature=0.0, thickness=0.01),
    "domain_rand": dict(
        rand_interval_s=10, randomize_friction=True, friction_range=[0.5, 1.25],
        randomize_restitution=False, restitution_range=[0, 1.0], randomize_base_mass=False,
        added_mass_range=[-1.0, 1.0], randomize_com_displacement=False,
        com_displacement_range=[-0.15, 0.15], randomize_motor_strength=False,
        motor_strength_range=[0.9, 1.1], randomize_Kp_factor=False, Kp_factor_range=[0.8, 1.3],
        randomize_Kd_factor=False, Kd_factor_range=[0.5, 1.5], push_robots=True, push_interval_s=15,
        max_push_vel_xy=1.0),
    "rewards": dict(
        only_positive_rewards=True, tracking_sigma=0.25, tracking_sigma_lat=0.25,
        tracking_sigma_long=0.25, tracking_sigma_yaw=0.25, soft_dof_pos_limit=1.0,
        soft_dof_vel_limit=1.0, soft_torque_limit=1.0, base_height_target=1.0, max_contact_force=100.0,
        use_terminal_body_height=False, terminal_body_height=0.20,
        scales=dict(
            termination=-0.0, tracking_lin_vel=1.0, tracking_ang_vel=0.5, lin_vel_z=-2.0,
            ang_vel_xy=-0.05, orientation=-0.0, torques=-0.00001, dof_vel=-0.0, dof_acc=-2.5e-7,
            base_height=-0.0, feet_air_time=1.0, collision=-1.0, feet_stumble=-0.0, action_rate=-0.01,
            stand_still=-0.0, tracking_lin_vel_lat=0.0, tracking_lin_vel_long=0.0)),
    "normalization": dict(
        obs_scales=dict(lin_vel=2.0, ang_vel=0.25, dof_pos=1.0, dof_vel=0.05, height_measurements=5.0,
                        body_height_cmd=2.0),
        clip_observations=100.0, clip_actions=100.0, friction_range=[0.05, 4.5],
        restitution_range=[0, 1.0], added_mass_range=[-1.0, 3.0], com_displacement_range=[-0.1, 0.1],
        motor_strength_range=[0.9, 1.1], Kp_factor_range=[0.8, 1.3], Kd_factor_range=[0.5, 1.5]),
    "noise": dict(
        add_noise=True, noise_level=1.0,
        noise_scales=dict(dof_pos=0.01, dof_vel=1.5, lin_vel=0.1, ang_vel=0.2, gravity=0.05,
                          height_measurements=0.1)),
    "viewer": dict(ref_env=0, pos=[-10, 0, 6], lookat=[0.0, 0, 3.0]),
    "sim": dict(dt=0.005, substeps=1, gravity=[0.0, 0.0, -9.81], up_axis=1, use_gpu_pipeline=True),
}

_QUADRUPED_DR = dict(
    randomize_base_mass=True, added_mass_range=[-1, 3], push_robots=False, max_push_vel_xy=0.5,
    randomize_friction=True, friction_range=[0.05, 4.5], randomize_restitution=True,
    restitution_range=[0.0, 1.0], restitution=0.5, randomize_com_displacement=True,
    com_displacement_range=[-0.1, 0.1], randomize_motor_strength=True, motor_strength_range=[0.9, 1.1],
    randomize_Kp_factor=False, Kp_factor_range=[0.8, 1.3], randomize_Kd_factor=False,
    Kd_factor_range=[0.5, 1.5], rand_interval_s=6)
_QUADRUPED_COMMANDS = dict(
    heading_command=False, resampling_time=10.0, command_curriculum=True, num_lin_vel_bins=30,
    num_ang_vel_bins=30, lin_vel_x=[-0.6, 0.6], lin_vel_y=[-0.6, 0.6], ang_vel_yaw=[-1, 1])


def _joint_angles(hip, thigh_front, thigh_rear, calf):
    out = {}
    for leg in ("FL", "RL", "FR", "RR"):
        out["%s_hip_joint" % leg] = hip if leg[1] == "L" else -hip
    for leg in ("FL", "RL", "FR", "RR"):
        out["%s_thigh_joint" % leg] = thigh_front if leg[0] == "F" else thigh_rear
    for leg in ("FL", "RL", "FR", "RR"):
        out["%s_calf_joint" % leg] = calf
    return out


PRESETS = {
    # mini_cheetah_config.py:8-105
    "mini_cheetah": {
        "init_state": dict(pos=[0.0, 0.0, 0.32], default_joint_angles=_joint_angles(0.1, -0.8, -0.8, 1.62)),
        "control": dict(control_type="P", stiffness={"joint": 20.0}, damping={"joint": 0.5},
                        action_scale=0.25, hip_scale_reduction=0.5, decimation=4),
        "asset": dict(file="{MINI_GYM_ROOT_DIR}/resources/robots/mini_cheetah/urdf/mini_cheetah.urdf",
                      foot_name="calf", penalize_contacts_on=[], terminate_after_contacts_on=["base", "thigh"],
                      self_collisions=0, flip_visual_attachments=False, fix_base_link=False),
        "rewards": dict(soft_dof_pos_limit=0.9, base_height_target=0.30,
                        scales=dict(torques=-0.0002, dof_pos_limits=-10.0, orientation=-5.0, base_height=-30.0)),
        "terrain": dict(mesh_type="trimesh", measure_heights=False, terrain_noise_magnitude=0.0,
                        teleport_robots=True, border_size=50,
                        terrain_proportions=[0, 0, 0, 0, 0, 0, 0, 0, 1.0], curriculum=False),
        "env": dict(num_observations=42, observe_vel=False, num_envs=4000),
        "commands": _QUADRUPED_COMMANDS,
        "domain_rand": _QUADRUPED_DR,
    },
    # go1_config.py:8-106
    "go1": {
        "init_state": dict(pos=[0.0, 0.0, 0.34], default_joint_angles=_joint_angles(0.1, 0.8, 1.0, -1.5)),
        "control": dict(control_type="P", stiffness={"joint": 20.0}, damping={"joint": 0.5},
                        action_scale=0.25, hip_scale_reduction=0.5, decimation=4),
        "asset": dict(file="{MINI_GYM_ROOT_DIR}/resources/robots/go1/urdf/go1.urdf", foot_name="foot",
                      penalize_contacts_on=["thigh", "calf"], terminate_after_contacts_on=["base"],
                      self_collisions=0, flip_visual_attachments=False, fix_base_link=False),
        "rewards": dict(soft_dof_pos_limit=0.9, base_height_target=0.34,
                        scales=dict(torques=-0.0001, action_rate=-0.01, dof_pos_limits=-10.0,
                                    orientation=-5.0, base_height=-30.0)),
        "terrain": dict(mesh_type="plane", measure_heights=False, terrain_noise_magnitude=0.0,
                        teleport_robots=False, border_size=50,
                        terrain_proportions=[0, 0, 0, 0, 0, 0, 0, 0, 1.0], curriculum=False),
        "env": dict(num_observations=42, observe_vel=False, num_envs=4096),
        "commands": _QUADRUPED_COMMANDS,
        "domain_rand": _QUADRUPED_DR,
    },
}


class _NamespaceMeta(type):
    def _update(cls, d=None, **kw):
        for k, v in dict(d or {}, **kw).items():
            cur = cls.__dict__.get(k)
            if isinstance(v, dict) and isinstance(cur, _NamespaceMeta):
                cur._update(v)
            else:
                setattr(cls, k, copy.deepcopy(v))

    def _public(cls):
        """Fresh deep dict of the public attributes (what `vars(ProtoCls)` gives in the reference)."""
        return public_vars(cls)


def _make_namespace(name, table, nested=("scales", "obs_scales", "noise_scales")):
    ns = {}
    for k, v in table.items():
        if isinstance(v, dict) and (k in nested or name == "Cfg"):
            ns[k] = _make_namespace(k, v, nested)
        else:
            ns[k] = copy.deepcopy(v)
    return _NamespaceMeta(name, (), ns)


def new_cfg():
    """A fresh `Cfg` tree with the reference's defaults."""
    return _make_namespace("Cfg", DEFAULTS)


Cfg = new_cfg()


def public_vars(ns):
    """Public attributes of a namespace class (ours or params_proto's), in definition order."""
    out = {}
    for klass in reversed(getattr(ns, "__mro__", (ns,))):
        d = type.__dict__["__dict__"].__get__(klass) if isinstance(klass, type) else vars(klass)
        for k, v in d.items():
            if k.startswith("_") or callable(v) and not isinstance(v, type) or isinstance(v, (classmethod, staticmethod, property)):
                continue
            out[k] = v if isinstance(v, type) else copy.deepcopy(v)
    return out


def _apply_preset(cnfg, preset):
    for group, values in PRESETS[preset].items():
        target = getattr(cnfg, group)
        for k, v in values.items():
            cur = getattr(target, k, None)
            if isinstance(v, dict) and isinstance(cur, type):
                for kk, vv in v.items():
                    setattr(cur, kk, copy.deepcopy(vv))
            else:
                setattr(target, k, copy.deepcopy(v))
    return cnfg


def config_mini_cheetah(Cnfg):
    """mini_cheetah_config.py:8 - mutates and returns the given Cfg tree."""
    return _apply_preset(Cnfg, "mini_cheetah")


def config_go1(Cnfg):
    """go1_config.py:8 - mutates and returns the given Cfg tree."""
    return _apply_preset(Cnfg, "go1")


def config_rough(Cnfg):
    """Mini Cheetah on a heightfield with measured heights (BASELINE.json configs[2]; not a shipped
    reference preset: SURVEY.md section 8 'Config provenance')."""
    Cnfg.terrain.mesh_type = "heightfield"
    Cnfg.terrain.measure_heights = True
    Cnfg.terrain.curriculum = True
    Cnfg.terrain.terrain_proportions = [0.1, 0.1, 0.35, 0.25, 0.2]
    Cnfg.env.num_observations = 42 + 17 * 11
    return Cnfg


# ----------------------------------------------------------------------------------------------
# freeze: Cfg -> resolved constants
# ----------------------------------------------------------------------------------------------
def f32(x):
    return float(np.float32(x))


def get_scale_shift(rng):
    """mini_gym/utils/math_utils.py:35-38."""
    return 2.0 / (rng[1] - rng[0]), (rng[1] + rng[0]) / 2.0


class TerrainInfo:
    """What the env core reads from the reference's Terrain object (utils/terrain.py:24-70,166-184)."""

    def __init__(self, cfg_terrain, heightsamples=None, eval_terrain=None, eval_heightsamples=None):
        """`eval_terrain`: the evaluation Cfg's terrain namespace (train / eval split, utils/terrain.py:43-57, 166-184):
        its tiles are appended below the training tiles - `eval_x_offset` rows further (pixels), `eval_rows_offset`
        tile rows - and `eval_env_origins` carries the shifted platform origins."""
        t = cfg_terrain
        self.mesh_type = t.mesh_type
        self.custom = t.mesh_type in ("heightfield", "trimesh")
        self.x_offset = 0
        self.eval_x_offset = self.eval_rows_offset = 0
        self.eval_env_origins = None
        if self.custom and eval_terrain is not None:
            train = TerrainInfo(t, heightsamples)
            ev = TerrainInfo(eval_terrain, eval_heightsamples)
            self.eval_x_offset, self.eval_rows_offset = train.tot_rows, int(t.num_rows)
            self.tot_rows, self.tot_cols = train.tot_rows + ev.tot_rows, max(train.tot_cols, ev.tot_cols)
            hs = np.zeros((self.tot_rows, self.tot_cols), dtype=np.int16)
            hs[:train.tot_rows, :train.tot_cols] = train.heightsamples
            hs[train.tot_rows:, :ev.tot_cols] = ev.heightsamples
            self.heightsamples = hs
            self.env_origins = train.env_origins
            self.eval_env_origins = ev.env_origins.copy()
            self.eval_env_origins[:, :, 0] += self.eval_x_offset * eval_terrain.horizontal_scale      # terrain.py:177
            return
        if self.custom:
            per_env_w = int(t.terrain_length / t.horizontal_scale)
            per_env_l = int(t.terrain_width / t.horizontal_scale)
            border = int(t.border_size / t.horizontal_scale)
            self.tot_cols = int(t.num_cols * per_env_w) + 2 * border
            self.tot_rows = int(t.num_rows * per_env_l) + 2 * border
            if heightsamples is None:
                heightsamples = np.zeros((self.tot_rows, self.tot_cols), dtype=np.int16)
            heightsamples = np.ascontiguousarray(heightsamples, dtype=np.int16)
            assert heightsamples.shape == (self.tot_rows, self.tot_cols), heightsamples.shape
            self.heightsamples = heightsamples
            origins = np.zeros((t.num_rows, t.num_cols, 3))
            for i in range(t.num_rows):
                for j in range(t.num_cols):
                    sx, ex = border + i * per_env_l, border + (i + 1) * per_env_l
                    sy, ey = border + j * per_env_w, border + (j + 1) * per_env_w
                    origins[i, j] = [(i + 0.5) * t.terrain_length, (j + 0.5) * t.terrain_width,
                                     np.max(heightsamples[sx:ex, sy:ey]) * t.vertical_scale]
            self.env_origins = origins
        else:
            self.tot_rows = self.tot_cols = 0
            self.heightsamples = None
            self.env_origins = None


class EnvParams:
    """Resolved constants of one env: python-side mirror of RlEnvCfg + the derived names."""

    def to_struct(self):
        c = _lib.RlEnvCfg()
        for name, _ in c._fields_:
            v = getattr(self, name)
            cur = getattr(c, name)
            if hasattr(cur, "__len__"):
                v = list(v)
                for i, x in enumerate(v):
                    cur[i] = x
            else:
                setattr(c, name, v)
        return c


CONTROL_TYPES = {"P": 0, "V": 1, "T": 2}
COMMAND_SUM_EXTRAS = ["lin_vel_raw", "ang_vel_raw", "lin_vel_residual", "ang_vel_residual", "ep_timesteps"]


def freeze_env_cfg(cfg, robot: RobotSpec = None, terrain: TerrainInfo = None, sim_dt=None,
                   custom_reward_names=(), upstream_order=False, eval_cfg=None):
    """Resolve a Cfg tree into EnvParams.

    Follows legged_robot.py _parse_cfg :1417-1429 (dt is decimation * float32(sim.dt); the
    episode / interval lengths are ceil(seconds / dt)), _prepare_reward_function :1074-1110 (drop
    zero scales, multiply by dt in double), _init_buffers :1012-1028 (PD gains by joint-name
    substring), _process_dof_props :501-515 (soft position limits), _get_noise_scale_vec :882-932.
    """
    p = EnvParams()
    robot = robot or robot_for_asset(cfg.asset.file)
    terrain = terrain or TerrainInfo(cfg.terrain)
    p.robot, p.terrain = robot, terrain
    sim_dt = f32(cfg.sim.dt if sim_dt is None else sim_dt)
    dt = cfg.control.decimation * sim_dt  # python double, like the reference's self.dt
    p.dt_double = dt
    # train / eval split (base_task.py:43-49): the evaluation envs follow the training envs
    p.num_train_envs = int(cfg.env.num_envs)
    p.num_eval_envs = 0 if eval_cfg is None else int(eval_cfg.env.num_envs)
    p.num_envs = p.num_train_envs + p.num_eval_envs
    p.num_bodies = robot.num_bodies
    p.num_actions = int(cfg.env.num_actions)
    p.num_obs = int(cfg.env.num_observations)
    p.measure_heights = int(bool(cfg.terrain.measure_heights))
    px, py = list(cfg.terrain.measured_points_x), list(cfg.terrain.measured_points_y)
    p.height_points = np.array([[x, y] for x in px for y in py], dtype=np.float32)  # meshgrid ij, x-major (:1461-1466)
    p.num_height_points = len(p.height_points) if p.measure_heights else 0
    if cfg.control.control_type not in CONTROL_TYPES:
        raise NameError("Unknown controller type: %s" % cfg.control.control_type)  # legged_robot.py:685
    p.control_type = CONTROL_TYPES[cfg.control.control_type]
    p.decimation = int(cfg.control.decimation)
    p.sim_dt, p.dt = sim_dt, f32(dt)
    p.action_scale = f32(cfg.control.action_scale)
    p.hip_scale_reduction = f32(cfg.control.hip_scale_reduction)
    p.clip_actions = f32(cfg.normalization.clip_actions)
    p.clip_obs = f32(cfg.normalization.clip_observations)
    # PD gains / default pose by joint-name substring
    p.dof_names = list(robot.dof_names)
    p.p_gains, p.d_gains, p.default_dof_pos = [], [], []
    for name in robot.dof_names:
        p.default_dof_pos.append(f32(cfg.init_state.default_joint_angles[name]))
        kp = kd = 0.0
        for key in cfg.control.stiffness.keys():
            if key in name:
                kp, kd = cfg.control.stiffness[key], cfg.control.damping[key]
        p.p_gains.append(f32(kp)); p.d_gains.append(f32(kd))
    p.torque_limits = [f32(x) for x in robot.dof_effort]
    p.dof_vel_limits = [f32(x) for x in robot.dof_velocity]
    p.dof_pos_lo, p.dof_pos_hi = [], []
    soft = cfg.rewards.soft_dof_pos_limit
    for lo, hi in zip(robot.dof_lower, robot.dof_upper):
        lo32, hi32 = np.float32(lo), np.float32(hi)      # fp32 tensor arithmetic (:512-515)
        m = (lo32 + hi32) / np.float32(2)
        r = hi32 - lo32
        p.dof_pos_lo.append(float(m - np.float32(0.5) * r * np.float32(soft)))
        p.dof_pos_hi.append(float(m + np.float32(0.5) * r * np.float32(soft)))
    # body index sets
    p.feet_idx = robot.bodies_matching(cfg.asset.foot_name)
    if len(p.feet_idx) != 4:
        raise ValueError("expected 4 feet bodies for foot_name=%r, got %d" % (cfg.asset.foot_name, len(p.feet_idx)))
    term = robot.bodies_matching(list(cfg.asset.terminate_after_contacts_on))
    pen = robot.bodies_matching(list(cfg.asset.penalize_contacts_on))
    maxb = _lib.DEFINES["RL_MAX_BODIES"]
    p.n_term_bodies, p.term_idx = len(term), term + [0] * (maxb - len(term))
    p.n_pen_bodies, p.pen_idx = len(pen), pen + [0] * (maxb - len(pen))
    # rewards
    scales = public_vars(cfg.rewards.scales)
    reward_scales = {}
    for key, scale in scales.items():
        if scale != 0:
            reward_scales[key] = scale * dt
    p.reward_scales = reward_scales
    p.reward_names = [n for n in reward_scales if n != "termination"]
    maxt = _lib.DEFINES["RL_MAX_TERMS"]
    p.term_id, p.term_scale = [], []
    p.custom_terms = []        # user-defined `_reward_<name>` methods: evaluated by the env class after the fused launch
    for n in p.reward_names:
        if n not in _lib.REWARD_TERM_IDS or n in custom_reward_names:
            if n in custom_reward_names:
                p.custom_terms.append((n, f32(reward_scales[n])))
                continue
            raise AttributeError("'LeggedRobot' object has no attribute '_reward_%s'" % n)  # :1093
        p.term_id.append(_lib.REWARD_TERM_IDS[n]); p.term_scale.append(f32(reward_scales[n]))
    fused_names = [n for n in p.reward_names if n not in dict(p.custom_terms)]
    p.n_terms = len(p.term_id)
    if p.n_terms > maxt:
        raise NotImplementedError("%d enabled reward terms > %d supported by the fused kernel" % (p.n_terms, maxt))
    p.term_id += [0] * (maxt - p.n_terms); p.term_scale += [0.0] * (maxt - p.n_terms)
    p.has_termination = int("termination" in reward_scales)
    p.termination_scale = f32(reward_scales.get("termination", 0.0))
    p.term_mask = 0
    for i in p.term_id[:p.n_terms]:
        p.term_mask |= 1 << i
    # fixed accumulator rows (rl_b200.h): term i -> row i, termination, total / extras
    p.sum_rows = {n: i for i, n in enumerate(fused_names)}
    if p.has_termination:
        p.sum_rows["termination"] = _lib.DEFINES["RL_MAX_TERMS"]
    p.sum_names = list(p.sum_rows.keys())
    p.only_positive_rewards = int(bool(cfg.rewards.only_positive_rewards))
    p.tracking_sigma = f32(cfg.rewards.tracking_sigma)
    p.tracking_sigma_yaw = f32(cfg.rewards.tracking_sigma_yaw)
    p.base_height_target = f32(cfg.rewards.base_height_target)
    p.soft_dof_vel_limit = f32(cfg.rewards.soft_dof_vel_limit)
    p.soft_torque_limit = f32(cfg.rewards.soft_torque_limit)
    p.max_contact_force = f32(cfg.rewards.max_contact_force)
    p.use_terminal_body_height = int(bool(cfg.rewards.use_terminal_body_height))
    p.terminal_body_height = f32(cfg.rewards.terminal_body_height)
    p.global_reference = int(bool(cfg.commands.global_reference))
    # observations
    e = cfg.env
    p.observe_command = int(bool(e.observe_command)); p.observe_vel = int(bool(e.observe_vel))
    p.observe_only_ang_vel = int(bool(e.observe_only_ang_vel)); p.observe_only_lin_vel = int(bool(e.observe_only_lin_vel))
    p.observe_yaw = int(bool(e.observe_yaw)); p.add_noise = int(bool(cfg.noise.add_noise))
    os_ = cfg.normalization.obs_scales
    p.obs_scale_lin_vel, p.obs_scale_ang_vel = f32(os_.lin_vel), f32(os_.ang_vel)
    p.obs_scale_dof_pos, p.obs_scale_dof_vel = f32(os_.dof_pos), f32(os_.dof_vel)
    p.obs_scale_height = f32(os_.height_measurements)
    p.commands_scale = [f32(os_.lin_vel), f32(os_.lin_vel), f32(os_.ang_vel)]
    # noise scale vector (:882-932), float32 arithmetic like torch.ones(k) * a * b * c
    ns, lvl = cfg.noise.noise_scales, cfg.noise.noise_level

    def seg(n, *factors):
        v = np.ones(n, dtype=np.float32)
        for f_ in factors:
            v = v * np.float32(f_)
        return v
    core = [seg(3, ns.gravity, lvl)]
    if p.observe_command:
        core.append(np.zeros(3, np.float32))
    core += [seg(12, ns.dof_pos, lvl, os_.dof_pos), seg(12, ns.dof_vel, lvl, os_.dof_vel),
             np.zeros(p.num_actions, np.float32)]
    vec = np.concatenate(core)
    if p.observe_vel:
        vec = np.concatenate([seg(3, ns.lin_vel, lvl, os_.lin_vel), seg(3, ns.ang_vel, lvl, os_.ang_vel), vec])
    if p.observe_only_lin_vel:
        vec = np.concatenate([seg(3, ns.lin_vel, lvl, os_.lin_vel), vec])
    if p.observe_yaw:
        vec = np.concatenate([vec, np.zeros(1, np.float32)])
    if p.measure_heights:
        vec = np.concatenate([vec, seg(len(p.height_points), ns.height_measurements, lvl, os_.height_measurements)])
    p.noise_scale_vec = vec.astype(np.float32)
    ncore = len(p.noise_scale_vec) - p.num_height_points
    maxc = _lib.DEFINES["RL_MAX_CORE_OBS"]
    if ncore > maxc or p.num_obs - p.num_height_points > maxc:
        raise NotImplementedError("%d non-height observation columns > %d" % (ncore, maxc))
    p.noise_scale_core = [float(x) for x in p.noise_scale_vec[:ncore]] + [0.0] * (maxc - ncore)
    p.noise_scale_height = float(p.noise_scale_vec[-1]) if p.measure_heights else 0.0
    # NOTE (reference quirk): observe_only_ang_vel adds 3 observation columns (:370-372) but no
    # noise entries (:882-932), so the reference itself fails to broadcast there; we require the
    # widths to agree.
    width = 3 + 3 * p.observe_command + 24 + p.num_actions + 6 * p.observe_vel + 3 * p.observe_only_ang_vel \
        + 3 * p.observe_only_lin_vel + p.observe_yaw + p.num_height_points
    if width != p.num_obs:
        raise ValueError("env.num_observations=%d but the enabled observation groups produce %d columns" % (p.num_obs, width))
    p.add_noise = int(bool(cfg.noise.add_noise))
    if len(p.noise_scale_vec) != p.num_obs and p.add_noise:       # (without noise the reference never uses the vector)
        raise ValueError("noise vector has %d columns for %d observations (reference :882-932 has the same gap)"
                         % (len(p.noise_scale_vec), p.num_obs))
    # privileged observations
    n = cfg.normalization
    pairs = [(n.friction_range, e.priv_observe_friction), (n.restitution_range, e.priv_observe_restitution),
             (n.added_mass_range, e.priv_observe_base_mass), (n.com_displacement_range, e.priv_observe_com_displacement),
             (n.motor_strength_range, e.priv_observe_motor_strength)]
    p.priv_scale, p.priv_shift = [], []
    for rng, on in pairs:
        sc, sh = get_scale_shift(rng)
        p.priv_scale.append(f32(sc if on else 0)); p.priv_shift.append(f32(sh))
    p.num_privileged_obs = int(e.num_privileged_obs)
    if p.num_privileged_obs != _lib.DEFINES["RL_PRIV_DIM"]:
        raise NotImplementedError("num_privileged_obs must be 18 (friction, restitution, payload, com[3], motor[12])")
    # teleport
    t = cfg.terrain
    p.teleport_robots = int(bool(t.teleport_robots))
    if p.teleport_robots and not terrain.custom:
        raise AttributeError("x_offset")  # reference quirk 5: plane + teleport fails at :774
    xo = int(terrain.x_offset * t.horizontal_scale)
    p.teleport_lo_x = f32(t.teleport_thresh + xo)
    p.teleport_hi_x = f32(t.terrain_length * t.num_rows - t.teleport_thresh + xo)
    p.teleport_shift_x = f32(t.terrain_length * (t.num_rows - 1))
    p.teleport_lo_y = f32(t.teleport_thresh)
    p.teleport_hi_y = f32(t.terrain_width * t.num_cols - t.teleport_thresh)
    p.teleport_shift_y = f32(t.terrain_width * (t.num_cols - 1))
    # heights
    if p.measure_heights and t.mesh_type == "none":
        raise NameError("Can't measure height with terrain mesh type 'none'")  # :1485
    p.heights_plane = int(t.mesh_type == "plane")
    p.border_size, p.horizontal_scale, p.vertical_scale = f32(t.border_size), f32(t.horizontal_scale), f32(t.vertical_scale)
    p.hf_rows, p.hf_cols = int(terrain.tot_rows), int(terrain.tot_cols)
    # episode clocks
    p.max_episode_length_f = float(np.ceil(cfg.env.episode_length_s / dt))
    p.max_episode_length = int(p.max_episode_length_f)
    dr = cfg.domain_rand
    p.rand_interval = int(np.ceil(dr.rand_interval_s / dt))
    p.push_interval = int(np.ceil(dr.push_interval_s / dt))
    p.randomize_motor_strength = int(bool(dr.randomize_motor_strength))
    p.randomize_Kp_factor = int(bool(dr.randomize_Kp_factor))
    p.randomize_Kd_factor = int(bool(dr.randomize_Kd_factor))
    p.motor_strength_lo_span = lo_span(dr.motor_strength_range)
    p.Kp_factor_lo_span = lo_span(dr.Kp_factor_range)
    p.Kd_factor_lo_span = lo_span(dr.Kd_factor_range)
    p.push_robots = int(bool(dr.push_robots))
    p.push_lo_span = [f32(-dr.max_push_vel_xy), f32(dr.max_push_vel_xy - (-dr.max_push_vel_xy))]
    # the evaluation range's values of the fields the reference reads from the Cfg it hands to _teleport_robots :576,
    # _push_robots :588 and _randomize_dof_props :593 (everything else in the step comes from the training Cfg)
    ec = cfg if eval_cfg is None else eval_cfg
    et, edr = ec.terrain, ec.domain_rand
    p.eval_teleport_robots = int(bool(et.teleport_robots))
    if eval_cfg is not None and p.eval_teleport_robots and not terrain.custom:
        raise AttributeError("x_offset")
    exo = int((terrain.eval_x_offset if eval_cfg is not None else terrain.x_offset) * et.horizontal_scale)
    p.eval_teleport_lo_x = f32(et.teleport_thresh + exo)
    p.eval_teleport_hi_x = f32(et.terrain_length * et.num_rows - et.teleport_thresh + exo)
    p.eval_teleport_shift_x = f32(et.terrain_length * (et.num_rows - 1))
    p.eval_teleport_lo_y = f32(et.teleport_thresh)
    p.eval_teleport_hi_y = f32(et.terrain_width * et.num_cols - et.teleport_thresh)
    p.eval_teleport_shift_y = f32(et.terrain_width * (et.num_cols - 1))
    p.eval_randomize_motor_strength = int(bool(edr.randomize_motor_strength))
    p.eval_randomize_Kp_factor = int(bool(edr.randomize_Kp_factor))
    p.eval_randomize_Kd_factor = int(bool(edr.randomize_Kd_factor))
    p.eval_motor_strength_lo_span = lo_span(edr.motor_strength_range)
    p.eval_Kp_factor_lo_span = lo_span(edr.Kp_factor_range)
    p.eval_Kd_factor_lo_span = lo_span(edr.Kd_factor_range)
    p.eval_push_robots = int(bool(edr.push_robots))
    p.eval_push_interval = int(np.ceil(edr.push_interval_s / dt))
    p.eval_push_lo_span = [f32(-edr.max_push_vel_xy), f32(edr.max_push_vel_xy - (-edr.max_push_vel_xy))]
    p.timeout_resets = int(bool(upstream_order))      # :197-198, commented out in this fork
    # command curriculum constants (:602-607)
    p.resample_interval = int(cfg.commands.resampling_time / dt)
    return p


def lo_span(rng):
    return [f32(rng[0]), f32(rng[1] - rng[0])]
