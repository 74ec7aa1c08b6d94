"""Named test configurations shared by the golden generator (reference side) and the tests
(oracle / product side).  `case_cfg_hook(case)` mutates a Cfg tree - the reference's params_proto
`Cfg` or ours, the attribute names are the same."""
import numpy as np


def case_cfg_hook(case):
    def hook(Cfg):
        if case == "go1_alt":
            Cfg.env.observe_vel = True
            Cfg.env.observe_yaw = True
            Cfg.env.num_observations = 42 + 6 + 1
            Cfg.commands.global_reference = False
            Cfg.domain_rand.push_robots = True
            Cfg.domain_rand.push_interval_s = 0.1
            Cfg.domain_rand.randomize_Kp_factor = True
            Cfg.domain_rand.randomize_Kd_factor = True
            Cfg.domain_rand.rand_interval_s = 0.06
            Cfg.rewards.only_positive_rewards = False
            Cfg.rewards.use_terminal_body_height = True
            Cfg.rewards.terminal_body_height = 0.29
            Cfg.rewards.soft_dof_vel_limit = 0.1
            Cfg.rewards.soft_torque_limit = 0.2
            Cfg.rewards.max_contact_force = 10.0
            s = Cfg.rewards.scales
            s.termination = -5.0
            s.lin_vel_z = 0.0
            s.ang_vel_xy = 0.0
            s.dof_acc = 0.0
            s.energy = -0.001
            s.energy_expenditure = -0.002
            s.dof_vel = -0.0003
            s.survival = 0.5
            s.dof_vel_limits = -0.7
            s.torque_limits = -0.02
            s.stumble = -0.3
            s.stand_still = -0.1
            s.feet_contact_forces = -0.01
        if case == "mc_only_lin":                     # legged_robot.py:374-376 (+ its noise columns :914-917)
            Cfg.env.observe_only_lin_vel = True
            Cfg.env.num_observations = 42 + 3
        if case == "mc_only_ang":                     # :370-372; _get_noise_scale_vec has no entry for this block, so the
            Cfg.env.observe_only_ang_vel = True       # reference only runs it with the observation noise off
            Cfg.env.num_observations = 42 + 3
            Cfg.noise.add_noise = False
        if case == "mc_rough":
            Cfg.terrain.num_rows = 2
            Cfg.terrain.num_cols = 2
            Cfg.terrain.border_size = 5
            Cfg.terrain.max_init_terrain_level = 1
    return hook


def eval_cfg_hook(ev):
    """The evaluation Cfg of the train / eval split case ("mc_eval": `ev` starts as a copy of the training Cfg, the
    shipped Mini Cheetah flat-trimesh preset).  Every field the reference reads from the evaluation Cfg differs from the
    training one: terrain extent / teleport band, pushes, DOF-property ranges, initial-position ranges."""
    ev.env.num_envs = 16
    ev.terrain.num_rows = 3
    ev.terrain.num_cols = 4
    ev.terrain.teleport_thresh = 1.5
    ev.terrain.x_init_range = 0.4
    ev.terrain.y_init_range = 0.7
    ev.terrain.x_init_offset = 0.3
    ev.terrain.y_init_offset = -0.2
    ev.domain_rand.push_robots = True
    ev.domain_rand.push_interval_s = 0.1
    ev.domain_rand.max_push_vel_xy = 0.7
    ev.domain_rand.randomize_motor_strength = False
    ev.domain_rand.randomize_Kp_factor = True
    ev.domain_rand.Kp_factor_range = [0.7, 1.1]
    ev.domain_rand.randomize_Kd_factor = True
    ev.domain_rand.Kd_factor_range = [0.6, 1.2]
    ev.domain_rand.added_mass_range = [0.0, 1.0]
    ev.commands.max_forward_curriculum = 2.0


def build_eval_case(num_train):
    """(cfg, eval_cfg, robot, terrain) of the train / eval split case for our side."""
    from rapid_locomotion_rl_b200 import config as C
    from rapid_locomotion_rl_b200.robots import robot_for_asset
    cfg, ev = C.new_cfg(), C.new_cfg()
    for c in (cfg, ev):
        C.config_mini_cheetah(c)
        c.env.record_video = False
    cfg.env.num_envs = num_train
    eval_cfg_hook(ev)
    terrain = C.TerrainInfo(cfg.terrain, eval_terrain=ev.terrain)
    return cfg, ev, robot_for_asset(cfg.asset.file), terrain


# mc_rough_full: BASELINE configs[2] at the shipped terrain size (10 x 20 tiles of 8 m, border as configured: the
# 1800 x 2600 int16 table); mc_rough is the same on a 2 x 2-tile table
ENV_CASES = ["mc_flat", "go1", "go1_alt", "mc_rough", "mc_rough_full", "mc_only_lin", "mc_only_ang"]


def build_case(case, num_envs):
    """(cfg, robot, terrain) for our side of a named case."""
    from rapid_locomotion_rl_b200 import config as C
    from rapid_locomotion_rl_b200.robots import robot_for_asset
    from rapid_locomotion_rl_b200.sim import synthetic_heightfield
    cfg = C.new_cfg()
    if case.startswith("go1"):
        C.config_go1(cfg)
    else:
        C.config_mini_cheetah(cfg)
    if case.startswith("mc_rough"):
        C.config_rough(cfg)
    case_cfg_hook(case)(cfg)
    cfg.env.num_envs = num_envs
    cfg.env.record_video = False
    hs = None
    if case.startswith("mc_rough"):
        probe = C.TerrainInfo(cfg.terrain)
        hs = synthetic_heightfield(probe.tot_rows, probe.tot_cols, seed=3)
    terrain = C.TerrainInfo(cfg.terrain, hs)
    return cfg, robot_for_asset(cfg.asset.file), terrain


# ---- learner fixtures: portable, seed-generated inputs (numpy), so goldens only store reference OUTPUTS ----
LEARNER_SHAPES = {
    "std": (12,),
    "env_factor_encoder.0": (256, 18), "env_factor_encoder.2": (128, 256), "env_factor_encoder.4": (18, 128),
    "adaptation_module.0": (256, 630), "adaptation_module.2": (32, 256), "adaptation_module.4": (18, 32),
    "actor_body.0": (512, 60), "actor_body.2": (256, 512), "actor_body.4": (128, 256), "actor_body.6": (12, 128),
    "critic_body.0": (512, 60), "critic_body.2": (256, 512), "critic_body.4": (128, 256), "critic_body.6": (1, 128),
}


def learner_weights(seed=7):
    """state_dict-shaped float32 numpy weights, nn.Linear-like uniform(-1/sqrt(in), 1/sqrt(in))."""
    rng = np.random.RandomState(seed)
    sd = {}
    for name, shape in LEARNER_SHAPES.items():
        if name == "std":
            sd["std"] = np.ones(12, np.float32)
            continue
        bound = 1.0 / np.sqrt(shape[1])
        sd[name + ".weight"] = rng.uniform(-bound, bound, shape).astype(np.float32)
        sd[name + ".bias"] = rng.uniform(-bound, bound, shape[0]).astype(np.float32)
    for k in list(sd):
        if k.startswith("env_factor_encoder."):
            sd["encoder." + k[len("env_factor_encoder."):]] = sd[k]
    return sd


def learner_rollout_inputs(n_envs, n_steps, seed=9):
    """Per-step env outputs fed to PPO.act / process_env_step."""
    rng = np.random.RandomState(seed)
    steps = []
    for _ in range(n_steps):
        steps.append(dict(obs=rng.randn(n_envs, 42).astype(np.float32), priv=rng.uniform(-1, 1, (n_envs, 18)).astype(np.float32),
                          hist=rng.randn(n_envs, 630).astype(np.float32), rew=(rng.randn(n_envs) * 0.05).astype(np.float32),
                          done=rng.rand(n_envs) < 0.05))
    last = dict(obs=rng.randn(n_envs, 42).astype(np.float32), priv=rng.uniform(-1, 1, (n_envs, 18)).astype(np.float32))
    perm = rng.permutation(n_envs * n_steps).astype(np.int64)
    return steps, last, perm


def tensor_digest(a):
    """Compact exact fingerprint of an array: float64 sum, abs-sum and 64 strided samples."""
    a = np.asarray(a, dtype=np.float32).reshape(-1)
    idx = np.linspace(0, a.size - 1, min(64, a.size)).astype(np.int64)
    return np.concatenate([[a.astype(np.float64).sum(), np.abs(a.astype(np.float64)).sum()], a[idx].astype(np.float64)])


def hlp_weights(seed=17):
    """state_dict of the reference's high_level_policy ActorCritic with USE_LATENT = False: bodies + std only, first layers
    [512, 42] (high_level_policy/ppo/actor_critic.py:86-110)."""
    rng = np.random.RandomState(seed)
    sd = {"std": np.ones(12, np.float32)}
    for name, shape in LEARNER_SHAPES.items():
        if not name.startswith(("actor_body", "critic_body")):
            continue
        if name.endswith(".0"):
            shape = (shape[0], 42)
        bound = 1.0 / np.sqrt(shape[1])
        sd[name + ".weight"] = rng.uniform(-bound, bound, shape).astype(np.float32)
        sd[name + ".bias"] = rng.uniform(-bound, bound, shape[0]).astype(np.float32)
    return sd
