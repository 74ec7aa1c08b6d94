"""Fused env kernels (csrc/env_step.cu, env_reset.cu, gac.cu, history.cu) through the drop-in
LeggedRobot API, against (a) the reference's golden fixtures, (b) the pinned oracle at the
BASELINE sizes, (c) size-independent properties at 32768 envs.

Tolerances (BASELINE.json north_star): resets, terminations, counters, contact flags, terrain
levels, curriculum bins and weights bit-exact; torques / rewards / observations / accumulators
within 1e-5 relative (atol 2e-6 covers cancellation in sums of signed terms).
"""
import os

import numpy as np
import pytest
import torch

import statekit
from cases import ENV_CASES, build_case

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 2e-6
DEV = "cuda:0"


def make_env(case, n, **kw):
    from rapid_locomotion_rl_b200.envs import LeggedRobot
    cfg, robot, terrain = build_case(case, n)
    return LeggedRobot(cfg, sim_device=DEV, headless=True, terrain=terrain, **kw), cfg, robot, terrain


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def sub(g, prefix):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def load(golden_dir, case):
    return dict(np.load(os.path.join(golden_dir, "env_%s.npz" % case), allow_pickle=False))


def check_step(env, obs, priv, rew, reset, want_obs, want_priv, want_rew, want_reset, label):
    torch.cuda.synchronize()
    np.testing.assert_allclose(obs.cpu().numpy(), want_obs, rtol=RTOL, atol=ATOL, err_msg=label + " obs")
    np.testing.assert_allclose(priv.cpu().numpy(), want_priv, rtol=RTOL, atol=ATOL, err_msg=label + " priv")
    np.testing.assert_allclose(rew.cpu().numpy(), want_rew, rtol=RTOL, atol=ATOL, err_msg=label + " rew")
    assert np.array_equal(reset.cpu().numpy(), want_reset), label + " reset_buf must be bit-exact"


@pytest.mark.parametrize("case", ENV_CASES)
def test_step_vs_golden(golden_dir, case):
    g = load(golden_dir, case)
    env, cfg, robot, terrain = make_env(case, 48)
    for s in range(3):
        pre = "step%d/" % s
        statekit.apply_to_product(env, sub(g, pre + "before/"))
        env._inject = dict(noise_u=cu(g[pre + "noise_u"]), dr_u=cu(g[pre + "dr_u"]), push_u=cu(g[pre + "push_u"]))
        obs, priv, rew, reset, _ = env.step(cu(g[pre + "actions"]))
        check_step(env, obs, priv, rew, reset, g[pre + "obs"], g[pre + "priv"], g[pre + "rew"], g[pre + "reset"],
                   "%s step %d" % (case, s))
        if case == "mc_rough":
            np.testing.assert_allclose(env.measured_heights.cpu().numpy(), g[pre + "measured_heights"], rtol=0, atol=0)
        got = statekit.state_from_product(env)
        statekit.assert_state_close(got, sub(g, pre + "after/"), RTOL, ATOL, skip=("contact_forces",),
                                    label="%s step %d" % (case, s))


def random_persistent_state(rng, n, sum_names):
    s = dict(
        commands=np.concatenate([rng.uniform(-1, 1, (n, 3)), np.zeros((n, 1))], 1).astype(np.float32),
        last_actions=rng.normal(0, 1, (n, 12)).astype(np.float32),
        last_dof_vel=rng.normal(0, 3, (n, 12)).astype(np.float32),
        motor_strengths=np.repeat(rng.uniform(0.9, 1.1, (n, 1)), 12, 1).astype(np.float32),
        Kp_factors=np.repeat(rng.uniform(0.8, 1.3, (n, 1)), 12, 1).astype(np.float32),
        Kd_factors=np.repeat(rng.uniform(0.5, 1.5, (n, 1)), 12, 1).astype(np.float32),
        friction_coeffs=rng.uniform(0.05, 4.5, n).astype(np.float32),
        restitutions=rng.uniform(0, 1, n).astype(np.float32), payloads=rng.uniform(-1, 3, n).astype(np.float32),
        com_displacements=rng.uniform(-0.1, 0.1, (n, 3)).astype(np.float32),
        feet_air_time=(rng.uniform(0, 0.6, (n, 4)) * (rng.random((n, 4)) < 0.7)).astype(np.float32),
        last_contacts=rng.random((n, 4)) < 0.3, episode_length_buf=rng.integers(0, 1001, n).astype(np.int64))
    s["commands"][: n // 8, :2] *= 0.05
    for k in sum_names:
        s["episode_sums/" + k] = rng.normal(0, 1, n).astype(np.float32)
    return s


@pytest.mark.parametrize("case,n", [("mc_flat", 4000), ("go1", 4133), ("mc_rough", 1000), ("go1_alt", 777),
                                    ("mc_rough_full", 4000), ("mc_only_lin", 515)])
def test_step_vs_oracle_at_size(case, n):
    """Same seeded inputs through the pinned oracle (CPU fp32) and the kernel, two consecutive steps."""
    from oracle.env_oracle import OracleEnv
    from rapid_locomotion_rl_b200.sim import synthetic_state
    env, cfg, robot, terrain = make_env(case, n)
    o = OracleEnv(cfg, robot, terrain)
    rng = np.random.default_rng(n)
    st = random_persistent_state(rng, n, list(o.episode_sums.keys()))
    for k in o.command_sums:
        st["command_sums/" + k] = rng.normal(0, 1, n).astype(np.float32)
    for k in ("env_origins", "terrain_levels", "terrain_types"):   # drawn at construction by the product
        st[k] = getattr(env, k).cpu().numpy()
    p = env.params
    for step in range(2):
        sim = synthetic_state(1000 + step, n, robot.num_bodies, 12, np.float32(p.default_dof_pos), p.feet_idx,
                              p.term_idx[:p.n_term_bodies], z0=0.3, teleport_band_frac=0.05)
        if case == "mc_rough":   # keep robots over the small 2x2 test terrain
            sim["root_states"][:, :2] = rng.uniform(0.5, 15.5, (n, 2))
        st.update(sim)
        statekit.apply_to_oracle(o, st)
        statekit.apply_to_product(env, st)
        actions = rng.normal(0, 1, (n, 12)).astype(np.float32)
        noise_u = rng.random((n, p.num_obs)).astype(np.float32)
        dr_u = rng.random((3, n)).astype(np.float32); push_u = rng.random((2, n)).astype(np.float32)
        env._inject = dict(noise_u=cu(noise_u), dr_u=cu(dr_u), push_u=cu(push_u))
        obs, priv, rew, reset, _ = env.step(cu(actions))
        oo, op, orr, ors = o.step(torch.from_numpy(actions), noise_u=torch.from_numpy(noise_u),
                                  dr_u=torch.from_numpy(dr_u), push_u=torch.from_numpy(push_u))
        check_step(env, obs, priv, rew, reset, oo.numpy(), op.numpy(), orr.numpy(), ors.numpy(), "%s n=%d step %d" % (case, n, step))
        want = statekit.state_from_oracle(o)
        statekit.assert_state_close(statekit.state_from_product(env), want, RTOL, ATOL, label="%s n=%d" % (case, n))
        if case.startswith("mc_rough"):
            assert np.array_equal(env.measured_heights.cpu().numpy(), o.measured_heights.numpy()), "heights are exact"
        st = want  # carry the state into the next step


@pytest.mark.parametrize("case", ENV_CASES)
def test_reset_vs_golden(golden_dir, case):
    g = load(golden_dir, case)
    env, cfg, robot, terrain = make_env(case, 48)
    statekit.apply_to_product(env, sub(g, "reset/before/"))
    env._reset_u8.copy_(cu(g["step2/reset"].astype(np.uint8)))
    env._inject = dict(reset_dr_u=cu(g["reset/dr_u"]), init_u=cu(g["reset/init_u"]), level_u=cu(g["reset/level_u"]))
    env.common_step_counter = 7  # not a multiple of max_episode_length: the uniform curriculum stays put
    env.reset_idx(cu(g["reset/ids"]))
    torch.cuda.synchronize()
    want = sub(g, "reset/after/")
    got = statekit.state_from_product(env)
    statekit.assert_state_close(got, want, RTOL, ATOL,
                                exact_keys=("root_states", "dof_state", "last_actions", "last_dof_vel", "feet_air_time",
                                            "env_origins", "motor_strengths", "Kp_factors", "Kd_factors"),
                                skip=("contact_forces", "torques", "base_lin_vel", "base_ang_vel", "projected_gravity",
                                      "last_root_vel", "joint_pos_target"), label=case + " reset")
    assert np.array_equal(env.reset_buf.cpu().numpy(), g["reset/reset_buf"])
    for k, v in sub(g, "reset/extras/").items():
        np.testing.assert_allclose(env.extras["train/episode"][k].item(), v, rtol=RTOL, atol=ATOL, err_msg=k)


def test_reset_mask_and_empty():
    """Edge cases of reset_idx: empty id list is a no-op (:238); all envs; history rows are zeroed."""
    from rapid_locomotion_rl_b200.envs import HistoryWrapper, VelocityTrackingEasyEnv
    cfg, robot, terrain = build_case("mc_flat", 300)
    env = HistoryWrapper(VelocityTrackingEasyEnv(sim_device=DEV, headless=True, cfg=cfg, terrain=terrain))
    env.env.episode_length_buf.fill_(5)
    env.reset_idx(torch.zeros(0, dtype=torch.long, device=DEV))
    assert (env.env.episode_length_buf == 5).all()
    for _ in range(3):
        env.step(torch.randn(300, 12, device=DEV))
    assert env.obs_history.abs().sum() > 0
    ids = torch.tensor([0, 17, 299], device=DEV)
    env.reset_idx(ids)
    torch.cuda.synchronize()
    assert (env.obs_history[ids] == 0).all() and (env.env.episode_length_buf[ids] == 0).all()
    assert env.obs_history[1].abs().sum() > 0 and env.env.episode_length_buf[1] == 3 + 5
    env.reset_idx(torch.arange(300, device=DEV))
    assert (env.env.episode_length_buf == 0).all() and (env.env.last_actions == 0).all()


@pytest.mark.parametrize("case", ["mc_flat", "go1"])
def test_resample_vs_golden(golden_dir, case):
    """GAC: weights (float64) and bins bit-exact, commands bit-exact given the reference's MT19937 uniforms."""
    g = load(golden_dir, case)
    env, cfg, robot, terrain = make_env(case, 48)
    assert env.curriculum.weights.sum() == 30.0
    for rnd in range(2):
        pre = "resample%d/" % rnd
        env.curriculum.weights = g[pre + "before/weights"]
        env._env_command_bins.copy_(cu(g[pre + "before/bins"]))
        env.commands.copy_(cu(g[pre + "before/commands"]))
        for k, row in env.command_sums.items():
            row.copy_(cu(g[pre + "before/command_sums/" + k]))
        env._inject = dict(gac_u_bin=cu(g[pre + "u_bin"]), gac_u_cell=cu(g[pre + "u_cell"]))
        env._resample_commands(cu(g[pre + "ids"]))
        torch.cuda.synchronize()
        assert np.array_equal(env.curriculum.weights, g[pre + "after/weights"]), "weights must be bit-exact"
        assert np.array_equal(env.env_command_bins.cpu().numpy(), g[pre + "after/bins"]), "bins must be bit-exact"
        assert np.array_equal(env.commands.cpu().numpy(), g[pre + "after/commands"])
        for k, row in env.command_sums.items():
            assert np.array_equal(row.cpu().numpy(), g[pre + "after/command_sums/" + k]), k
        assert int(env.curriculum.hit_count.abs().sum()) == 0, "workspace re-armed"


def test_resample_numpy_replay_matches_reference_stream(golden_dir):
    """gac_rng='numpy' replays RandomState(curriculum_seed) in the reference's draw order."""
    g = load(golden_dir, "mc_flat")
    env, cfg, robot, terrain = make_env("mc_flat", 48, gac_rng="numpy")
    env._resample_commands(torch.arange(48, device=DEV))   # the initial all-env draw the fixture also made
    pre = "resample0/"
    env.curriculum.weights = g[pre + "before/weights"]
    env._env_command_bins.copy_(cu(g[pre + "before/bins"]))
    for k, row in env.command_sums.items():
        row.copy_(cu(g[pre + "before/command_sums/" + k]))
    env._resample_commands(cu(g[pre + "ids"]))
    torch.cuda.synchronize()
    assert np.array_equal(env.env_command_bins.cpu().numpy(), g[pre + "after/bins"])
    ids = g[pre + "ids"]
    assert np.array_equal(env.commands.cpu().numpy()[ids], g[pre + "after/commands"][ids])


def test_resample_statistics_philox():
    """Device RNG path: bins follow the weight distribution (chi-square) and commands fall inside their cell."""
    env, cfg, robot, terrain = make_env("mc_flat", 32768)
    env._resample_commands(torch.arange(32768, device=DEV))
    torch.cuda.synchronize()
    bins = env.env_command_bins.cpu().numpy()
    w = env.curriculum.weights
    live = np.nonzero(w)[0]
    assert set(np.unique(bins)) <= set(live)
    counts = np.bincount(bins, minlength=len(w))[live]
    expected = 32768 / len(live)
    chi2 = ((counts - expected) ** 2 / expected).sum()
    assert chi2 < 2.5 * len(live), chi2     # 30 live bins: mean 29, sd 7.6
    grid = env.curriculum.grid.T[bins]
    half = np.array(list(env.curriculum.bin_sizes.values())) / 2
    cmd = env.commands.cpu().numpy()
    big = np.linalg.norm(cmd[:, :2], axis=1) > 0
    assert (np.abs(cmd[big, :3] - grid[big]) <= half + 1e-6).all()
    assert (np.linalg.norm(cmd[:, :2], axis=1)[big] > 0.2).all()


def test_history_ring_matches_shift_concat():
    from rapid_locomotion_rl_b200.envs import HistoryWrapper, VelocityTrackingEasyEnv
    cfg, robot, terrain = build_case("mc_flat", 257)
    env = HistoryWrapper(VelocityTrackingEasyEnv(sim_device=DEV, headless=True, cfg=cfg, terrain=terrain))
    ref = torch.zeros(257, 15 * 42, device=DEV)
    for i in range(40):
        out, rew, done, info = env.step(torch.randn(257, 12, device=DEV))
        ref = torch.cat((ref[:, 42:], out["obs"]), dim=-1)   # history_wrapper.py:23
        assert torch.equal(out["obs_history"], ref), "step %d" % i
    assert info["privileged_obs"] is out["privileged_obs"]
    assert info["joint_pos"].shape == (257, 12)    # lazy numpy extras


def test_philox_noise_statistics_and_determinism():
    env, cfg, robot, terrain = make_env("mc_flat", 8192, seed=123)
    a = torch.zeros(8192, 12, device=DEV)
    env.common_step_counter = 10
    o1 = env.step(a)[0].clone()
    env.episode_length_buf.zero_(); env.common_step_counter = 10
    o2 = env.step(a)[0].clone()
    assert torch.equal(o1, o2), "same (seed, step) -> same noise"
    o3 = env.step(a)[0].clone()
    assert not torch.equal(o1, o3)
    # gravity columns: sim state is the upright default => clean value (0,0,-1), noise U(-0.05, 0.05)
    nz = (o1[:, :3] - torch.tensor([0.0, 0.0, -1.0], device=DEV)) / 0.05
    assert nz.abs().max() <= 1.0 + 1e-6
    assert abs(nz.mean().item()) < 0.02 and abs(nz.var().item() - 1.0 / 3.0) < 0.02
    assert (o1[:, 30:42] == 0).all()   # action columns carry no noise


def test_full_size_properties():
    """32768 envs (BASELINE configs[4] per GPU): properties checked with plain torch on the device."""
    from rapid_locomotion_rl_b200.sim import synthetic_state
    n = 32768
    env, cfg, robot, terrain = make_env("mc_flat", n)
    p = env.params
    st = synthetic_state(0, n, robot.num_bodies, 12, np.float32(p.default_dof_pos), p.feet_idx, p.term_idx[:p.n_term_bodies])
    statekit.apply_to_product(env, st)
    ep0 = torch.randint(0, 1001, (n,), device=DEV)
    env.episode_length_buf.copy_(ep0)
    actions = torch.randn(n, 12, device=DEV) * 60
    obs, priv, rew, reset, _ = env.step(actions)
    torch.cuda.synchronize()
    assert torch.equal(env.episode_length_buf, ep0 + 1)
    cf = env.contact_forces[:, env.termination_contact_indices, :]
    assert torch.equal(reset, torch.any(torch.norm(cf, dim=-1) > 1.0, dim=1))
    assert torch.equal(obs[:, 30:42], torch.clip(actions, -100, 100))
    assert torch.equal(env.last_actions, torch.clip(actions, -100, 100))
    assert torch.equal(env.last_dof_vel, env.dof_vel)
    assert (env.torques.abs() <= env.torque_limits + 0).all()
    assert (rew >= 0).all()                                   # only_positive_rewards
    assert (obs.abs() <= 100).all() and (priv.abs() <= 100).all()
    # standalone torque entry == fused torques
    t_fused = env.torques.clone()
    env.step(actions)  # Kp/Kd/motor unchanged unless re-drawn; compare on envs that were not re-drawn
    keep = (env.episode_length_buf % p.rand_interval != 0) & ((env.episode_length_buf - 1) % p.rand_interval != 0)
    t_alone = env._compute_torques(actions).clone()
    torch.cuda.synchronize()
    assert torch.equal(t_alone[keep], t_fused[keep])


def test_api_errors():
    from rapid_locomotion_rl_b200 import _lib
    from rapid_locomotion_rl_b200.envs import LeggedRobot
    cfg, robot, terrain = build_case("mc_flat", 16)
    cfg.control.control_type = "X"
    with pytest.raises(NameError):
        LeggedRobot(cfg, sim_device=DEV, terrain=terrain)
    cfg, robot, terrain = build_case("mc_flat", 16)
    cfg.rewards.scales.feet_stumble = -1.0      # scale name without a _reward_ method (quirk 15)
    with pytest.raises(AttributeError):
        LeggedRobot(cfg, sim_device=DEV, terrain=terrain)
    cfg, robot, terrain = build_case("go1", 16)
    cfg.terrain.teleport_robots = True          # plane + teleport: AttributeError x_offset (quirk 5)
    with pytest.raises(AttributeError):
        LeggedRobot(cfg, sim_device=DEV)
    cfg, robot, terrain = build_case("mc_flat", 16)
    env = LeggedRobot(cfg, sim_device=DEV, terrain=terrain)
    with pytest.raises(ValueError):
        env.step(torch.zeros(15, 12, device=DEV))
    with pytest.raises(_lib.RlError):
        LeggedRobot(cfg, sim_device="cpu", terrain=terrain)


def test_host_io_zero_copy_matches_device_step():
    """bind_host_io / step_host (simulator tensors and outputs in pinned host memory, read and written by
    the kernel in place) gives exactly what the device-resident step gives."""
    import numpy as np
    from cases import build_case
    from rapid_locomotion_rl_b200.envs import LeggedRobot
    from rapid_locomotion_rl_b200.sim import synthetic_state
    n = 4133
    outs = []
    for host in (False, True):
        cfg, robot, terrain = build_case("mc_flat", n)
        env = LeggedRobot(cfg, sim_device="cuda:0", headless=True, terrain=terrain, seed=11)
        p = env.params
        st = synthetic_state(5, n, robot.num_bodies, 12, np.float32(p.default_dof_pos), p.feet_idx, p.term_idx[:p.n_term_bodies])
        actions = torch.from_numpy(np.random.default_rng(1).normal(0, 1, (n, 12)).astype(np.float32))
        env.commands[:, :3] = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, (n, 3)).astype(np.float32)).cuda()
        if not host:
            env.sim.root_states.copy_(torch.from_numpy(st["root_states"]))
            env.sim.dof_state.copy_(torch.from_numpy(st["dof_state"]).view(-1, 2))
            env.sim.contact_forces.copy_(torch.from_numpy(st["contact_forces"]).view(-1, 3))
            for _ in range(3):
                obs, priv, rew, reset, _x = env.step(actions.cuda())
            torch.cuda.synchronize()
            outs.append([t.cpu().clone() for t in (obs, priv, rew, reset.to(torch.uint8), env.torques, env.episode_sums["total"])])
        else:
            pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            h_root, h_dof, h_con = pin(st["root_states"]), pin(st["dof_state"].reshape(-1, 2)), pin(st["contact_forces"].reshape(-1, 3))
            h_obs = torch.zeros(n, env.num_obs).pin_memory(); h_priv = torch.zeros(n, 18).pin_memory()
            h_rew = torch.zeros(n).pin_memory(); h_reset = torch.zeros(n, dtype=torch.uint8).pin_memory()
            env.bind_host_io(h_root, h_dof, h_con, h_obs, h_priv, h_rew, h_reset)
            h_act = actions.pin_memory()
            for _ in range(3):
                env.step_host(h_act)
            torch.cuda.synchronize()
            outs.append([t.clone() for t in (h_obs, h_priv, h_rew, h_reset)] + [env.torques.cpu(), env.episode_sums["total"].cpu()])
            with pytest.raises(ValueError):
                env.step_host(actions)          # not pinned
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("case", ["mc_flat", "go1"])
def test_packed_io_matches_device_step(case):
    """pack_io (simulator rows + actions as views of one device block, outputs of another: one H2D and one D2H copy
    per step) gives exactly what the step on separately allocated tensors gives."""
    import numpy as np
    from cases import build_case
    from rapid_locomotion_rl_b200.envs import LeggedRobot
    from rapid_locomotion_rl_b200.sim import synthetic_state
    n = 4132
    outs = []
    for packed in (False, True):
        cfg, robot, terrain = build_case(case, n)
        env = LeggedRobot(cfg, sim_device="cuda:0", headless=True, terrain=terrain, seed=11)
        p = env.params
        st = synthetic_state(5, n, robot.num_bodies, 12, np.float32(p.default_dof_pos), p.feet_idx, p.term_idx[:p.n_term_bodies])
        actions = torch.from_numpy(np.random.default_rng(1).normal(0, 1, (n, 12)).astype(np.float32))
        env.commands[:, :3] = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, (n, 3)).astype(np.float32)).cuda()
        if not packed:
            for _ in range(3):                  # the simulator's rows are re-sent every step in both variants
                env.sim.root_states.copy_(torch.from_numpy(st["root_states"]))
                env.sim.dof_state.copy_(torch.from_numpy(st["dof_state"]).view(-1, 2))
                env.sim.contact_forces.copy_(torch.from_numpy(st["contact_forces"]).view(-1, 3))
                obs, priv, rew, reset, _x = env.step(actions.cuda())
            torch.cuda.synchronize()
            outs.append([t.cpu().clone() for t in (obs, priv, rew, reset.to(torch.uint8), env.torques, env.episode_sums["total"])])
        else:
            d_in, d_out, lay_in, lay_out = env.pack_io()
            assert d_in.numel() == n * (13 + 24 + 3 * robot.num_bodies + 12) * 4
            h_in = torch.empty(d_in.numel(), dtype=torch.uint8).pin_memory()
            h_out = torch.empty(d_out.numel(), dtype=torch.uint8).pin_memory()
            hv = LeggedRobot.host_views(h_in, lay_in)
            hv["root_states"].copy_(torch.from_numpy(st["root_states"]))
            hv["dof_state"].copy_(torch.from_numpy(st["dof_state"]).view(-1, 2))
            hv["contact_forces"].copy_(torch.from_numpy(st["contact_forces"]).view(-1, 3))
            hv["actions"].copy_(actions)
            for _ in range(3):
                d_in.copy_(h_in, non_blocking=True)
                env.step(env.packed_actions)
                h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            ho = LeggedRobot.host_views(h_out, lay_out)
            outs.append([ho["obs"].clone(), ho["priv"].clone(), ho["rew"].clone(), ho["reset"].clone(),
                         env.torques.cpu(), env.episode_sums["total"].cpu()])
            cfg2, _r2, terrain2 = build_case("mc_flat", 50)
            env2 = LeggedRobot(cfg2, sim_device="cuda:0", headless=True, terrain=terrain2, seed=1)
            with pytest.raises(ValueError):
                env2.pack_io()                  # sections would not be 16-byte aligned
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("case,n,rows_mode", [("mc_flat", 4000, 1), ("go1", 4128, 1), ("mc_flat", 32768, 3), ("mc_flat", 32768, 2),
                                              ("go1", 4128, 2), ("mc_flat", 96, 2), ("mc_flat", 65536, 2),
                                              ("mc_flat", 4000, 4), ("go1", 4128, 4), ("mc_flat", 32768, 4), ("mc_flat", 4000, 3)])
def test_rows_kernel_matches_quad_kernel(case, n, rows_mode):
    """The all-TMA kernel (csrc/env_step_rows.cu, packed state blocks) and the one-warp-per-leg kernel it replaces give
    identical bits for every output and every piece of state: fused step and post-physics entry, Philox noise (no
    injection), three consecutive steps (the second re-draws Kp / Kd / motor strength for a third of the envs)."""
    from rapid_locomotion_rl_b200 import _lib
    from rapid_locomotion_rl_b200.sim import synthetic_state
    lib = _lib.lib()
    results = []
    # rows_mode: 1 = automatic, 2 = persistent two-buffer variant (tile queue), 3 = one tile per CTA (four warps), 4 = the
    # 16-warp wide variant (what 'automatic' picks for grids of at most two CTAs per SM)
    for mode in (rows_mode, 0):
        prev = lib.rl_debug_env_rows(mode)
        try:
            env, cfg, robot, terrain = make_env(case, n)
            p = env.params
            rng = np.random.default_rng(7)
            st = random_persistent_state(rng, n, list(env.episode_sums.keys()))
            st["episode_length_buf"][::3] = p.rand_interval - 2        # re-draw on the second step
            for k in env.command_sums:
                st["command_sums/" + k] = rng.normal(0, 1, n).astype(np.float32)
            for k in ("env_origins", "terrain_levels", "terrain_types"):
                st[k] = getattr(env, k).cpu().numpy()
            st.update(synthetic_state(3, n, robot.num_bodies, 12, np.float32(p.default_dof_pos), p.feet_idx,
                                      p.term_idx[:p.n_term_bodies], z0=0.3, teleport_band_frac=0.05))
            statekit.apply_to_product(env, st)
            actions = cu(rng.normal(0, 1, (n, 12)).astype(np.float32))
            outs = []
            for s in range(3):
                if s == 2:                      # the gymapi-shaped entry pair
                    env._compute_torques(actions)
                    env.post_physics_step(actions)
                    o = (env.obs_buf, env.privileged_obs_buf, env.rew_buf, env.reset_buf)
                else:
                    o = env.step(actions)[:4]
                torch.cuda.synchronize()
                outs.append([t.clone() for t in o] + [env.torques.clone()])
            state = statekit.state_from_product(env)
            results.append((outs, state))
        finally:
            lib.rl_debug_env_rows(prev)
    (oa, sa), (ob, sb) = results
    def where(a, b):
        a, b = np.asarray(a), np.asarray(b)
        bad = np.argwhere(a != b)
        rows = np.unique(bad[:, 0])
        return "%d elements in %d envs differ; first envs %s (mod 32: %s), first columns %s" % (
            len(bad), len(rows), rows[:8].tolist(), (rows[:8] % 32).tolist(), bad[:8, 1:].tolist())
    names = ("obs", "priv", "rew", "reset", "torques")
    for s, (xa, xb) in enumerate(zip(oa, ob)):
        for i, (a, b) in enumerate(zip(xa, xb)):
            assert torch.equal(a, b), "step %d %s: %s" % (s, names[i], where(a.cpu().numpy(), b.cpu().numpy()))
    for k in sa:
        assert np.array_equal(np.asarray(sa[k]), np.asarray(sb[k])), "state %r: %s" % (k, where(sa[k], sb[k]))


@pytest.mark.parametrize("case,n", [("mc_rough", 1000), ("mc_rough_full", 4000)])
def test_heights_prepass_matches_in_kernel_sampling(case, n, monkeypatch):
    """The terrain heights sampled by the pre-pass launch (csrc/heights.cu, one warp per env; the default) and inside the
    step kernel (RL_ENV_HEIGHTS_PREPASS=0) give identical bits: observations (Philox noise, no injection), rewards,
    measured_heights, every piece of state - over three steps with envs inside the teleport band."""
    from rapid_locomotion_rl_b200.sim import synthetic_state
    results = []
    for mode in ("1", "0"):
        monkeypatch.setenv("RL_ENV_HEIGHTS_PREPASS", mode)
        env, cfg, robot, terrain = make_env(case, n)
        assert (getattr(env, "_height_mean", None) is not None) == (mode == "1")
        p = env.params
        rng = np.random.default_rng(11)
        st = random_persistent_state(rng, n, list(env.episode_sums.keys()))
        for k in env.command_sums:
            st["command_sums/" + k] = rng.normal(0, 1, n).astype(np.float32)
        for k in ("env_origins", "terrain_levels", "terrain_types"):
            st[k] = getattr(env, k).cpu().numpy()
        st.update(synthetic_state(5, n, robot.num_bodies, 12, np.float32(p.default_dof_pos), p.feet_idx,
                                  p.term_idx[:p.n_term_bodies], z0=0.3, teleport_band_frac=0.05))
        statekit.apply_to_product(env, st)
        actions = cu(rng.normal(0, 1, (n, 12)).astype(np.float32))
        outs = []
        for s_ in range(3):
            o = env.step(actions)[:4]
            torch.cuda.synchronize()
            outs.append([t.clone() for t in o] + [env.measured_heights.clone(), env.root_states.clone()])
        results.append((outs, statekit.state_from_product(env)))
    (oa, sa), (ob, sb) = results
    names = ("obs", "priv", "rew", "reset", "measured_heights", "root_states")
    for s_, (xa, xb) in enumerate(zip(oa, ob)):
        for i, (a, b) in enumerate(zip(xa, xb)):
            assert torch.equal(a, b), "step %d %s differs in %d elements" % (s_, names[i], int((a != b).sum()))
    for k in sa:
        assert np.array_equal(np.asarray(sa[k]), np.asarray(sb[k])), "state %r differs" % k


def test_uniform_command_curriculum_vs_golden(golden_dir):
    """_update_command_curriculum_uniform (legged_robot.py:851-880) against the reference's recorded ranges: twelve
    scripted calls (on / off the max_episode_length step, above / below the thresholds, up to the clip)."""
    g = np.load(os.path.join(golden_dir, "curriculum_uniform.npz"))
    from rapid_locomotion_rl_b200.envs import LeggedRobot
    cfg, robot, terrain = build_case("mc_flat", 48)
    c = cfg.commands
    c.command_curriculum, c.yaw_command_curriculum = True, True
    c.max_forward_curriculum, c.max_reverse_curriculum, c.max_yaw_curriculum = 1.5, 0.5, 1.3
    env = LeggedRobot(cfg, sim_device=DEV, headless=True, terrain=terrain)
    assert env.max_episode_length == float(g["max_episode_length"])
    r = cfg.command_ranges
    assert [list(r["lin_vel_x"]), list(r["ang_vel_yaw"])] == g["ranges0"].tolist()
    for i in range(int(g["n_calls"])):
        env.episode_sums["tracking_lin_vel"].copy_(cu(g["call%d/lin" % i]))
        env.episode_sums["tracking_ang_vel"].copy_(cu(g["call%d/ang" % i]))
        env.common_step_counter = int(g["call%d/step" % i])
        env._update_command_curriculum_uniform(cu(g["call%d/ids" % i]))
        got = [[float(x) for x in r["lin_vel_x"]], [float(x) for x in r["ang_vel_yaw"]]]
        assert np.allclose(got, g["call%d/ranges" % i], rtol=0, atol=1e-12), (i, got, g["call%d/ranges" % i].tolist())


def test_upstream_order_switch():
    """upstream_order=True restores the call sites this fork commented out: time-outs enter reset_buf
    (:197-198), terminated envs are reset inside step (:177) and get new commands (:246), commands are resampled every
    resampling_time (:581).  The default keeps the fork's order: no resets, no time-outs, no resampling inside step."""
    from rapid_locomotion_rl_b200.envs import LeggedRobot
    from rapid_locomotion_rl_b200.sim import synthetic_state
    n = 256
    envs = {}
    for up in (False, True):
        cfg, robot, terrain = build_case("mc_flat", n)
        env = LeggedRobot(cfg, sim_device=DEV, headless=True, terrain=terrain, seed=4, upstream_order=up)
        p = env.params
        assert p.timeout_resets == int(up)
        st = synthetic_state(2, n, robot.num_bodies, 12, np.float32(p.default_dof_pos), p.feet_idx, p.term_idx[:p.n_term_bodies])
        st["contact_forces"][:, p.term_idx[:p.n_term_bodies]] = 0.0
        st["contact_forces"][:8, p.term_idx[0], 2] = 50.0                  # envs 0-7 terminate by contact
        statekit.apply_to_product(env, st)
        ep = torch.full((n,), 10, dtype=torch.long, device=DEV)
        ep[8:16] = p.max_episode_length                                     # envs 8-15 time out (ep + 1 > max)
        ep[16:24] = p.resample_interval - 1                                 # envs 16-23 hit the resampling interval
        env.episode_length_buf.copy_(ep)
        env.commands[:, :3] = 7.0                                           # a value no draw produces
        env.last_actions[:] = 1.0
        obs, priv, rew, reset, extras = env.step(torch.zeros(n, 12, device=DEV))
        torch.cuda.synchronize()
        envs[up] = (env, reset.clone())
    env, reset = envs[False]
    assert reset[:8].all() and not reset[8:].any()                          # contact terminations only
    assert torch.equal(env.episode_length_buf[:8], torch.full((8,), 11, device=DEV))       # nobody was reset
    assert (env.commands[:, 0] == 7.0).all()
    env, reset = envs[True]
    assert reset[:16].all() and not reset[16:].any()                        # + time-outs
    assert env.time_out_buf[8:16].all() and not env.time_out_buf[:8].any()
    assert (env.episode_length_buf[:16] == 0).all() and (env.episode_length_buf[24:] == 11).all()
    assert float(env.last_actions[:16].abs().max()) == 0.0 and float(env.last_actions[24:].min()) == 0.0   # step wrote zeros anyway
    torch.testing.assert_close(env.dof_pos[:16], env.default_dof_pos.expand(16, 12))
    assert (env.commands[:24, 0] != 7.0).all() and (env.commands[24:, 0] == 7.0).all()     # reset envs + interval envs resampled


# ---- train / eval env split (legged_robot.py:37-46, :204-225, :456-469) ---------------------------------------------------
def make_eval_env(n_train=40, **kw):
    from cases import build_eval_case
    from rapid_locomotion_rl_b200.envs import LeggedRobot
    cfg, ev, robot, terrain = build_eval_case(n_train)
    return LeggedRobot(cfg, sim_device=DEV, headless=True, eval_cfg=ev, terrain=terrain, **kw), cfg, ev


def test_eval_split_construction_vs_golden(golden_dir):
    """Env counts, per-range env origins / terrain-origin tables (the evaluation tiles lie below the training tiles), the
    episode length left by the second _parse_cfg call, the evaluation push interval."""
    g = load(golden_dir, "mc_eval")
    env, cfg, ev = make_eval_env()
    assert (env.num_envs, env.num_train_envs, env.num_eval_envs) == (56, 40, 16)
    np.testing.assert_array_equal(env.terrain_types.cpu().numpy(), g["init/terrain_types"])
    np.testing.assert_allclose(env.terrain_origins_t.cpu().numpy(), g["init/train_terrain_origins"], rtol=0, atol=0)
    np.testing.assert_allclose(env.terrain_origins_eval_t.cpu().numpy(), g["init/eval_terrain_origins"], rtol=0, atol=0)
    # (terrain levels are random draws; with the same levels the origins are table look-ups)
    lv, ty = env.terrain_levels.cpu().numpy(), env.terrain_types.cpu().numpy()
    want = np.concatenate([g["init/train_terrain_origins"][lv[:40], ty[:40]], g["init/eval_terrain_origins"][lv[40:], ty[40:]]])
    np.testing.assert_allclose(env.env_origins.cpu().numpy(), want, rtol=0, atol=0)
    assert float(env.max_episode_length) == float(g["const/max_episode_length_attr"])
    assert float(ev.domain_rand.push_interval) == float(g["const/eval_push_interval"])
    assert env.terrain.eval_x_offset == int(g["const/eval_x_offset"])


def test_eval_split_step_vs_golden(golden_dir):
    """Three steps of 40 training + 16 evaluation envs: teleport bands, pushes and DOF-property re-draws follow each
    range's own Cfg, everything else the training Cfg - against the unmodified reference run with `eval_cfg`."""
    g = load(golden_dir, "mc_eval")
    env, cfg, ev = make_eval_env()
    for s in range(3):
        pre = "step%d/" % s
        statekit.apply_to_product(env, sub(g, pre + "before/"))
        env._inject = dict(noise_u=cu(g[pre + "noise_u"]), dr_u=cu(g[pre + "dr_u"]), push_u=cu(g[pre + "push_u"]))
        obs, priv, rew, reset, _ = env.step(cu(g[pre + "actions"]))
        check_step(env, obs, priv, rew, reset, g[pre + "obs"], g[pre + "priv"], g[pre + "rew"], g[pre + "reset"],
                   "mc_eval step %d" % s)
        got = statekit.state_from_product(env)
        statekit.assert_state_close(got, sub(g, pre + "after/"), RTOL, ATOL, skip=("contact_forces",),
                                    label="mc_eval step %d" % s)
    # the golden inputs do exercise the split: evaluation envs were pushed, teleported with their own band, re-drawn
    a, b = g["step1/before/root_states"], g["step1/after/root_states"]
    assert (a[40:, 7:9] != b[40:, 7:9]).any() and (a[:40, 7:9] == b[:40, 7:9]).all()
    a, b = g["step0/before/root_states"], g["step0/after/root_states"]
    assert (a[40:44, :2] != b[40:44, :2]).any() and (a[44, :2] == b[44, :2]).all()
    assert (g["step0/before/Kp_factors"][40:] != g["step0/after/Kp_factors"][40:]).any()


def test_eval_split_reset_vs_golden(golden_dir):
    """reset_idx over ids of both ranges (each range's DOF-property ranges and initial-position ranges; the evaluation
    rollout results saved once into episode_sums_eval), then reset_evaluation_envs."""
    g = load(golden_dir, "mc_eval")
    env, cfg, ev = make_eval_env()
    statekit.apply_to_product(env, sub(g, "reset/before/"))
    env._inject = dict(reset_dr_u=cu(g["reset/dr_u"]), init_u=cu(g["reset/init_u"]))
    env._reset_u8.copy_(cu(g["step2/reset"].astype(np.uint8)))
    env.extras = {}
    env.reset_idx(cu(g["reset/ids"], torch.long))
    torch.cuda.synchronize()
    got = statekit.state_from_product(env)
    statekit.assert_state_close(got, sub(g, "reset/after/"), RTOL, ATOL, skip=("contact_forces", "torques", "joint_pos_target",
                                "base_lin_vel", "base_ang_vel", "projected_gravity", "last_root_vel"), label="mc_eval reset")
    assert np.array_equal(env.reset_buf.cpu().numpy(), g["reset/reset_buf"])
    for k, v in sub(g, "reset/extras/").items():
        np.testing.assert_allclose(float(env.extras["train/episode"][k]), float(v), rtol=1e-5, atol=2e-6, err_msg=k)
    assert sorted(env.extras["eval/episode"].keys()) == list(g["reset/eval_extras_keys"])
    for k, v in sub(g, "reset/episode_sums_eval/").items():
        np.testing.assert_allclose(env.episode_sums_eval[k].cpu().numpy(), v, rtol=RTOL, atol=ATOL, err_msg=k)

    # ---- reset_evaluation_envs ----
    st = sub(g, "evalreset/before/")
    statekit.apply_to_product(env, st)
    for k, v in sub(g, "evalreset/before/episode_sums_eval/").items():
        env.episode_sums_eval[k].copy_(cu(v))
    env._inject = dict(reset_dr_u=cu(g["evalreset/dr_u"]), init_u=cu(g["evalreset/init_u"]))
    env.reset_evaluation_envs()
    torch.cuda.synchronize()
    got = statekit.state_from_product(env)
    statekit.assert_state_close(got, sub(g, "evalreset/after/"), RTOL, ATOL, skip=("contact_forces", "torques", "joint_pos_target",
                                "base_lin_vel", "base_ang_vel", "projected_gravity", "last_root_vel"), label="mc_eval eval reset")
    for k, v in sub(g, "evalreset/after/episode_sums_eval/").items():
        np.testing.assert_allclose(env.episode_sums_eval[k].cpu().numpy(), v, rtol=0, atol=0, err_msg=k)
    assert sorted(env.extras.get("eval/episode", {}).keys()) == list(g["evalreset/eval_extras_keys"])


def test_eval_split_runner_iteration():
    """One Runner iteration with evaluation envs: the storage holds the training envs only, evaluation envs are stepped with
    the student policy and reset every eval_freq iterations."""
    from cases import build_eval_case
    from rapid_locomotion_rl_b200.envs import VelocityTrackingEasyEnv, HistoryWrapper
    from rapid_locomotion_rl_b200.ppo import Runner
    cfg, ev, robot, terrain = build_eval_case(64)
    ev.env.num_envs = 32
    from rapid_locomotion_rl_b200 import config as C
    terrain = C.TerrainInfo(cfg.terrain, eval_terrain=ev.terrain)
    env = HistoryWrapper(VelocityTrackingEasyEnv(sim_device=DEV, headless=True, cfg=cfg, eval_cfg=ev, terrain=terrain))
    runner = Runner(env, device=DEV)
    hist = runner.learn(2, eval_freq=1)
    assert len(hist) == 2 and all(np.isfinite(h["mean_value_loss"]) for h in hist)
    assert runner.alg.storage.observations.shape[1] == 64
    assert int(env.episode_length_buf[64:].max()) <= runner.num_steps_per_env      # evaluation envs were reset
