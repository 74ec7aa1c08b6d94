"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference) on CPU fp32 through the shims in tests/_ref_shims.

    python tests/golden/make_golden.py            # all cases (one subprocess per config: the
                                                  # reference's Cfg is a process-global class)
    python tests/golden/make_golden.py mc_flat    # one case

Randomness is injected: torch.rand / rand_like / randint_like are replaced, for the duration of a
reference call, by a scripted queue fed from uniforms that are also written to the fixture, so the
oracle and the CUDA kernels can be driven with exactly the numbers the reference consumed.
tests/test_oracle_vs_golden.py then requires oracle/env_oracle.py to reproduce these tensors bit
for bit on the CPU - this is what pins the oracle.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "_ref_shims"))

CASES = ["mc_flat", "go1", "go1_alt", "mc_rough", "mc_rough_full", "mc_only_lin", "mc_only_ang", "learner", "rollout", "checkpoint", "curriculum_uniform", "eval_split", "hlp"]
N_ENVS = 48
N_STEPS = 3


from cases import case_cfg_hook  # noqa: E402


def synth_inputs(rng, env_like, n, nb, default_dof_pos, feet, term, span_x, span_y, z0):
    """Per-step simulator state (SURVEY.md 8(d) distribution, scaled to the terrain extent)."""
    root = np.zeros((n, 13), np.float32)
    root[:, 0] = rng.uniform(0.5, span_x - 0.5, n)
    root[:, 1] = rng.uniform(0.5, span_y - 0.5, n)
    root[:, 2] = z0 + rng.normal(0, 0.02, n)
    q = np.concatenate([rng.normal(0, 0.15, (n, 3)), np.ones((n, 1))], 1)
    root[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    root[:, 7:13] = rng.normal(0, 0.5, (n, 6))
    dof = np.zeros((n, 12, 2), np.float32)
    dof[:, :, 0] = default_dof_pos[None] + rng.normal(0, 0.6, (n, 12))
    dof[:, :, 1] = rng.normal(0, 6.0, (n, 12))
    con = rng.normal(0, 0.5, (n, nb, 3)).astype(np.float32)
    big = rng.random((n, len(term))) < 0.1
    for k, b in enumerate(term):
        con[big[:, k], b] *= 20
    for b in feet:
        con[:, b, 2] = np.abs(rng.normal(0, 30, n)) * (rng.random(n) < 0.5)
    actions = (rng.normal(0, 1, (n, 12)) * np.where(rng.random((n, 1)) < 0.1, 150.0, 1.0)).astype(np.float32)
    return root, dof, con.astype(np.float32), actions


class ScriptedRand:
    """Replaces torch.rand / rand_like / randint_like with a scripted FIFO for one reference call."""

    def __init__(self, torch, queue):
        self.torch, self.queue = torch, list(queue)

    def __enter__(self):
        t = self.torch
        self.saved = (t.rand, t.rand_like, t.randint_like)

        def pop(shape, what):
            assert self.queue, "reference asked for %s%s but the script is empty" % (what, tuple(shape))
            v = self.queue.pop(0)
            assert tuple(v.shape) == tuple(shape), "%s: scripted %s, requested %s" % (what, tuple(v.shape), tuple(shape))
            return v.clone()

        def rand(*shape, **kw):
            if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
                shape = tuple(shape[0])
            return pop(shape, "rand")

        def rand_like(x, **kw):
            return pop(x.shape, "rand_like")

        def randint_like(x, *a, **kw):
            return pop(x.shape, "randint_like").to(x.dtype)
        t.rand, t.rand_like, t.randint_like = rand, rand_like, randint_like
        return self

    def __exit__(self, *exc):
        t = self.torch
        t.rand, t.rand_like, t.randint_like = self.saved
        if exc[0] is None:
            assert not self.queue, "%d scripted draws were not consumed" % len(self.queue)
        return False


def gen_env_case(case):
    import torch
    import harness
    import statekit
    from rapid_locomotion_rl_b200 import sim as psim
    torch.manual_seed(0)
    torch.set_num_threads(1)
    robot = "go1" if case.startswith("go1") else "mini_cheetah"
    rough = case.startswith("mc_rough")
    hf = (lambda r, c: psim.synthetic_heightfield(r, c, seed=3)) if rough else None
    env, Cfg = harness.make_reference_env(robot, N_ENVS, rough=rough, height_fn=hf, cfg_hook=case_cfg_hook(case))
    e = env.env
    N, NB = e.num_envs, e.num_bodies
    rng = np.random.default_rng(abs(hash(case)) % (2 ** 31) if False else {"mc_flat": 11, "go1": 12, "go1_alt": 13, "mc_rough": 14, "mc_rough_full": 15, "mc_only_lin": 16, "mc_only_ang": 17}[case])
    t = Cfg.terrain
    span_x, span_y = t.terrain_length * t.num_rows, t.terrain_width * t.num_cols
    out = {}
    if rough and case != "mc_rough_full":          # (the full table is 9.4 MB: regenerated from its seed by the tests)
        out["heightsamples"] = e.terrain.heightsamples
    if rough:
        out["heightsamples_shape"] = np.array(e.terrain.heightsamples.shape)
        out["heightsamples_sum"] = np.int64(e.terrain.heightsamples.astype(np.int64).sum())

    # ---- randomise the persistent state the step reads -------------------------------------------
    def T(a, dtype=torch.float):
        return torch.from_numpy(np.asarray(a)).to(dtype)
    e.commands[:, :3] = T(rng.uniform(-1, 1, (N, 3)).astype(np.float32))
    e.commands[: N // 8, :2] *= 0.05                      # some near-zero commands (air-time / stand-still gates)
    e.last_actions[:] = T(rng.normal(0, 1, (N, 12)).astype(np.float32))
    e.last_dof_vel[:] = T(rng.normal(0, 3, (N, 12)).astype(np.float32))
    e.motor_strengths[:] = T(rng.uniform(0.9, 1.1, (N, 1)).astype(np.float32)).repeat(1, 12)
    e.Kp_factors[:] = T(rng.uniform(0.8, 1.3, (N, 1)).astype(np.float32)).repeat(1, 12)
    e.Kd_factors[:] = T(rng.uniform(0.5, 1.5, (N, 1)).astype(np.float32)).repeat(1, 12)
    e.friction_coeffs[:] = T(rng.uniform(0.05, 4.5, N).astype(np.float32))
    e.restitutions[:] = T(rng.uniform(0, 1, N).astype(np.float32))
    e.payloads[:] = T(rng.uniform(-1, 3, N).astype(np.float32))
    e.com_displacements[:] = T(rng.uniform(-0.1, 0.1, (N, 3)).astype(np.float32))
    e.feet_air_time[:] = T((rng.uniform(0, 0.6, (N, 4)) * (rng.random((N, 4)) < 0.7)).astype(np.float32))
    e.last_contacts = T(rng.random((N, 4)) < 0.3, torch.bool)
    ri = int(Cfg.domain_rand.rand_interval)
    ep = rng.integers(0, 1001, N)
    ep[::7] = ri * rng.integers(1, 3, len(ep[::7])) - 1     # these hit the DOF-property re-draw on step 1
    if Cfg.domain_rand.push_robots:
        pi = int(Cfg.domain_rand.push_interval)
        ep[1::9] = pi * rng.integers(1, 50, len(ep[1::9])) - 2   # pushed on step 2
    e.episode_length_buf[:] = T(ep, torch.long)
    for k in e.episode_sums:
        e.episode_sums[k][:] = T(rng.normal(0, 1, N).astype(np.float32))
    for k in e.command_sums:
        e.command_sums[k][:] = T(rng.normal(0, 1, N).astype(np.float32))

    feet = e.feet_indices.tolist(); term = e.termination_contact_indices.tolist()
    z0 = 0.30 if robot == "mini_cheetah" else 0.34
    dflt = e.default_dof_pos[0].numpy()
    meta = dict(case=case, robot=robot, rough=rough, n_envs=N, n_steps=N_STEPS)
    for s in range(N_STEPS):
        root, dof, con, actions = synth_inputs(rng, e, N, NB, dflt, feet, term, span_x, span_y, z0)
        if Cfg.terrain.teleport_robots:   # a few robots inside the teleport band on each side
            root[0, 0] = 0.7; root[1, 0] = span_x - 1.2; root[2, 1] = 1.1; root[3, 1] = span_y - 0.3
        e.all_root_states[:] = T(root); e.all_dof_state[:] = T(dof.reshape(-1, 2)); e.all_contact_forces[:] = T(con.reshape(-1, 3))
        noise_u = rng.random((N, e.num_obs)).astype(np.float32)
        dr_u = rng.random((3, N)).astype(np.float32)
        push_u = rng.random((2, N)).astype(np.float32)
        # canonical "before" state: what this step will read
        before = statekit.state_from_reference(e)
        before["root_states"], before["dof_state"], before["contact_forces"] = root, dof, con
        # script the draws in the order the reference makes them (:588 push, :593 re-draw, :392 noise)
        epn = e.episode_length_buf.numpy() + 1
        queue = []
        if Cfg.domain_rand.push_robots:
            ids = np.nonzero(epn % int(Cfg.domain_rand.push_interval) == 0)[0]
            queue.append(T(push_u[:, ids].T.copy()))
        ids = np.nonzero(epn % ri == 0)[0]
        if len(ids) > 0:
            for k, flag in enumerate((Cfg.domain_rand.randomize_motor_strength, Cfg.domain_rand.randomize_Kp_factor,
                                      Cfg.domain_rand.randomize_Kd_factor)):
                if flag:
                    queue.append(T(dr_u[k, ids].copy()))
        if Cfg.noise.add_noise:
            queue.append(T(noise_u))
        with ScriptedRand(torch, queue):
            obs, priv, rew, reset, _ = type(e).__mro__[1].step(e, T(actions))   # LeggedRobot.step, skipping the numpy extras
        after = statekit.state_from_reference(e)
        pre = "step%d/" % s
        for k, v in before.items():
            out[pre + "before/" + k] = v
        for k, v in after.items():
            out[pre + "after/" + k] = v
        out[pre + "actions"] = actions; out[pre + "noise_u"] = noise_u; out[pre + "dr_u"] = dr_u; out[pre + "push_u"] = push_u
        out[pre + "obs"] = obs.numpy().copy(); out[pre + "priv"] = priv.numpy().copy()
        out[pre + "rew"] = rew.numpy().copy(); out[pre + "reset"] = reset.numpy().copy()
        if rough:
            out[pre + "measured_heights"] = e.measured_heights.numpy().copy()

    # ---- reset_idx on a subset (:227-290) ------------------------------------------------------------
    ids = np.sort(rng.choice(N, N // 3, replace=False))
    before = statekit.state_from_reference(e)
    if not e.custom_origins:
        before["root_states"] = e.all_root_states.numpy().copy()
    reset_dr_u = rng.random((3, N)).astype(np.float32)
    init_u = rng.random((2, N)).astype(np.float32)
    level_u = rng.random(N).astype(np.float32)
    if Cfg.terrain.curriculum:
        # make some robots walk far / stay put so terrain levels move both ways; some at the last level
        e.root_states[ids[::2], :2] = e.env_origins[ids[::2], :2] + 5.0
        e.terrain_levels[ids[1::4]] = Cfg.terrain.max_terrain_level - 1
        e.root_states[ids[1::4], :2] = e.env_origins[ids[1::4], :2] + 6.0
        before = statekit.state_from_reference(e)
    queue = []
    if Cfg.terrain.curriculum:
        queue.append(T(np.minimum((level_u[ids] * Cfg.terrain.max_terrain_level).astype(np.int64),
                                  Cfg.terrain.max_terrain_level - 1), torch.long))
    for k, flag in enumerate((Cfg.domain_rand.randomize_motor_strength, Cfg.domain_rand.randomize_Kp_factor,
                              Cfg.domain_rand.randomize_Kd_factor)):
        if flag:
            queue.append(T(reset_dr_u[k, ids].copy()))
    if e.custom_origins:
        queue.append(T(init_u[:, ids].T.copy()))
    e.extras = {}
    with ScriptedRand(torch, queue):
        e.reset_idx(T(ids, torch.long))
    after = statekit.state_from_reference(e)
    if not e.custom_origins:
        after["root_states"] = e.all_root_states.numpy().copy()   # plane: the reset writes the sim tensor (:733)
    after["dof_state"] = e.all_dof_state.numpy().reshape(N, 12, 2).copy()
    for k, v in before.items():
        out["reset/before/" + k] = v
    for k, v in after.items():
        out["reset/after/" + k] = v
    out["reset/ids"] = ids; out["reset/dr_u"] = reset_dr_u; out["reset/init_u"] = init_u; out["reset/level_u"] = level_u
    out["reset/reset_buf"] = e.reset_buf.numpy().copy()
    for k, v in e.extras["train/episode"].items():
        if k.startswith("rew_"):
            out["reset/extras/" + k] = np.float32(v)

    # ---- _resample_commands (:595-626) with the reference's own MT19937 stream ------------------------
    all_ids = T(np.arange(N), torch.long)
    e._resample_commands(all_ids)                    # initial draw so every env owns a bin
    lin_scale, ang_scale = e.reward_scales["tracking_lin_vel"], e.reward_scales["tracking_ang_vel"]
    e.command_sums["tracking_lin_vel"][:] = T((rng.uniform(0.5, 1.1, N) * 500 * lin_scale).astype(np.float32))
    e.command_sums["tracking_ang_vel"][:] = T((rng.uniform(0.2, 0.8, N) * 500 * ang_scale).astype(np.float32))
    for rnd in range(2):
        ids = np.sort(rng.choice(N, N // 2, replace=False))
        pre = "resample%d/" % rnd
        out[pre + "ids"] = ids
        out[pre + "before/weights"] = e.curriculum.weights.copy()
        out[pre + "before/bins"] = e.env_command_bins.copy().astype(np.int64)
        out[pre + "before/commands"] = e.commands.numpy().copy()
        for k, v in e.command_sums.items():
            out[pre + "before/command_sums/" + k] = v.numpy().copy()
        st = e.curriculum.rng.get_state()
        e._resample_commands(T(ids, torch.long))
        r2 = np.random.RandomState(); r2.set_state(st)
        u_bin = np.zeros(N); u_cell = np.zeros((N, 3))
        u_bin[ids] = r2.random_sample(len(ids))
        u_cell[ids] = np.stack([r2.random_sample(3) for _ in ids])
        out[pre + "u_bin"] = u_bin; out[pre + "u_cell"] = u_cell
        out[pre + "after/weights"] = e.curriculum.weights.copy()
        out[pre + "after/bins"] = e.env_command_bins.copy().astype(np.int64)
        out[pre + "after/commands"] = e.commands.numpy().copy()
        for k, v in e.command_sums.items():
            out[pre + "after/command_sums/" + k] = v.numpy().copy()
        e.command_sums["tracking_lin_vel"][:] = T((rng.uniform(0.5, 1.1, N) * 500 * lin_scale).astype(np.float32))
        e.command_sums["tracking_ang_vel"][:] = T((rng.uniform(0.2, 0.8, N) * 500 * ang_scale).astype(np.float32))

    # ---- constants of the frozen config: known answers for config freezing ---------------------------
    out["const/dt"] = np.float64(e.dt); out["const/max_episode_length"] = np.float64(e.max_episode_length)
    out["const/rand_interval"] = np.float64(Cfg.domain_rand.rand_interval)
    out["const/push_interval"] = np.float64(Cfg.domain_rand.push_interval)
    out["const/p_gains"] = e.p_gains.numpy(); out["const/d_gains"] = e.d_gains.numpy()
    out["const/default_dof_pos"] = e.default_dof_pos.numpy(); out["const/torque_limits"] = e.torque_limits.numpy()
    out["const/dof_pos_limits"] = e.dof_pos_limits.numpy(); out["const/dof_vel_limits"] = e.dof_vel_limits.numpy()
    out["const/noise_scale_vec"] = e.noise_scale_vec.numpy()
    out["const/feet_indices"] = e.feet_indices.numpy(); out["const/termination_contact_indices"] = e.termination_contact_indices.numpy()
    out["const/penalised_contact_indices"] = e.penalised_contact_indices.numpy()
    out["const/reward_names"] = np.array(e.reward_names)
    out["const/reward_scales"] = np.array([e.reward_scales[k] for k in e.reward_scales])
    out["const/reward_scale_keys"] = np.array(list(e.reward_scales.keys()))
    out["const/initial_weight_sum"] = np.float64(30.0)
    out["meta"] = np.array(repr(meta))
    path = os.path.join(HERE, "env_%s.npz" % case)
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def gen_eval_split_case():
    """Train / eval env split (legged_robot.py:37-46, :204-225, :456-469; base_task.py:43-49): 40 training envs with the
    shipped Mini Cheetah Cfg + 16 evaluation envs with cases.eval_cfg_hook's Cfg, through three steps (teleport, pushes and
    DOF-property re-draws per range), a reset_idx over ids of both ranges and reset_evaluation_envs."""
    import torch
    import harness
    import statekit
    from cases import eval_cfg_hook
    torch.manual_seed(0)
    torch.set_num_threads(1)
    N_TRAIN = 40
    env, Cfg = harness.make_reference_env("mini_cheetah", N_TRAIN, eval_hook=eval_cfg_hook)
    e = env.env
    ev = env.eval_cfg_used
    N, NB, NT = e.num_envs, e.num_bodies, e.num_train_envs
    assert NT == N_TRAIN and N == N_TRAIN + 16
    rng = np.random.default_rng(31)
    out = {}

    def T(a, dtype=torch.float):
        return torch.from_numpy(np.asarray(a)).to(dtype)
    out["init/env_origins"] = e.env_origins.numpy().copy()
    out["init/terrain_types"] = e.terrain_types.numpy().copy()
    out["init/eval_terrain_origins"] = ev.terrain.terrain_origins.numpy().copy()
    out["init/train_terrain_origins"] = Cfg.terrain.terrain_origins.numpy().copy()
    e.commands[:, :3] = T(rng.uniform(-1, 1, (N, 3)).astype(np.float32))
    e.last_actions[:] = T(rng.normal(0, 1, (N, 12)).astype(np.float32))
    e.last_dof_vel[:] = T(rng.normal(0, 3, (N, 12)).astype(np.float32))
    e.motor_strengths[:] = T(rng.uniform(0.9, 1.1, (N, 1)).astype(np.float32)).repeat(1, 12)
    e.Kp_factors[:] = T(rng.uniform(0.8, 1.3, (N, 1)).astype(np.float32)).repeat(1, 12)
    e.Kd_factors[:] = T(rng.uniform(0.5, 1.5, (N, 1)).astype(np.float32)).repeat(1, 12)
    e.friction_coeffs[:] = T(rng.uniform(0.05, 4.5, N).astype(np.float32))
    e.restitutions[:] = T(rng.uniform(0, 1, N).astype(np.float32))
    e.payloads[:] = T(rng.uniform(-1, 3, N).astype(np.float32))
    e.com_displacements[:] = T(rng.uniform(-0.1, 0.1, (N, 3)).astype(np.float32))
    e.feet_air_time[:] = T((rng.uniform(0, 0.6, (N, 4)) * (rng.random((N, 4)) < 0.7)).astype(np.float32))
    e.last_contacts = T(rng.random((N, 4)) < 0.3, torch.bool)
    ri = int(Cfg.domain_rand.rand_interval)
    pi_eval = int(ev.domain_rand.push_interval)
    ep = rng.integers(0, 1001, N)
    ep[::5] = ri * rng.integers(1, 3, len(ep[::5])) - 1          # DOF-property re-draw on step 1 (both ranges)
    ep[NT + 1::3] = pi_eval * rng.integers(1, 50, len(ep[NT + 1::3])) - 2   # evaluation envs pushed on step 2
    e.episode_length_buf[:] = T(ep, torch.long)
    for k in e.episode_sums:
        e.episode_sums[k][:] = T(rng.normal(0, 1, N).astype(np.float32))
    for k in e.command_sums:
        e.command_sums[k][:] = T(rng.normal(0, 1, N).astype(np.float32))
    feet = e.feet_indices.tolist(); term = e.termination_contact_indices.tolist()
    dflt = e.default_dof_pos[0].numpy()
    tt, te = Cfg.terrain, ev.terrain
    xo_eval = int(te.x_offset * te.horizontal_scale)
    for s in range(N_STEPS):
        root, dof, con, actions = synth_inputs(rng, e, N, NB, dflt, feet, term, tt.terrain_length * tt.num_rows,
                                               tt.terrain_width * tt.num_cols, 0.30)
        # evaluation robots live on the evaluation tiles (x shifted by the training rows)
        root[NT:, 0] = xo_eval + rng.uniform(0.5, te.terrain_length * te.num_rows - 0.5, N - NT)
        root[NT:, 1] = rng.uniform(0.5, te.terrain_width * te.num_cols - 0.5, N - NT)
        # robots inside the teleport bands of either range
        root[0, 0] = 0.7; root[1, 0] = tt.terrain_length * tt.num_rows - 1.2; root[2, 1] = 1.1
        root[NT, 0] = xo_eval + 0.9; root[NT + 1, 0] = xo_eval + te.terrain_length * te.num_rows - 0.4
        root[NT + 2, 1] = 1.2; root[NT + 3, 1] = te.terrain_width * te.num_cols - 0.8
        root[NT + 4, 0] = xo_eval + 1.8          # inside the TRAINING threshold (2.0) but not the evaluation one (1.5)
        e.all_root_states[:] = T(root); e.all_dof_state[:] = T(dof.reshape(-1, 2)); e.all_contact_forces[:] = T(con.reshape(-1, 3))
        noise_u = rng.random((N, e.num_obs)).astype(np.float32)
        dr_u = rng.random((3, N)).astype(np.float32)
        push_u = rng.random((2, N)).astype(np.float32)
        before = statekit.state_from_reference(e)
        before["root_states"], before["dof_state"], before["contact_forces"] = root, dof, con
        epn = e.episode_length_buf.numpy() + 1
        queue = []
        # :588 pushes per range (training Cfg: off), :593 re-draws per range, :392 noise
        for lo, hi, c in ((0, NT, Cfg), (NT, N, ev)):
            if c.domain_rand.push_robots:
                ids = lo + np.nonzero(epn[lo:hi] % int(c.domain_rand.push_interval) == 0)[0]
                queue.append(T(push_u[:, ids].T.copy()))
        ids_all = np.nonzero(epn % ri == 0)[0]
        for lo, hi, c in ((0, NT, Cfg), (NT, N, ev)):
            ids = ids_all[(ids_all >= lo) & (ids_all < hi)]
            if len(ids) == 0:
                continue
            for k, flag in enumerate((c.domain_rand.randomize_motor_strength, c.domain_rand.randomize_Kp_factor,
                                      c.domain_rand.randomize_Kd_factor)):
                if flag:
                    queue.append(T(dr_u[k, ids].copy()))
        queue.append(T(noise_u))
        with ScriptedRand(torch, queue):
            obs, priv, rew, reset, _ = type(e).__mro__[1].step(e, T(actions))
        after = statekit.state_from_reference(e)
        pre = "step%d/" % s
        for k, v in before.items():
            out[pre + "before/" + k] = v
        for k, v in after.items():
            out[pre + "after/" + k] = v
        out[pre + "actions"] = actions; out[pre + "noise_u"] = noise_u; out[pre + "dr_u"] = dr_u; out[pre + "push_u"] = push_u
        out[pre + "obs"] = obs.numpy().copy(); out[pre + "priv"] = priv.numpy().copy()
        out[pre + "rew"] = rew.numpy().copy(); out[pre + "reset"] = reset.numpy().copy()

    # ---- reset_idx over ids of both ranges (:227-290 through _call_train_eval) ----
    ids = np.sort(np.concatenate([rng.choice(NT, NT // 3, replace=False), NT + rng.choice(N - NT, (N - NT) // 2, replace=False)]))
    before = statekit.state_from_reference(e)
    reset_dr_u = rng.random((3, N)).astype(np.float32)
    init_u = rng.random((2, N)).astype(np.float32)
    queue = []
    for lo, hi, c in ((0, NT, Cfg), (NT, N, ev)):                     # :247 per range
        sel = ids[(ids >= lo) & (ids < hi)]
        for k, flag in enumerate((c.domain_rand.randomize_motor_strength, c.domain_rand.randomize_Kp_factor,
                                  c.domain_rand.randomize_Kd_factor)):
            if flag:
                queue.append(T(reset_dr_u[k, sel].copy()))
    for lo, hi in ((0, NT), (NT, N)):                                  # :251 per range (custom origins: xy offset draw)
        sel = ids[(ids >= lo) & (ids < hi)]
        queue.append(T(init_u[:, sel].T.copy()))
    e.extras = {}
    with ScriptedRand(torch, queue):
        e.reset_idx(T(ids, torch.long))
    after = statekit.state_from_reference(e)
    after["dof_state"] = e.all_dof_state.numpy().reshape(N, 12, 2).copy()
    for k, v in before.items():
        out["reset/before/" + k] = v
    for k, v in after.items():
        out["reset/after/" + k] = v
    out["reset/ids"] = ids; out["reset/dr_u"] = reset_dr_u; out["reset/init_u"] = init_u
    out["reset/reset_buf"] = e.reset_buf.numpy().copy()
    for k, v in e.extras["train/episode"].items():
        if k.startswith("rew_"):
            out["reset/extras/" + k] = np.float32(v)
    out["reset/eval_extras_keys"] = np.array(sorted(e.extras["eval/episode"].keys()))
    for k, v in e.episode_sums_eval.items():
        out["reset/episode_sums_eval/" + k] = v.numpy().copy()

    # ---- reset_evaluation_envs (:204-225) ----
    for k in e.episode_sums:
        e.episode_sums[k][:] = T(rng.normal(0, 1, N).astype(np.float32))
    before = statekit.state_from_reference(e)
    saved_before = {k: v.numpy().copy() for k, v in e.episode_sums_eval.items()}
    n_eval = N - NT
    ev_dr_u = rng.random((3, N)).astype(np.float32)
    ev_init_u = rng.random((2, N)).astype(np.float32)
    sel = np.arange(NT, N)
    queue = []
    for k, flag in enumerate((ev.domain_rand.randomize_motor_strength, ev.domain_rand.randomize_Kp_factor,
                              ev.domain_rand.randomize_Kd_factor)):
        if flag:
            queue.append(T(ev_dr_u[k, sel].copy()))
    queue.append(T(ev_init_u[:, sel].T.copy()))
    # (extras still holds "eval/episode" from the reset above: reset_evaluation_envs :215 writes into it and raises
    # KeyError when no evaluation env was ever reset through reset_idx)
    with ScriptedRand(torch, queue):
        e.reset_evaluation_envs()
    after = statekit.state_from_reference(e)
    after["dof_state"] = e.all_dof_state.numpy().reshape(N, 12, 2).copy()
    for k, v in before.items():
        out["evalreset/before/" + k] = v
    for k, v in saved_before.items():
        out["evalreset/before/episode_sums_eval/" + k] = v
    for k, v in after.items():
        out["evalreset/after/" + k] = v
    out["evalreset/dr_u"] = ev_dr_u; out["evalreset/init_u"] = ev_init_u
    for k, v in e.episode_sums_eval.items():
        out["evalreset/after/episode_sums_eval/" + k] = v.numpy().copy()
    out["evalreset/eval_extras_keys"] = np.array(sorted(e.extras.get("eval/episode", {}).keys()))
    out["const/max_episode_length_attr"] = np.float64(e.max_episode_length)
    out["const/eval_push_interval"] = np.float64(ev.domain_rand.push_interval)
    out["const/eval_x_offset"] = np.int64(te.x_offset)
    path = os.path.join(HERE, "env_mc_eval.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def gen_learner_case():
    """GAE (rollout_storage.py:76-90) and the PPO update (ppo.py:94-178) of the reference."""
    import torch
    import harness
    harness.install()
    import isaacgym  # noqa: F401  (fake)
    from mini_gym_learn.ppo.rollout_storage import RolloutStorage
    torch.manual_seed(0)
    torch.set_num_threads(1)
    rng = np.random.default_rng(21)
    out = {}
    T_, N = 24, 50
    st = RolloutStorage(N, T_, [42], [18], [630], [12], device="cpu")
    rewards = rng.normal(0, 0.05, (T_, N, 1)).astype(np.float32)
    values = rng.normal(0, 1.0, (T_, N, 1)).astype(np.float32)
    dones = (rng.random((T_, N, 1)) < 0.05).astype(np.uint8)
    last_values = rng.normal(0, 1.0, (N, 1)).astype(np.float32)
    st.rewards[:] = torch.from_numpy(rewards); st.values[:] = torch.from_numpy(values)
    st.dones[:] = torch.from_numpy(dones)
    st.compute_returns(torch.from_numpy(last_values), 0.99, 0.95)
    out["gae/rewards"] = rewards; out["gae/values"] = values; out["gae/dones"] = dones
    out["gae/last_values"] = last_values
    out["gae/returns"] = st.returns.numpy().copy(); out["gae/advantages"] = st.advantages.numpy().copy()
    # ---- PPO.update (ppo.py:94-178) on a small synthetic rollout --------------------------------------
    # inputs (weights, env outputs, permutation) come from tests/cases.py seeds; only the tensors the
    # reference PRODUCES are stored
    from cases import learner_weights, learner_rollout_inputs, tensor_digest
    from mini_gym_learn.ppo import ActorCritic
    from mini_gym_learn.ppo.ppo import PPO
    torch.manual_seed(1)
    N2, T2 = 64, 8
    ac = ActorCritic(42, 18, 630, 12)
    ac.load_state_dict({k: torch.from_numpy(v) for k, v in learner_weights().items()})
    ppo = PPO(ac, device="cpu")
    ppo.init_storage(N2, T2, [42], [18], [630], [12])
    steps, last, perm = learner_rollout_inputs(N2, T2)
    T = torch.from_numpy
    for sd in steps:
        ppo.act(T(sd["obs"]), T(sd["priv"]), T(sd["hist"]))
        ppo.process_env_step(T(sd["rew"]), T(sd["done"]), {"env_bins": torch.zeros(N2)})
    ppo.compute_returns(T(last["obs"]), T(last["priv"]))
    st = ppo.storage
    flat = lambda x: x.flatten(0, 1).numpy().copy()
    for name, tns in (("actions", st.actions), ("values", st.values), ("returns", st.returns), ("old_logp", st.actions_log_prob),
                      ("advantages", st.advantages), ("old_mu", st.mu), ("old_sigma", st.sigma)):
        out["ppo/storage/" + name] = flat(tns)
    real_randperm = torch.randperm
    torch.randperm = lambda n, **kw: T(perm).clone()
    try:
        res = ppo.update()
    finally:
        torch.randperm = real_randperm
    out["ppo/result"] = np.array(res, dtype=np.float64)
    out["ppo/final_lr"] = np.float64(ppo.learning_rate)
    for k, v in ac.state_dict().items():
        if not k.startswith("encoder."):
            out["ppo/final_digest/" + k] = tensor_digest(v.detach().numpy())
    path = os.path.join(HERE, "learner.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def gen_hlp_case():
    """The reference's second learner, high_level_policy/ppo (tanh networks, USE_LATENT = False): forward passes and one
    full PPO.update of the UNMODIFIED classes on the seeded rollout of the learner case."""
    import torch
    import harness
    harness.install()
    import isaacgym  # noqa: F401  (fake)
    import high_level_policy
    assert high_level_policy.USE_LATENT is False
    from high_level_policy.ppo import ActorCritic
    from high_level_policy.ppo.ppo import PPO
    from cases import hlp_weights, learner_rollout_inputs, tensor_digest
    torch.manual_seed(1)
    torch.set_num_threads(1)
    out = {}
    N2, T2 = 64, 8
    ac = ActorCritic(42, 18, 630, 12)
    out["keys"] = np.array(sorted(ac.state_dict().keys()))
    ac.load_state_dict({k: torch.from_numpy(v) for k, v in hlp_weights().items()})
    T = torch.from_numpy
    steps, last, perm = learner_rollout_inputs(N2, T2)
    with torch.no_grad():
        out["fwd/mean"] = ac.act_teacher(T(steps[0]["obs"]), T(steps[0]["priv"])).numpy().copy()
        out["fwd/student"] = ac.act_student(T(steps[0]["obs"]), T(steps[0]["hist"])).numpy().copy()
        out["fwd/value"] = ac.evaluate(T(steps[0]["obs"]), T(steps[0]["priv"])).numpy().copy()
    ppo = PPO(ac, device="cpu")
    ppo.init_storage(N2, T2, [42], [18], [630], [12])
    for sd in steps:
        ppo.act(T(sd["obs"]), T(sd["priv"]), T(sd["hist"]))
        ppo.process_env_step(T(sd["rew"]), T(sd["done"]), {})
    ppo.compute_returns(T(last["obs"]), T(last["priv"]))
    st = ppo.storage
    flat = lambda x: x.flatten(0, 1).numpy().copy()
    for name, tns in (("actions", st.actions), ("values", st.values), ("returns", st.returns), ("old_logp", st.actions_log_prob),
                      ("advantages", st.advantages), ("old_mu", st.mu), ("old_sigma", st.sigma)):
        out["ppo/storage/" + name] = flat(tns)
    real_randperm = torch.randperm
    torch.randperm = lambda n, **kw: T(perm).clone()
    try:
        res = ppo.update()
    finally:
        torch.randperm = real_randperm
    out["ppo/result"] = np.array(res, dtype=np.float64)
    out["ppo/final_lr"] = np.float64(ppo.learning_rate)
    for k, v in ac.state_dict().items():
        out["ppo/final_digest/" + k] = tensor_digest(v.detach().numpy())
    path = os.path.join(HERE, "hlp.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def gen_rollout_case():
    """The rollout loop body of Runner.learn (mini_gym_learn/ppo/__init__.py:126-141) with the reference's own env,
    wrapper, ActorCritic, PPO and RolloutStorage: PPO.act -> HistoryWrapper.step -> PPO.process_env_step for T steps.
    The actions fed to the env are a FIXED sequence (so the env side of the transitions can be compared tightly while
    the policy outputs carry their bf16 tolerance); the simulator state is rewritten before every step (scripted
    physics); a fixed `time_outs` pattern exercises the gamma * V bootstrap of ppo.py:81-83; observation noise is off
    and no DOF-property re-draw falls into the window, so nothing is random."""
    import torch
    import harness
    import statekit
    torch.manual_seed(0)
    torch.set_num_threads(1)

    def hook(Cfg):
        Cfg.noise.add_noise = False
    env, Cfg = harness.make_reference_env("mini_cheetah", N_ENVS, cfg_hook=hook)
    e = env.env
    N, NB, T_ = e.num_envs, e.num_bodies, 6
    from cases import learner_weights
    from mini_gym_learn.ppo import ActorCritic
    from mini_gym_learn.ppo.ppo import PPO
    ac = ActorCritic(42, 18, 630, 12)
    ac.load_state_dict({k: torch.from_numpy(v) for k, v in learner_weights().items()})
    ppo = PPO(ac, device="cpu")
    ppo.init_storage(N, T_, [42], [18], [630], [12])
    rng = np.random.default_rng(31)
    T = lambda a, dt=torch.float: torch.from_numpy(np.asarray(a)).to(dt)
    out = {}
    # ---- persistent state, observation buffers and history at the start of the rollout ----
    e.commands[:, :3] = T(rng.uniform(-1, 1, (N, 3)).astype(np.float32))
    e.last_actions[:] = T(rng.normal(0, 1, (N, 12)).astype(np.float32))
    e.last_dof_vel[:] = T(rng.normal(0, 3, (N, 12)).astype(np.float32))
    e.motor_strengths[:] = T(rng.uniform(0.9, 1.1, (N, 1)).astype(np.float32)).repeat(1, 12)
    e.friction_coeffs[:] = T(rng.uniform(0.05, 4.5, N).astype(np.float32))
    e.restitutions[:] = T(rng.uniform(0, 1, N).astype(np.float32))
    e.payloads[:] = T(rng.uniform(-1, 3, N).astype(np.float32))
    e.com_displacements[:] = T(rng.uniform(-0.1, 0.1, (N, 3)).astype(np.float32))
    e.feet_air_time[:] = T((rng.uniform(0, 0.6, (N, 4)) * (rng.random((N, 4)) < 0.7)).astype(np.float32))
    e.last_contacts = T(rng.random((N, 4)) < 0.3, torch.bool)
    e.episode_length_buf[:] = T(rng.integers(0, 200, N), torch.long)
    for k in e.episode_sums:
        e.episode_sums[k][:] = 0.
    for k in e.command_sums:
        e.command_sums[k][:] = 0.
    e.obs_buf[:] = T(rng.normal(0, 1, (N, 42)).astype(np.float32))
    e.privileged_obs_buf[:] = T(rng.uniform(-1, 1, (N, 18)).astype(np.float32))
    env.obs_history[:] = 0.
    for k, v in statekit.state_from_reference(e).items():
        out["before/" + k] = v
    out["before/obs_buf"] = e.obs_buf.numpy().copy(); out["before/privileged_obs_buf"] = e.privileged_obs_buf.numpy().copy()
    time_outs = T(np.arange(N) % 5 == 0, torch.bool)
    bins = T(rng.integers(0, 5202, N).astype(np.float32))
    out["time_outs"] = time_outs.numpy(); out["env_bins"] = bins.numpy()
    t = Cfg.terrain
    span_x, span_y = t.terrain_length * t.num_rows, t.terrain_width * t.num_cols
    feet = e.feet_indices.tolist(); term = e.termination_contact_indices.tolist()
    dflt = e.default_dof_pos[0].numpy()
    real_normal = torch.normal
    torch.normal = lambda mean, std, *a, **k: mean.clone() if torch.is_tensor(mean) else real_normal(mean, std, *a, **k)
    try:
        od = env.get_observations()
        for s_ in range(T_):
            root, dof, con, actions = synth_inputs(rng, e, N, NB, dflt, feet, term, span_x, span_y, 0.30)
            root[:, 0] = np.clip(root[:, 0], 3.0, span_x - 3.0); root[:, 1] = np.clip(root[:, 1], 3.0, span_y - 3.0)   # no teleport
            e.all_root_states[:] = T(root); e.all_dof_state[:] = T(dof.reshape(-1, 2)); e.all_contact_forces[:] = T(con.reshape(-1, 3))
            out["step%d/root_states" % s_] = root; out["step%d/dof_state" % s_] = dof
            out["step%d/contact_forces" % s_] = con; out["step%d/actions" % s_] = actions
            with torch.inference_mode():
                ppo.act(od["obs"], od["privileged_obs"], od["obs_history"])
                ppo.transition.actions = T(actions)                       # the fixed sequence replaces the sample
                od, rew, done, infos = env.step(T(actions))
                ppo.process_env_step(rew, done, {"env_bins": bins, "time_outs": time_outs})
    finally:
        torch.normal = real_normal
    st = ppo.storage
    for name in ("observations", "privileged_observations", "observation_histories", "actions", "rewards", "dones", "values",
                 "actions_log_prob", "mu", "sigma", "env_bins"):
        out["storage/" + name] = getattr(st, name).numpy().copy()
    out["final/obs"] = od["obs"].numpy().copy(); out["final/obs_history"] = od["obs_history"].numpy().copy()
    path = os.path.join(HERE, "rollout.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def gen_checkpoint_case():
    """The trained policy the reference ships (runs/rapid-locomotion/example/train/201852.132488/checkpoints/
    ac_weights_last.pt) through the reference's own ActorCritic: teacher / student actions and values on a seeded batch
    (actor_critic.py:149-173), and the composition scripts/play.py deploys (body(cat(obs, adaptation_module(hist))))."""
    import torch
    import harness
    harness.install()
    import isaacgym  # noqa: F401  (fake)
    from mini_gym_learn.ppo import ActorCritic
    from cases import tensor_digest
    path = os.path.join(harness.REFERENCE_ROOT, "runs/rapid-locomotion/example/train/201852.132488/checkpoints/ac_weights_last.pt")
    sd = torch.load(path, map_location="cpu", weights_only=True)
    ac = ActorCritic(42, 18, 630, 12)
    ac.load_state_dict(sd)
    ac.eval()
    g = torch.Generator().manual_seed(5)
    n = 512
    obs, priv, hist = torch.randn(n, 42, generator=g), torch.rand(n, 18, generator=g) * 2 - 1, torch.randn(n, 630, generator=g) * 0.5
    out = {"input_digest": np.concatenate([tensor_digest(x.numpy()) for x in (obs, priv, hist)])}    # inputs come from the seed
    with torch.no_grad():
        out["act_teacher"] = ac.act_teacher(obs, priv).numpy()
        out["act_student"] = ac.act_student(obs, hist).numpy()
        out["evaluate"] = ac.evaluate(obs, priv).numpy()
        out["act_inference"] = ac.act_inference({"obs": obs, "obs_history": hist, "privileged_obs": priv}).numpy()
        out["latent_student"] = ac.adaptation_module(hist).numpy()
    for k, v in sd.items():
        out["digest/" + k] = tensor_digest(v.numpy())
    out["std"] = sd["std"].numpy()
    path = os.path.join(HERE, "checkpoint.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def gen_curriculum_uniform_case():
    """_update_command_curriculum_uniform (legged_robot.py:851-880; called from reset_idx :827): the command ranges widen by
    0.2 per call while the mean tracking reward of the reset envs clears the threshold, only on steps that are multiples
    of max_episode_length, clipped at the configured maxima.  A scripted sequence of calls; the ranges after every call
    are the known answers."""
    import torch
    import harness

    def hook(Cfg):
        Cfg.commands.command_curriculum = True
        Cfg.commands.yaw_command_curriculum = True
        Cfg.commands.max_forward_curriculum = 1.5
        Cfg.commands.max_reverse_curriculum = 0.5
        Cfg.commands.max_yaw_curriculum = 1.3
    env, Cfg = harness.make_reference_env("mini_cheetah", N_ENVS, cfg_hook=hook, history=False)
    e = env
    N = e.num_envs
    rng = np.random.default_rng(41)
    L = float(e.cfg.env.max_episode_length)
    lin_thr = Cfg.commands.forward_curriculum_threshold * e.reward_scales["tracking_lin_vel"]
    ang_thr = Cfg.commands.yaw_curriculum_threshold * e.reward_scales["tracking_ang_vel"]
    out = {"max_episode_length": np.float64(L), "lin_threshold": np.float64(lin_thr), "ang_threshold": np.float64(ang_thr)}
    out["ranges0"] = np.array([Cfg.command_ranges["lin_vel_x"], Cfg.command_ranges["ang_vel_yaw"]], dtype=np.float64)
    script = []
    # (step counter multiple?, lin factor of threshold, ang factor): 12 calls
    plan = [(1, 1.3, 0.7), (0, 1.5, 1.5), (2, 1.2, 1.4), (3, 0.9, 1.2), (4, 1.1, 0.9), (5, 1.4, 1.6), (6, 1.05, 1.05), (7, 2.0, 2.0),
            (8, 1.5, 1.5), (9, 1.5, 1.5), (10, 1.5, 1.5), (11, 0.5, 1.5)]
    for i, (mult, lf, af) in enumerate(plan):
        ids = np.sort(rng.choice(N, N // 2, replace=False))
        lin = (rng.uniform(0.9, 1.1, N) * lf * lin_thr * L).astype(np.float32)
        ang = (rng.uniform(0.9, 1.1, N) * af * ang_thr * L).astype(np.float32)
        e.episode_sums["tracking_lin_vel"][:] = torch.from_numpy(lin)
        e.episode_sums["tracking_ang_vel"][:] = torch.from_numpy(ang)
        e.common_step_counter = int(mult * L) if mult else int(3 * L) + 7
        e._update_command_curriculum_uniform(torch.from_numpy(ids), e.cfg, e.episode_sums)
        out["call%d/ids" % i] = ids; out["call%d/lin" % i] = lin; out["call%d/ang" % i] = ang
        out["call%d/step" % i] = np.int64(e.common_step_counter)
        out["call%d/ranges" % i] = np.array([Cfg.command_ranges["lin_vel_x"], Cfg.command_ranges["ang_vel_yaw"]], dtype=np.float64)
    out["n_calls"] = np.int64(len(plan))
    path = os.path.join(HERE, "curriculum_uniform.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), out["ranges0"].tolist(), "->", out["call%d/ranges" % (len(plan) - 1)].tolist())


if __name__ == "__main__":
    which = sys.argv[1:] or None
    if which is None:
        for c in CASES:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), c])
    else:
        for c in which:
            if c == "learner":
                gen_learner_case()
            elif c == "rollout":
                gen_rollout_case()
            elif c == "checkpoint":
                gen_checkpoint_case()
            elif c == "curriculum_uniform":
                gen_curriculum_uniform_case()
            elif c == "eval_split":
                gen_eval_split_case()
            elif c == "hlp":
                gen_hlp_case()
            else:
                gen_env_case(c)
