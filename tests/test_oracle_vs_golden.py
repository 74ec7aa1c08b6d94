"""Pins the oracle: oracle/env_oracle.py must reproduce, bit for bit on the CPU, what the
unmodified reference produced (tests/golden/env_*.npz, written by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

import statekit
from cases import ENV_CASES, build_case
from oracle.env_oracle import OracleCurriculum, OracleEnv, resample_commands


def load(golden_dir, case):
    return dict(np.load(os.path.join(golden_dir, "env_%s.npz" % case), allow_pickle=False))


def sub(g, prefix):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


@pytest.fixture(scope="module", autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("case", ENV_CASES)
def test_frozen_constants(golden_dir, case):
    g = load(golden_dir, case)
    cfg, robot, terrain = build_case(case, 48)
    o = OracleEnv(cfg, robot, terrain)
    assert o.dt == float(g["const/dt"])
    assert o.max_episode_length == float(g["const/max_episode_length"])
    assert o.rand_interval == int(g["const/rand_interval"]) and o.push_interval == int(g["const/push_interval"])
    for name in ("p_gains", "d_gains", "default_dof_pos", "torque_limits", "dof_pos_limits", "dof_vel_limits",
                 "noise_scale_vec", "feet_indices", "termination_contact_indices", "penalised_contact_indices"):
        assert np.array_equal(getattr(o, name).numpy(), g["const/" + name]), name
    assert o.reward_names == list(g["const/reward_names"])
    assert list(o.reward_scales.keys()) == list(g["const/reward_scale_keys"])
    assert np.array_equal(np.array(list(o.reward_scales.values())), g["const/reward_scales"])
    # the product's freeze step resolves the same numbers
    from rapid_locomotion_rl_b200.config import freeze_env_cfg
    p = freeze_env_cfg(cfg, robot, terrain)
    assert p.reward_names == o.reward_names
    assert np.array_equal(np.float32(p.p_gains), g["const/p_gains"])
    assert np.array_equal(np.float32(p.default_dof_pos), g["const/default_dof_pos"][0])
    assert np.array_equal(np.float32(p.dof_pos_lo), g["const/dof_pos_limits"][:, 0])
    assert np.array_equal(np.float32(p.dof_pos_hi), g["const/dof_pos_limits"][:, 1])
    assert np.array_equal(p.noise_scale_vec, g["const/noise_scale_vec"])
    assert p.max_episode_length == 1001 or case == "none"
    assert np.array_equal(np.float32(p.term_scale[:p.n_terms]),
                          np.float32([o.reward_scales[n] for n in o.reward_names]))


@pytest.mark.parametrize("case", ENV_CASES)
def test_step_bit_exact(golden_dir, case):
    g = load(golden_dir, case)
    cfg, robot, terrain = build_case(case, 48)
    if case == "mc_rough":
        assert np.array_equal(terrain.heightsamples, g["heightsamples"])
    o = OracleEnv(cfg, robot, terrain)
    for s in range(3):
        pre = "step%d/" % s
        statekit.apply_to_oracle(o, sub(g, pre + "before/"))
        obs, priv, rew, reset = o.step(torch.from_numpy(g[pre + "actions"]), noise_u=torch.from_numpy(g[pre + "noise_u"]),
                                       dr_u=torch.from_numpy(g[pre + "dr_u"]), push_u=torch.from_numpy(g[pre + "push_u"]))
        assert np.array_equal(obs.numpy(), g[pre + "obs"]), "obs step %d" % s
        assert np.array_equal(priv.numpy(), g[pre + "priv"])
        assert np.array_equal(rew.numpy(), g[pre + "rew"])
        assert np.array_equal(reset.numpy(), g[pre + "reset"])
        if case == "mc_rough":
            assert np.array_equal(o.measured_heights.numpy(), g[pre + "measured_heights"])
        got = statekit.state_from_oracle(o)
        want = sub(g, pre + "after/")
        statekit.assert_state_close(got, want, rtol=0, atol=0, exact_keys=tuple(want.keys()),
                                    skip=("contact_forces",), label="%s step %d" % (case, s))


@pytest.mark.parametrize("case", ENV_CASES)
def test_reset_bit_exact(golden_dir, case):
    g = load(golden_dir, case)
    cfg, robot, terrain = build_case(case, 48)
    o = OracleEnv(cfg, robot, terrain)
    statekit.apply_to_oracle(o, sub(g, "reset/before/"))
    o.reset_buf = torch.from_numpy(g["step2/reset"].copy())   # the flags the last step left behind
    ids = torch.from_numpy(g["reset/ids"])
    means = o.reset_idx(ids, dr_u=torch.from_numpy(g["reset/dr_u"]), init_u=torch.from_numpy(g["reset/init_u"]),
                        level_u=torch.from_numpy(g["reset/level_u"]))
    want = sub(g, "reset/after/")
    got = statekit.state_from_oracle(o)
    statekit.assert_state_close(got, want, rtol=0, atol=0, exact_keys=tuple(want.keys()),
                                skip=("contact_forces", "torques", "base_lin_vel", "base_ang_vel", "projected_gravity",
                                      "last_root_vel", "joint_pos_target"), label=case + " reset")
    assert np.array_equal(o.reset_buf.numpy(), g["reset/reset_buf"])
    for k, v in sub(g, "reset/extras/").items():
        assert np.float32(means[k]) == v, k


@pytest.mark.parametrize("case", ["mc_flat", "go1"])
def test_resample_bit_exact(golden_dir, case):
    g = load(golden_dir, case)
    cfg, robot, terrain = build_case(case, 48)
    o = OracleEnv(cfg, robot, terrain)
    c = cfg.commands
    cur = OracleCurriculum(c.curriculum_seed, x_vel=(c.limit_vel_x[0], c.limit_vel_x[1], 51),
                           y_vel=(c.limit_vel_y[0], c.limit_vel_y[1], 2), yaw_vel=(c.limit_vel_yaw[0], c.limit_vel_yaw[1], 51))
    cur.set_to([c.lin_vel_x[0], c.lin_vel_y[0], c.ang_vel_yaw[0]], [c.lin_vel_x[1], c.lin_vel_y[1], c.ang_vel_yaw[1]])
    assert cur.weights.sum() == 30.0  # 30/5202 = 0.006 "command area" of the reference's run log
    for rnd in range(2):
        pre = "resample%d/" % rnd
        cur.weights = g[pre + "before/weights"].copy()
        bins = g[pre + "before/bins"].copy()
        o.commands = torch.from_numpy(g[pre + "before/commands"].copy())
        for k in o.command_sums:
            o.command_sums[k] = torch.from_numpy(g[pre + "before/command_sums/" + k].copy())
        ids = torch.from_numpy(g[pre + "ids"])
        resample_commands(o, cur, bins, ids, u_bin=g[pre + "u_bin"], u_cell=g[pre + "u_cell"])
        assert np.array_equal(cur.weights, g[pre + "after/weights"])
        assert np.array_equal(bins, g[pre + "after/bins"])
        assert np.array_equal(o.commands.numpy(), g[pre + "after/commands"])
        for k in o.command_sums:
            assert np.array_equal(o.command_sums[k].numpy(), g[pre + "after/command_sums/" + k]), k
        assert (g[pre + "after/weights"] != g[pre + "before/weights"]).any(), "fixture exercises the weight update"


def test_resample_mt19937_stream(golden_dir):
    """Without injected uniforms the oracle draws from RandomState(seed) in the reference's order."""
    g = load(golden_dir, "mc_flat")
    r = np.random.RandomState(100)
    n = 48
    r.random_sample(n); [r.random_sample(3) for _ in range(n)]   # the initial all-env resample
    ids = g["resample0/ids"]
    assert np.array_equal(r.random_sample(len(ids)), g["resample0/u_bin"][ids])
