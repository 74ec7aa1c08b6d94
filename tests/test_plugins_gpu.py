"""Plugin surface of the drop-in env (reference: mini_gym/envs/base/legged_robot.py:1074-1093 `_reward_<name>`
resolution, :190 check_termination, :314 compute_reward, :342 compute_observations, :1469 _get_heights): user-defined reward
terms and overridden hooks take effect on top of the fused launch."""
import numpy as np
import pytest
import torch

import statekit
from cases import build_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _state(env, robot, n, seed=5):
    from rapid_locomotion_rl_b200.sim import synthetic_state
    p = env.params
    st = synthetic_state(seed, n, robot.num_bodies, 12, np.float32(p.default_dof_pos), p.feet_idx, p.term_idx[:p.n_term_bodies],
                         z0=0.3)
    rng = np.random.default_rng(seed)
    st["commands"] = np.concatenate([rng.uniform(-1, 1, (n, 3)), np.zeros((n, 1))], 1).astype(np.float32)
    st["last_actions"] = rng.normal(0, 1, (n, 12)).astype(np.float32)
    st["episode_length_buf"] = rng.integers(0, 200, n).astype(np.int64)
    return st, rng


def _make(cls, case, n, hook=None):
    cfg, robot, terrain = build_case(case, n)
    if hook:
        hook(cfg)
    env = cls(cfg, sim_device=DEV, headless=True, terrain=terrain)
    return env, robot


def _step(env, robot, n, seed=5):
    st, rng = _state(env, robot, n, seed)
    statekit.apply_to_product(env, st)
    actions = torch.from_numpy(rng.normal(0, 1, (n, 12)).astype(np.float32)).to(DEV)
    noise = torch.from_numpy(rng.random((n, env.num_obs)).astype(np.float32)).to(DEV)
    env._inject = dict(noise_u=noise)
    out = [t.clone() for t in env.step(actions)[:4]]
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("n", [96, 1000])          # 96: all-TMA kernel (multiple of 32); 1000: one-warp-per-leg kernel
@pytest.mark.parametrize("positive", [False, True])
def test_user_defined_reward_term(n, positive):
    from rapid_locomotion_rl_b200.envs import LeggedRobot

    class WithFoo(LeggedRobot):
        def _reward_foo(self):
            return torch.square(self.base_lin_vel[:, 0]) + 0.1 * self.commands[:, 2] + 3.0

    def hook(scale):
        def h(cfg):
            cfg.rewards.only_positive_rewards = positive
            if scale is not None:
                cfg.rewards.scales.foo = scale
        return h
    base, robot = _make(LeggedRobot, "mc_flat", n, hook(None))
    plug, _ = _make(WithFoo, "mc_flat", n, hook(-0.5))
    assert "foo" in plug.reward_names and "foo" in plug.episode_sums and "foo" in plug.command_sums
    ob, pb, rb, db = _step(base, robot, n)
    op, pp, rp, dp = _step(plug, robot, n)
    assert torch.equal(ob, op) and torch.equal(pb, pp) and torch.equal(db, dp)
    term = (torch.square(plug.base_lin_vel[:, 0]) + 0.1 * plug.commands[:, 2] + 3.0) * np.float32(-0.5 * plug.dt)
    torch.testing.assert_close(plug.episode_sums["foo"], term, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(plug.command_sums["foo"], term, rtol=1e-6, atol=1e-7)
    raw = plug._rew_raw
    want = torch.clip(raw + term, min=0.) if positive else raw + term
    torch.testing.assert_close(rp, want, rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(rb, torch.clip(raw, min=0.) if positive else raw, rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(plug.episode_sums["total"], want, rtol=1e-5, atol=2e-6)
    assert (rp != rb).any()
    # the accumulators of a plugin term are reported and zeroed on reset like the built-in ones (:264-267)
    ids = torch.arange(0, n, 3, device=DEV)
    mean = plug.episode_sums["foo"][ids].mean().item()
    plug.reset_idx(ids)
    assert abs(float(plug.extras["train/episode"]["rew_foo"]) - mean) < 1e-6
    assert float(plug.episode_sums["foo"][ids].abs().max()) == 0.0 and float(plug.episode_sums["foo"][1].abs()) > 0


def test_missing_reward_method_raises_like_the_reference():
    from rapid_locomotion_rl_b200.envs import LeggedRobot
    cfg, robot, terrain = build_case("mc_flat", 32)
    cfg.rewards.scales.bar = 1.0
    with pytest.raises(AttributeError, match="_reward_bar"):
        LeggedRobot(cfg, sim_device=DEV, terrain=terrain)


def test_overridden_builtin_term_runs_in_python_and_matches_the_kernel():
    from rapid_locomotion_rl_b200.envs import LeggedRobot

    class MyTorques(LeggedRobot):
        calls = 0

        def _reward_torques(self):                       # legged_robot.py:1523 restated by the user
            type(self).calls += 1
            return torch.sum(torch.square(self.torques), dim=1)
    n = 256
    base, robot = _make(LeggedRobot, "mc_flat", n)
    mine, _ = _make(MyTorques, "mc_flat", n)
    assert mine.params.n_terms == base.params.n_terms - 1 and [t[0] for t in mine._custom_terms] == ["torques"]
    _, _, rb, _ = _step(base, robot, n)
    _, _, rm, _ = _step(mine, robot, n)
    assert MyTorques.calls >= 1
    torch.testing.assert_close(rm, rb, rtol=1e-5, atol=2e-6)
    torch.testing.assert_close(mine.episode_sums["torques"], base.episode_sums["torques"], rtol=1e-5, atol=1e-7)


def test_overridden_hooks_take_effect():
    from rapid_locomotion_rl_b200.envs import LeggedRobot

    class Hooked(LeggedRobot):
        def check_termination(self):
            super().check_termination()
            self.reset_buf |= self.root_states[:, 2] < 0.30

        def compute_observations(self):
            super().compute_observations()
            self.obs_buf[:, 0] += 1.0

        def compute_reward(self):
            super().compute_reward()
            self.rew_buf *= 2.0
    n = 512
    base, robot = _make(LeggedRobot, "mc_flat", n)
    hk, _ = _make(Hooked, "mc_flat", n)
    ob, pb, rb, db = _step(base, robot, n)
    oh, ph, rh, dh = _step(hk, robot, n)
    low = hk.root_states[:, 2] < 0.30
    assert low.any() and (~low).any()
    assert torch.equal(dh, db | low) and (dh != db).any()
    torch.testing.assert_close(oh[:, 0], ob[:, 0] + 1.0)
    assert torch.equal(oh[:, 1:], ob[:, 1:])
    torch.testing.assert_close(rh, 2.0 * rb)

    class BadHeights(LeggedRobot):
        def _get_heights(self, env_ids=None, cfg=None):
            return None
    cfg, _, terrain = build_case("mc_flat", 32)
    with pytest.raises(NotImplementedError):
        BadHeights(cfg, sim_device=DEV, terrain=terrain)


def test_get_heights_is_a_real_method():
    """_get_heights() (:1469-1503) as its own launch equals what the fused step measured, for all envs and for a subset."""
    from rapid_locomotion_rl_b200.envs import LeggedRobot
    n = 640
    env, robot = _make(LeggedRobot, "mc_rough", n)
    st, rng = _state(env, robot, n)
    st["root_states"][:, :2] = rng.uniform(1.0, 15.0, (n, 2))          # over the 2 x 2 test terrain, away from the teleport band
    statekit.apply_to_product(env, st)
    env.step(torch.zeros(n, 12, device=DEV))
    torch.cuda.synchronize()
    h = env._get_heights()
    assert h.shape == (n, 187) and torch.equal(h, env.measured_heights)
    ids = torch.tensor([5, 77, 639, 0], device=DEV)
    assert torch.equal(env._get_heights(ids), env.measured_heights[ids])
    flat, _ = _make(LeggedRobot, "mc_flat", 32)
    assert float(flat._get_heights().abs().max()) == 0.0
