"""Runner: rollout -> GAE -> PPO.update loop on the product kernels, eagerly and with the
24-step rollout captured in CUDA graphs (one per history-ring phase)."""
import pytest
import torch

from cases import build_case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
DEV = "cuda:0"


def _make(n_envs, graph, **kw):
    from rapid_locomotion_rl_b200.envs import HistoryWrapper, VelocityTrackingEasyEnv
    from rapid_locomotion_rl_b200.ppo import Runner
    cfg, robot, terrain = build_case("mc_flat", n_envs)
    env = HistoryWrapper(VelocityTrackingEasyEnv(sim_device=DEV, headless=True, cfg=cfg, terrain=terrain, seed=3))

    def physics(e):           # stands in for the simulator: the state tensors move between steps
        inner = getattr(e, "env", e)
        inner.sim.dof_state.add_(torch.randn_like(inner.sim.dof_state) * 0.01)
        inner.sim.root_states[:, 7:13].add_(torch.randn_like(inner.sim.root_states[:, 7:13]) * 0.01)
    torch.manual_seed(0)
    return Runner(env, device=DEV, graph_rollout=graph, physics=physics, **kw), env


@pytest.mark.parametrize("graph", [False, True])
def test_runner_learn(graph):
    runner, env = _make(256, graph)
    w0 = runner.alg.actor_critic.flat.clone()
    seen = []
    hist = runner.learn(8, log=seen.append)          # 8 iterations: every ring phase (5) and graph re-use
    assert len(hist) == 8 and seen == hist
    for rec in hist:
        for k in ("mean_value_loss", "mean_surrogate_loss", "adaptation_loss"):
            assert rec[k] == rec[k] and abs(rec[k]) < 1e6, (k, rec)
    assert runner.tot_timesteps == 8 * 24 * 256 and runner.current_learning_iteration == 8
    assert not torch.equal(w0, runner.alg.actor_critic.flat) and torch.isfinite(runner.alg.actor_critic.flat).all()
    if graph:
        assert len(runner._graphs) == 5              # lcm(24, 15) / 24 ring phases
    # the history view ends with the newest observation and the rollout filled the storage
    od = env.get_observations()
    torch.testing.assert_close(od["obs_history"][:, -42:], od["obs"])
    inner = env.env
    assert inner.common_step_counter >= 8 * 24


def test_graph_rollout_writes_what_eager_writes():
    """Same env state and weights: one eager rollout and one graph-replayed rollout store identical
    observations / rewards / dones when the action noise is the same (zero std makes it so)."""
    outs = {}
    for graph in (False, True):
        runner, env = _make(128, graph)
        runner.physics = None
        ac = runner.alg.actor_critic
        ac.std.data.zero_()
        od = env.get_observations()
        obs, priv, hist = od["obs"], od["privileged_obs"], od["obs_history"]
        for _ in range(3):                            # iteration 1 eager warm-up, 2 captures, 3 replays a new phase
            with torch.inference_mode():
                obs, priv, hist = runner._rollout(obs, priv, hist)
            st = runner.alg.storage
            snap = {k: getattr(st, k).clone() for k in ("observations", "privileged_observations", "observation_histories",
                                                        "actions", "rewards", "dones", "values", "mu")}
            st.clear()
        torch.cuda.synchronize()
        outs[graph] = snap
    for k in outs[False]:
        torch.testing.assert_close(outs[True][k], outs[False][k], rtol=1e-5, atol=1e-6, msg=k)


def test_add_transitions_fused_copy():
    """RolloutStorage.add_transitions (rollout_storage.py:54-71) as one launch: every field of the [t] slice
    equals the source, including the strided history-ring view and the bool dones."""
    from rapid_locomotion_rl_b200.ppo import RolloutStorage
    N, T = 333, 4
    st = RolloutStorage(N, T, [42], [18], [630], [12], DEV)
    g = torch.Generator(device=DEV).manual_seed(0)
    ring = torch.randn(N, 2 * 630, device=DEV, generator=g)
    for s in range(T):
        t = RolloutStorage.Transition()
        t.observations = torch.randn(N + 5, 42, device=DEV, generator=g)[:N]
        t.critic_observations = t.observations
        t.privileged_observations = torch.randn(N, 18, device=DEV, generator=g)
        t.observation_histories = ring[:, 42 * (s + 1):42 * (s + 1) + 630]        # row-strided view, like HistoryWrapper's
        t.actions = torch.randn(N, 12, device=DEV, generator=g)
        t.rewards = torch.randn(N, device=DEV, generator=g)
        t.dones = torch.rand(N, device=DEV, generator=g) < 0.3
        t.values = torch.randn(N, 1, device=DEV, generator=g)
        t.actions_log_prob = torch.randn(N, device=DEV, generator=g)
        t.action_mean = torch.randn(N, 12, device=DEV, generator=g)
        t.action_sigma = torch.rand(N, 12, device=DEV, generator=g)
        t.env_bins = torch.randint(0, 5202, (N,), device=DEV, generator=g).float()
        st.add_transitions(t)
        torch.cuda.synchronize()
        assert torch.equal(st.observations[s], t.observations) and torch.equal(st.privileged_observations[s], t.privileged_observations)
        assert torch.equal(st.observation_histories[s], t.observation_histories)
        assert torch.equal(st.actions[s], t.actions) and torch.equal(st.mu[s], t.action_mean) and torch.equal(st.sigma[s], t.action_sigma)
        assert torch.equal(st.rewards[s, :, 0], t.rewards) and torch.equal(st.values[s], t.values)
        assert torch.equal(st.actions_log_prob[s, :, 0], t.actions_log_prob) and torch.equal(st.env_bins[s, :, 0], t.env_bins)
        assert torch.equal(st.dones[s, :, 0], t.dones.to(torch.uint8))
    assert st.step == T
    with pytest.raises(AssertionError):
        st.add_transitions(t)


@pytest.mark.parametrize("n_envs,timeouts", [(160, False), (4000, True)])
def test_fused_rollout_glue_stores_what_the_reference_call_order_stores(n_envs, timeouts):
    """Runner with the fused step glue (rl_rollout_boundary / rl_rollout_act: 4-5 launches per step) and Runner driving
    PPO.act -> env.step -> PPO.process_env_step like the reference (mini_gym_learn/ppo/__init__.py:126-141): the same
    env state, weights and (zero) Normal draws must leave identical transitions in the storage - every field, two
    rollouts (the second starts from a shifted history ring).  With time-outs the gamma * V bootstrap of ppo.py:81-83
    is part of the stored reward."""
    outs = {}
    for fused in (False, True):
        runner, env = _make(n_envs, False, fused_rollout=fused)
        runner.physics = None
        inner = env.env
        assert runner._fused_ok() == fused
        if timeouts:
            inner._time_out_u8.copy_((torch.arange(n_envs, device=DEV) % 7 == 0).to(torch.uint8))
            inner.extras["time_outs"] = inner.time_out_buf[:n_envs]
        runner.inject_normal = torch.zeros(n_envs, 12, device=DEV)       # the same (zero) action noise on both paths
        od = env.get_observations()
        obs, priv, hist = od["obs"], od["privileged_obs"], od["obs_history"]
        snaps = []
        for _ in range(2):
            with torch.inference_mode():
                obs, priv, hist = runner._rollout(obs, priv, hist)
            st = runner.alg.storage
            assert st.step == runner.num_steps_per_env
            snaps.append({k: getattr(st, k).clone() for k in (
                "observations", "privileged_observations", "observation_histories", "actions", "rewards", "dones", "values",
                "mu", "sigma", "actions_log_prob", "env_bins")})
            st.clear()
        torch.cuda.synchronize()
        outs[fused] = (snaps, obs.clone(), hist.clone())
    for a, b in zip(outs[True][0], outs[False][0]):
        for k in a:
            torch.testing.assert_close(a[k], b[k], rtol=1e-5, atol=1e-6, msg=k)
        for k in ("observations", "privileged_observations", "observation_histories", "dones", "env_bins"):
            assert torch.equal(a[k], b[k]), k
    assert torch.equal(outs[True][1], outs[False][1]) and torch.equal(outs[True][2], outs[False][2])
    if timeouts:
        a = outs[True][0][0]
        boot = a["rewards"][:, ::7, 0] - 0.99 * a["values"][:, ::7, 0]
        plain = a["rewards"][:, 1::7, 0]
        assert boot.abs().max() < 10 and plain.abs().max() < 10         # finite and of reward magnitude


@pytest.mark.parametrize("fused", [True, False])
def test_rollout_vs_reference_loop(golden_dir, fused):
    """tests/golden/rollout.npz: the reference's own env + HistoryWrapper + PPO + RolloutStorage run through the rollout
    loop body of Runner.learn (mini_gym_learn/ppo/__init__.py:126-141) for 6 steps with a fixed action sequence, scripted
    simulator state and a time-out pattern.  The product Runner (fused glue and reference call order) must leave the same
    transitions in its storage: env-side fields to 1e-5 (dones / bins exactly), policy-side fields to the bf16 tolerance."""
    import os
    import numpy as np
    import statekit
    from rapid_locomotion_rl_b200.envs import HistoryWrapper, VelocityTrackingEasyEnv
    from rapid_locomotion_rl_b200.ppo import Runner, RunnerArgs
    from cases import learner_weights
    g = dict(np.load(os.path.join(golden_dir, "rollout.npz")))
    N, T = g["storage/actions"].shape[1], g["storage/actions"].shape[0]
    cfg, robot, terrain = build_case("mc_flat", N)
    cfg.noise.add_noise = False
    env = HistoryWrapper(VelocityTrackingEasyEnv(sim_device=DEV, headless=True, cfg=cfg, terrain=terrain, seed=3))
    old_T = RunnerArgs.num_steps_per_env
    RunnerArgs.num_steps_per_env = T
    try:
        runner = Runner(env, device=DEV, graph_rollout=False, fused_rollout=fused)
    finally:
        RunnerArgs.num_steps_per_env = old_T
    inner = env.env
    ac = runner.alg.actor_critic
    ac.load_state_dict({k: torch.from_numpy(v) for k, v in learner_weights().items()})
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    before = {k[len("before/"):]: v for k, v in g.items() if k.startswith("before/")}
    statekit.apply_to_product(inner, before)
    inner.obs_buf.copy_(cu(before["obs_buf"])); inner.privileged_obs_buf.copy_(cu(before["privileged_obs_buf"]))
    env._ring.zero_()
    inner._time_out_u8.copy_(cu(g["time_outs"].astype(np.uint8)))
    inner.extras["time_outs"] = inner.time_out_buf[:N]
    inner.extras["env_bins"] = cu(g["env_bins"])

    def set_sim(t):
        if t < T:
            inner.sim.root_states.copy_(cu(g["step%d/root_states" % t]))
            inner.sim.dof_state.copy_(cu(g["step%d/dof_state" % t]).view(-1, 2))
            inner.sim.contact_forces.copy_(cu(g["step%d/contact_forces" % t]).view(-1, 3))
    step = {"t": 0}

    def physics(e):                      # called after every env.step: the simulator state of the NEXT step
        step["t"] += 1
        set_sim(step["t"])

    def actions(t, buf, slot):
        buf.copy_(cu(g["step%d/actions" % t]))
        if slot is not None:
            slot.copy_(buf)
    runner.physics, runner.action_hook = physics, actions
    runner.inject_normal = torch.zeros(N, 12, device=DEV)
    set_sim(0)
    od = env.get_observations()
    with torch.inference_mode():
        obs, priv, hist = runner._rollout(od["obs"], od["privileged_obs"], od["obs_history"])
    torch.cuda.synchronize()
    st = runner.alg.storage
    got = lambda name: getattr(st, name).cpu().numpy()
    for name in ("observations", "privileged_observations", "observation_histories", "actions"):
        np.testing.assert_allclose(got(name), g["storage/" + name], rtol=1e-5, atol=2e-6, err_msg=name)
    assert np.array_equal(got("dones"), g["storage/dones"]) and np.array_equal(got("env_bins"), g["storage/env_bins"])
    for name in ("values", "mu", "actions_log_prob", "sigma"):                      # bf16 policy pass
        np.testing.assert_allclose(got(name), g["storage/" + name], rtol=2e-2, atol=2e-2, err_msg=name)
    # rewards: exact env part, plus gamma * V (bf16 tolerance) where the time-out pattern is set (ppo.py:81-83)
    tmo = g["time_outs"]
    np.testing.assert_allclose(got("rewards")[:, ~tmo], g["storage/rewards"][:, ~tmo], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(got("rewards")[:, tmo], g["storage/rewards"][:, tmo], rtol=2e-2, atol=2e-2)
    boot = got("rewards")[:, tmo, 0] - 0.99 * got("values")[:, tmo, 0]
    plain = g["storage/rewards"][:, tmo, 0] - 0.99 * g["storage/values"][:, tmo, 0]
    np.testing.assert_allclose(boot, plain, rtol=1e-4, atol=1e-5)                   # the env part under the bootstrap
    np.testing.assert_allclose(obs.cpu().numpy(), g["final/obs"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(hist.cpu().numpy(), g["final/obs_history"], rtol=1e-5, atol=2e-6)
