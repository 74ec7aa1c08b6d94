"""Runner: rollout -> GAE -> PPO.update loop on the product kernels, eagerly and with the
24-step rollout captured in CUDA graphs (one per history-ring phase)."""
import pytest
import torch

from cases import build_case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
DEV = "cuda:0"


def _make(n_envs, graph):
    from rapid_locomotion_rl_b200.envs import HistoryWrapper, VelocityTrackingEasyEnv
    from rapid_locomotion_rl_b200.ppo import Runner
    cfg, robot, terrain = build_case("mc_flat", n_envs)
    env = HistoryWrapper(VelocityTrackingEasyEnv(sim_device=DEV, headless=True, cfg=cfg, terrain=terrain, seed=3))

    def physics(e):           # stands in for the simulator: the state tensors move between steps
        inner = getattr(e, "env", e)
        inner.sim.dof_state.add_(torch.randn_like(inner.sim.dof_state) * 0.01)
        inner.sim.root_states[:, 7:13].add_(torch.randn_like(inner.sim.root_states[:, 7:13]) * 0.01)
    torch.manual_seed(0)
    return Runner(env, device=DEV, graph_rollout=graph, physics=physics), env


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
