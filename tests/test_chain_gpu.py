"""Fused MLP chain kernel (csrc/chain.cu) against the per-layer tcgen05 GEMM path and fp32 torch.

The per-layer path is itself checked against fp32 autograd in test_gemm_gpu.py / test_ppo_gpu.py; here
the chain must reproduce its bf16 activations (same rounding points: bf16 operands, fp32 accumulate,
bf16 stored activations) and the fp32 network outputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ac(seed=0):
    from rapid_locomotion_rl_b200.ppo import ActorCritic
    torch.manual_seed(seed)
    return ActorCritic(42, 18, 630, 12, device="cuda:0")


def _ref_teacher(ac, obs, priv):
    import torch.nn.functional as F
    bfr = lambda x: x.to(torch.bfloat16).float()
    def mlp(seq, x):
        lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
        for i, l in enumerate(lin):
            x = F.linear(x, bfr(l.weight), l.bias)
            if i < len(lin) - 1:
                x = bfr(F.elu(x))
        return x
    lat = bfr(mlp(ac.env_factor_encoder, bfr(priv)))
    x = torch.cat((bfr(obs), lat), dim=-1)
    return mlp(ac.actor_body, x), mlp(ac.critic_body, x)


@pytest.mark.parametrize("rows", [1, 128, 300, 4000, 24000])
@pytest.mark.parametrize("save", [False, True])
def test_teacher_forward_chain_matches_layers(rows, save):
    ac = _ac()
    g = torch.Generator(device="cuda").manual_seed(rows)
    obs = torch.randn(rows, 42, device="cuda", generator=g)
    priv = torch.rand(rows, 18, device="cuda", generator=g) * 2 - 1
    ac.workspace(rows, backward=save)
    outs = {}
    for use_chain in (False, True):
        ac.use_chain = use_chain
        w = ac.workspace(rows, backward=save)
        for k in ("H1", "H2", "Y1", "A2", "A3", "C2", "C3", "mean", "value"):
            w[k].zero_()
        ac._stage_obs(obs, rows)
        ac._stage(priv, w["Xp"], ac.num_priv, 0, w["Xp"].shape[1])
        ac.forward_teacher(rows, save=save)
        torch.cuda.synchronize()
        outs[use_chain] = {k: w[k][:rows].float().clone() for k in ("Xac", "H1", "H2", "Y1", "A2", "A3", "C2", "C3", "mean", "value")}
    mean_ref, value_ref = _ref_teacher(ac, obs, priv)
    torch.testing.assert_close(outs[True]["mean"], mean_ref, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(outs[True]["value"], value_ref, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(outs[True]["mean"], outs[False]["mean"], rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(outs[True]["value"], outs[False]["value"], rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(outs[True]["Xac"], outs[False]["Xac"], rtol=1e-2, atol=1e-2)
    if save:
        for k in ("H1", "H2", "Y1", "A2", "A3", "C2", "C3"):
            # identical rounding points; a bf16 ulp flip where the fp32 sums differ in the last bit
            torch.testing.assert_close(outs[True][k], outs[False][k], rtol=2e-2, atol=2e-2)
            assert (outs[True][k] != outs[False][k]).float().mean() < 0.02, k


def _random_storage(ppo, n_envs, T, seed):
    st = ppo.storage
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda t, s=1.0: t.copy_(torch.randn(t.shape, device="cuda", generator=g) * s)
    r(st.observations); st.privileged_observations.copy_(torch.rand(st.privileged_observations.shape, device="cuda", generator=g) * 2 - 1)
    r(st.observation_histories); r(st.actions); r(st.values); r(st.returns); r(st.advantages)
    st.actions_log_prob.fill_(-17.0); r(st.mu, 0.3); st.sigma.fill_(1.0)


@pytest.mark.parametrize("n_envs", [11, 1000])
def test_minibatch_step_chain_matches_layers(n_envs):
    """One full PPO minibatch step (forward, loss, dgrad chains, wgrads, adaptation step) on both learner
    paths from identical weights: same loss statistics and the same flat gradient."""
    from rapid_locomotion_rl_b200.ppo import PPO
    T = 24
    res = {}
    for use_chain in (False, True):
        ac = _ac(1)
        ac.use_chain = use_chain
        ppo = PPO(ac, device="cuda:0")
        ppo.init_storage(n_envs, T, [42], [18], [630], [12])
        _random_storage(ppo, n_envs, T, 5)
        ppo.debug_keep_grad = True
        idx = torch.randperm(n_envs * T, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
        ppo.minibatch_step(idx)
        torch.cuda.synchronize()
        w = ac._ws
        res[use_chain] = dict(grad=ppo.debug_grad.clone(), stats=ppo.debug_stats.clone(), flat=ac.flat.clone(),
                              acts={k: w[k][:n_envs * T].float().clone() for k in ("dA3", "dA2", "dY1", "dC3", "dC2", "dLat", "dH2", "dH1", "dD2", "dD1", "D1", "D2", "pred")})
    for k, a in res[True]["acts"].items():
        b = res[False]["acts"][k]
        rel = ((a - b).norm() / (b.norm() + 1e-20)).item()
        assert rel < 2e-2, (k, rel)
    ga, gb = res[True]["grad"], res[False]["grad"]
    cos = torch.nn.functional.cosine_similarity(ga, gb, dim=0).item()
    rel = ((ga - gb).norm() / gb.norm()).item()
    assert cos > 0.9995 and rel < 3e-2, (cos, rel)
    torch.testing.assert_close(res[True]["stats"], res[False]["stats"], rtol=2e-3, atol=1e-6)
    torch.testing.assert_close(res[True]["flat"], res[False]["flat"], rtol=0, atol=2.5e-3)   # one Adam step: |dw| <= lr
