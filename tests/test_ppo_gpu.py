"""Learner kernels (tcgen05 GEMMs + fused loss / Adam) through the drop-in ActorCritic / PPO API against
the pinned fp32 oracle (oracle/ppo_oracle.py) and the reference's golden PPO.update.

Stated tolerance (BASELINE.json north_star "bf16/tf32 tolerance"): operands are rounded to bf16 (8-bit
mantissa, rel 4e-3 per element), accumulation is fp32.  Network outputs: |err| <= 2e-2 + 2e-2*|ref|;
loss scalars 3e-2 relative; gradients: cosine >= 0.995 per weight matrix and relative L2 error <= 8e-2.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ppo_oracle
from test_learner_oracle import learner_case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = "cuda:0"


@pytest.fixture(params=["chain", "layers"], autouse=True)
def learner_path(request, monkeypatch):
    """Every test runs on both learner paths: fused MLP chains (csrc/chain.cu, the default) and one
    tcgen05 GEMM per layer (csrc/gemm_tc.cu)."""
    monkeypatch.setenv("RL_USE_CHAIN", "1" if request.param == "chain" else "0")
    return request.param


def make_ac():
    from cases import learner_weights
    from rapid_locomotion_rl_b200.ppo import ActorCritic
    ac = ActorCritic(42, 18, 630, 12, device=DEV)
    sd = {k: torch.from_numpy(v) for k, v in learner_weights().items()}
    ac.load_state_dict(sd)
    return ac, sd


def test_state_dict_keys_match_reference():
    ac, sd = make_ac()
    keys = list(ac.state_dict().keys())
    assert keys[0] == "std" and len(keys) == 35
    assert set(keys) == set(sd.keys())          # includes the duplicate `encoder.*` alias (actor_critic.py:55-56)
    assert sum(p.numel() for p in ac.parameters()) == 603037
    for k, v in ac.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k


@pytest.mark.parametrize("n", [4000, 333])
def test_forward_vs_oracle(n):
    ac, sd = make_ac()
    g = torch.Generator().manual_seed(n)
    obs, priv, hist = torch.randn(n, 42, generator=g), torch.rand(n, 18, generator=g) * 2 - 1, torch.randn(n, 630, generator=g)
    p = {k: v.float() for k, v in sd.items()}
    mean_ref, val_ref = ppo_oracle.actor_mean(p, obs, priv), ppo_oracle.critic_value(p, obs, priv)
    mean = ac.act_teacher(obs.to(DEV), priv.to(DEV)).cpu()
    val = ac.evaluate(obs.to(DEV), priv.to(DEV)).cpu()
    torch.testing.assert_close(mean, mean_ref, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(val, val_ref, rtol=2e-2, atol=2e-2)
    latent = ppo_oracle.mlp(p, "adaptation_module", (0, 2, 4), hist)
    stu_ref = ppo_oracle.mlp(p, "actor_body", (0, 2, 4, 6), torch.cat((obs, latent), -1))
    stu = ac.act_student(obs.to(DEV), hist.to(DEV)).cpu()
    torch.testing.assert_close(stu, stu_ref, rtol=2e-2, atol=2e-2)
    assert torch.nn.functional.cosine_similarity(mean.flatten(), mean_ref.flatten(), dim=0) > 0.999


def test_act_sampling_and_log_prob():
    ac, sd = make_ac()
    n = 2048
    obs, priv = torch.randn(n, 42, device=DEV), torch.rand(n, 18, device=DEV)
    z = torch.randn(n, 12, device=DEV)
    a = ac.act(obs, priv, inject_normal=z)
    mean = ac.action_mean
    torch.testing.assert_close(a, mean + ac.std.data * z, rtol=1e-6, atol=1e-6)
    ref_lp = ppo_oracle.normal_log_prob(a.cpu(), mean.cpu(), (mean * 0 + ac.std.data).cpu()).sum(-1)
    torch.testing.assert_close(ac.get_actions_log_prob(a).cpu(), ref_lp, rtol=1e-4, atol=1e-4)
    v = ac.evaluate(obs, priv)      # consumes the cached pass of act()
    assert v.shape == (n, 1)
    # Philox path: standard-normal statistics
    a2 = ac.act(obs, priv)
    zz = ((a2 - ac.action_mean) / ac.std.data).flatten()
    assert abs(zz.mean().item()) < 0.02 and abs(zz.std().item() - 1.0) < 0.02
    assert abs((zz ** 4).mean().item() - 3.0) < 0.2


def _load_storage(ppo, storage, n_envs, n_steps):
    st = ppo.storage
    shp = lambda t, d: t.reshape(n_steps, n_envs, d).to(DEV)
    st.observations.copy_(shp(storage["obs"], 42)); st.privileged_observations.copy_(shp(storage["priv"], 18))
    st.observation_histories.copy_(shp(storage["hist"], 630)); st.actions.copy_(shp(storage["actions"], 12))
    st.values.copy_(shp(storage["values"], 1)); st.returns.copy_(shp(storage["returns"], 1))
    st.actions_log_prob.copy_(shp(storage["old_logp"], 1)); st.advantages.copy_(shp(storage["advantages"], 1))
    st.mu.copy_(shp(storage["old_mu"], 12)); st.sigma.copy_(shp(storage["old_sigma"], 12))


def _grad_views(ac):
    out = {"std": ac.std_grad}
    names = {"env_factor_encoder": ac.env_factor_encoder, "adaptation_module": ac.adaptation_module,
             "actor_body": ac.actor_body, "critic_body": ac.critic_body}
    for pre, seq in names.items():
        for i, m in enumerate(seq):
            if isinstance(m, torch.nn.Linear):
                out["%s.%d.weight" % (pre, i)] = ac._grad_view[id(m.weight)]
                out["%s.%d.bias" % (pre, i)] = ac._grad_view[id(m.bias)]
    return out


def test_minibatch_gradients_vs_oracle(golden_dir):
    """One minibatch: loss statistics and every gradient tensor against fp32 autograd on the same rows."""
    from rapid_locomotion_rl_b200.ppo import PPO
    g = np.load(os.path.join(golden_dir, "learner.npz"))
    init, storage, perm = learner_case(g)
    ac, _ = make_ac()
    ppo = PPO(ac, device=DEV)
    ppo.init_storage(64, 8, [42], [18], [630], [12])
    _load_storage(ppo, storage, 64, 8)
    idx = perm[:256]
    o = ppo_oracle.PPOOracle(init)
    mb = {k: v[idx] for k, v in storage.items()}
    surr, vloss, aloss, kl = o.step(mb)
    ppo.debug_keep_grad = True
    offsets = {k: (v.data_ptr() - ac.flat_grad.data_ptr()) // 4 for k, v in _grad_views(ac).items()}
    ppo.minibatch_step(idx.to(DEV))
    torch.cuda.synchronize()
    stats = (ppo.debug_stats / 256).tolist()
    assert abs(stats[0] - surr) <= 3e-2 * abs(surr) + 1e-3, (stats[0], surr)
    assert abs(stats[1] - vloss) <= 3e-2 * abs(vloss) + 1e-3, (stats[1], vloss)
    assert abs(stats[2] - kl) <= 5e-2 * abs(kl) + 1e-3, (stats[2], kl)
    assert abs(stats[3] / 18 - aloss) <= 3e-2 * abs(aloss) + 1e-4, (stats[3] / 18, aloss)
    flat = ppo.debug_grad.cpu()
    for k in ppo_oracle.PARAM_ORDER:
        ref = o.last_adapt_grads[k] if k.startswith("adaptation_module") else o.last_grads[k]
        got = flat[offsets[k]:offsets[k] + ref.numel()].view(ref.shape)
        cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
        rel = ((got - ref).norm() / (ref.norm() + 1e-12)).item()
        assert cos >= 0.995 and rel <= 8e-2, "%s: cosine %.5f rel L2 %.4f" % (k, cos, rel)


def test_update_vs_golden(golden_dir):
    """Full PPO.update (5 epochs x 4 minibatches) against the reference's recorded losses / learning rate."""
    from rapid_locomotion_rl_b200.ppo import PPO
    g = np.load(os.path.join(golden_dir, "learner.npz"))
    init, storage, perm = learner_case(g)
    ac, _ = make_ac()
    ppo = PPO(ac, device=DEV)
    ppo.init_storage(64, 8, [42], [18], [630], [12])
    _load_storage(ppo, storage, 64, 8)
    ppo.storage.step = 8
    real = torch.randperm
    torch.randperm = lambda n, **kw: perm.to(DEV)
    try:
        res = ppo.update()
    finally:
        torch.randperm = real
    ref = g["ppo/result"]
    assert abs(res[0] - ref[0]) <= 5e-2 * abs(ref[0]) + 1e-3, (res, ref)
    assert abs(res[1] - ref[1]) <= 5e-2 * abs(ref[1]) + 2e-3, (res, ref)
    assert abs(res[2] - ref[2]) <= 5e-2 * abs(ref[2]) + 1e-4, (res, ref)
    assert abs(ppo.learning_rate - float(g["ppo/final_lr"])) <= 1e-9 + 0.34 * float(g["ppo/final_lr"]), ppo.learning_rate
    assert ppo.storage.step == 0
    # the weights moved, stayed finite, and the bf16 shadows follow the fp32 masters
    assert torch.isfinite(ac.flat).all()
    L = ac.L_act[0]
    torch.testing.assert_close(L.wb[:, :L.inp].float(), L.w.to(torch.bfloat16).float())
    torch.testing.assert_close(L.wbt[:, :L.out].float(), L.w.t().to(torch.bfloat16).float())


def test_gather_history_matches_indexing():
    """rl_ppo_gather_history == observation_histories[batch_idx] (rollout_storage.py:124) rounded to bf16, zero
    padded to the workspace pitch; odd dimensions take the scalar path."""
    import ctypes as C
    from rapid_locomotion_rl_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator().manual_seed(3)
    for rows, dim, ld, B in ((5000, 630, 632, 3001), (700, 45, 48, 129), (64, 7, 9, 64)):
        hist = torch.randn(rows, dim, generator=g).to(DEV)
        idx = torch.randint(0, rows, (B,), generator=g).to(DEV)
        out = torch.full((B, ld), 7.0, dtype=torch.bfloat16, device=DEV)
        _lib.check(lib.rl_ppo_gather_history(hist.data_ptr(), idx.data_ptr(), B, dim, out.data_ptr(), ld, _lib.current_stream()))
        torch.cuda.synchronize()
        assert torch.equal(out[:, :dim], hist[idx].to(torch.bfloat16))
        assert float(out[:, dim:].float().abs().max()) == 0.0
    assert lib.rl_ppo_gather_history(None, idx.data_ptr(), B, dim, out.data_ptr(), ld, _lib.current_stream()) != 0


def test_lagged_schedule_matches_serial_update(golden_dir, learner_path, monkeypatch):
    """PPO.update with the adaptation module one minibatch behind on the side branch (two captured graphs + flush,
    ONE gradient reduction point per minibatch) must produce the serial schedule's parameters and statistics: the
    schedules differ only in WHEN independent kernels run (fp32 atomics order aside)."""
    if learner_path != "chain":
        pytest.skip("the side branch exists on the chain path only")
    from rapid_locomotion_rl_b200.ppo import PPO
    g = np.load(os.path.join(golden_dir, "learner.npz"))
    init, storage, perm = learner_case(g)
    out = {}
    for name, env in (("serial", dict(RL_PPO_OVERLAP="0")), ("lagged", {})):
        monkeypatch.delenv("RL_PPO_OVERLAP", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ac, _ = make_ac()
        ppo = PPO(ac, device=DEV)
        ppo.init_storage(64, 8, [42], [18], [630], [12])
        _load_storage(ppo, storage, 64, 8)
        real = torch.randperm
        torch.randperm = lambda n, **kw: perm.to(DEV)
        try:
            for _ in range(2):               # second update re-uses the captured graphs
                ppo.storage.step = 8
                res = ppo.update()
        finally:
            torch.randperm = real
        torch.cuda.synchronize()
        assert ppo._graph_lag == (name == "lagged")
        out[name] = (ac.flat.clone(), res, ppo.learning_rate)
    d = (out["lagged"][0] - out["serial"][0]).abs().max().item()
    assert d <= 4e-4, d                                  # 40 Adam steps of |dw| <= lr = 1e-3 each
    np.testing.assert_allclose(out["lagged"][1], out["serial"][1], rtol=5e-3, atol=1e-6)
    assert abs(out["lagged"][2] - out["serial"][2]) <= 1e-12
