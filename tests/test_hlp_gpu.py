"""The high_level_policy learner family (reference: high_level_policy/ppo - tanh networks, USE_LATENT = False, no command
bins, 200 steps per env) on the fused kernels, against goldens produced by the UNMODIFIED reference classes
(tests/golden/make_golden.py hlp -> hlp.npz).

Stated tolerance (bf16 operands, fp32 accumulation, tanh.approx): network outputs 2e-2 abs + 2e-2 rel; loss scalars 5 %;
the KL-adaptive learning rate must take the reference's schedule step for step; weight updates as in test_ppo_gpu.
"""
import os

import numpy as np
import pytest
import torch

from cases import hlp_weights, learner_rollout_inputs, tensor_digest

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_hlp_ac():
    from rapid_locomotion_rl_b200.high_level_policy.ppo import ActorCritic
    ac = ActorCritic(42, 18, 630, 12, device=DEV)
    ac.load_state_dict({k: torch.from_numpy(v) for k, v in hlp_weights().items()})
    return ac


def test_hlp_state_dict_is_the_reference_layout(golden_dir):
    g = np.load(os.path.join(golden_dir, "hlp.npz"))
    ac = make_hlp_ac()
    sd = ac.state_dict()
    assert sorted(sd.keys()) == list(g["keys"])
    want = hlp_weights()
    for k, v in sd.items():
        assert tuple(v.shape) == want[k].shape, k
        assert np.array_equal(v.cpu().numpy(), want[k]), k
    # the latent path is pinned to zero: encoder output layer and the latent columns of both first layers
    assert float(ac.env_factor_encoder[-1].weight.abs().max()) == 0.0 and float(ac.env_factor_encoder[-1].bias.abs().max()) == 0.0
    assert float(ac.actor_body[0].weight[:, 42:].abs().max()) == 0.0 and float(ac.critic_body[0].weight[:, 42:].abs().max()) == 0.0


def test_hlp_forward_vs_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "hlp.npz"))
    ac = make_hlp_ac()
    steps, last, perm = learner_rollout_inputs(64, 8)
    T = lambda a: torch.from_numpy(a).to(DEV)
    obs, priv, hist = T(steps[0]["obs"]), T(steps[0]["priv"]), T(steps[0]["hist"])
    mean = ac.act_teacher(obs, priv).cpu().numpy()
    student = ac.act_student(obs, hist).cpu().numpy()
    value = ac.evaluate(obs, priv).cpu().numpy()
    for got, key in ((mean, "fwd/mean"), (student, "fwd/student"), (value, "fwd/value")):
        np.testing.assert_allclose(got, g[key], rtol=2e-2, atol=2e-2, err_msg=key)
    # (without a latent the student and the teacher are the same function of the observations)
    np.testing.assert_array_equal(mean, student)


def test_hlp_update_vs_golden(golden_dir):
    from rapid_locomotion_rl_b200.high_level_policy.ppo import PPO
    g = np.load(os.path.join(golden_dir, "hlp.npz"))
    ac = make_hlp_ac()
    ppo = PPO(ac, device=DEV)
    ppo.init_storage(64, 8, [42], [18], [630], [12])
    steps, last, perm = learner_rollout_inputs(64, 8)
    st = ppo.storage
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    st.observations.copy_(T(np.stack([s["obs"] for s in steps])))
    st.privileged_observations.copy_(T(np.stack([s["priv"] for s in steps])))
    st.observation_histories.copy_(T(np.stack([s["hist"] for s in steps])))
    shp = lambda k, d: T(g["ppo/storage/" + k]).reshape(8, 64, d)
    st.actions.copy_(shp("actions", 12)); st.values.copy_(shp("values", 1)); st.returns.copy_(shp("returns", 1))
    st.actions_log_prob.copy_(shp("old_logp", 1)); st.advantages.copy_(shp("advantages", 1))
    st.mu.copy_(shp("old_mu", 12)); st.sigma.copy_(shp("old_sigma", 12))
    st.step = 8
    real = torch.randperm
    torch.randperm = lambda n, **kw: torch.from_numpy(perm).to(DEV)
    try:
        res = ppo.update()
    finally:
        torch.randperm = real
    ref = g["ppo/result"]
    assert abs(res[0] - ref[0]) <= 5e-2 * abs(ref[0]) + 1e-3, (res, ref)
    assert abs(res[1] - ref[1]) <= 5e-2 * abs(ref[1]) + 2e-3, (res, ref)
    assert res[2] == 0 and ref[2] == 0                      # no adaptation module to train
    ref_lr = float(g["ppo/final_lr"])
    assert abs(ppo.learning_rate - ref_lr) <= 1e-6 * ref_lr, (ppo.learning_rate, ref_lr)
    init_w = hlp_weights()
    dw, dw_ref = [], []
    for k, v in ac.state_dict().items():
        ref_d = g["ppo/final_digest/" + k]
        got = tensor_digest(v.detach().cpu().numpy())
        start = tensor_digest(init_w[k])
        dw.append(got[2:] - start[2:]); dw_ref.append(ref_d[2:] - start[2:])
    dw, dw_ref = np.concatenate(dw), np.concatenate(dw_ref)
    cos = float(dw @ dw_ref / (np.linalg.norm(dw) * np.linalg.norm(dw_ref) + 1e-30))
    print("hlp final-weight update vs reference: cosine %.4f, max |diff| %.2e, |dw_ref| max %.2e" %
          (cos, np.abs(dw - dw_ref).max(), np.abs(dw_ref).max()))
    assert cos >= 0.995 and np.median(np.abs(dw - dw_ref)) <= 1e-4, (cos, np.median(np.abs(dw - dw_ref)))
    # the pinned latent path did not move
    assert float(ac.env_factor_encoder[-1].weight.abs().max()) == 0.0
    assert float(ac.actor_body[0].weight[:, 42:].abs().max()) == 0.0 and float(ac.critic_body[0].weight[:, 42:].abs().max()) == 0.0
    assert torch.isfinite(ac.flat).all()


def test_hlp_runner_and_export(tmp_path):
    """Runner of the family: 200 steps per env, bodies-only checkpoint, TorchScript body with a [., num_obs] first layer."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cases import build_case
    from rapid_locomotion_rl_b200.envs import VelocityTrackingEasyEnv, HistoryWrapper
    from rapid_locomotion_rl_b200.high_level_policy.ppo import Runner, RunnerArgs
    cfg, robot, terrain = build_case("mc_flat", 64)
    env = HistoryWrapper(VelocityTrackingEasyEnv(sim_device=DEV, headless=True, cfg=cfg, terrain=terrain))
    runner = Runner(env, device=DEV)
    assert runner.num_steps_per_env == RunnerArgs.num_steps_per_env == 200
    hist = runner.learn(1, save_dir=str(tmp_path))
    assert np.isfinite(hist[0]["mean_value_loss"]) and hist[0]["adaptation_loss"] == 0
    ck = os.path.join(str(tmp_path), "checkpoints")
    assert sorted(os.listdir(ck)) == ["ac_weights_000000.pt", "ac_weights_last.pt", "body_latest.jit"]
    sd = torch.load(os.path.join(ck, "ac_weights_last.pt"), weights_only=True)
    assert sd["actor_body.0.weight"].shape == (512, 42) and not any(k.startswith(("encoder", "adaptation")) for k in sd)
    body = torch.jit.load(os.path.join(ck, "body_latest.jit"))
    obs = torch.randn(5, 42)
    want = runner.alg.actor_critic.act_inference({"obs": obs.to(DEV), "obs_history": torch.zeros(5, 630, device=DEV)}).cpu()
    torch.testing.assert_close(body(obs), want, rtol=2e-2, atol=2e-2)
