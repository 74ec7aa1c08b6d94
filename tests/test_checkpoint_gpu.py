"""Checkpoint / export compatibility: the trained policy the reference ships loads into the product
ActorCritic and acts like the reference's ActorCritic (tests/golden/checkpoint.npz, produced by the reference class on a
seeded batch); the exported TorchScript modules are the ones the reference's Runner saves (mini_gym_learn/ppo/
__init__.py:227-242) and compose into the deployed policy (scripts/play.py)."""
import os

import numpy as np
import pytest
import torch

from cases import tensor_digest

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CKPT = "runs/rapid-locomotion/example/train/201852.132488/checkpoints/ac_weights_last.pt"


def _checkpoint_path():
    for base in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        p = os.path.join(base, CKPT)
        if os.path.isfile(p):
            return p
    pytest.skip("the reference checkpoint is not staged on this machine (oracle/make_ref.py)")


def _inputs(n=512):
    g = torch.Generator().manual_seed(5)
    return torch.randn(n, 42, generator=g), torch.rand(n, 18, generator=g) * 2 - 1, torch.randn(n, 630, generator=g) * 0.5


def test_shipped_checkpoint_loads_and_acts_like_the_reference(golden_dir):
    from rapid_locomotion_rl_b200.ppo import ActorCritic
    g = np.load(os.path.join(golden_dir, "checkpoint.npz"))
    sd = torch.load(_checkpoint_path(), map_location="cpu", weights_only=True)
    for k, v in sd.items():                      # the file is the one the golden was made from
        assert np.array_equal(tensor_digest(v.numpy()), g["digest/" + k]), k
    ac = ActorCritic(42, 18, 630, 12, device=DEV)
    missing = ac.load_state_dict(sd)             # strict: all 35 keys, including the duplicate `encoder.*` alias
    assert not missing.missing_keys and not missing.unexpected_keys
    for k, v in ac.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k
    obs, priv, hist = _inputs()
    assert np.allclose(np.concatenate([tensor_digest(x.numpy()) for x in (obs, priv, hist)]), g["input_digest"])
    ac.eval()
    o, p, h = obs.to(DEV), priv.to(DEV), hist.to(DEV)
    # Stated tolerance of the bf16 tensor-core path on the TRAINED weights: on these (out-of-distribution, unit-normal)
    # inputs the policy outputs reach |a| ~ 36 and rounding operands to bf16 alone moves single outputs by up to 0.12
    # (emulated on the CPU: bf16 inputs / weights / activations, fp32 accumulation), so the bound is on the whole output:
    # relative L2 error <= 1e-2, cosine > 0.9999, and no element off by more than 0.1 + 5e-2 |ref|.
    for name, got in (("act_teacher", ac.act_teacher(o, p)), ("act_student", ac.act_student(o, h)), ("evaluate", ac.evaluate(o, p)),
                      ("act_inference", ac.act_inference({"obs": o, "obs_history": h, "privileged_obs": p})),
                      ("act_expert_is_teacher", ac.act_expert({"obs": o, "privileged_obs": p}))):
        ref = torch.from_numpy(g["act_teacher" if name == "act_expert_is_teacher" else name])
        got = got.float().cpu()
        rel = ((got - ref).norm() / ref.norm()).item()
        cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
        # (the critic's outputs are ten times smaller than the actor's - mean |V| 0.44 - with the same absolute rounding
        # noise from its 512-wide first layer: emulated bf16 error 0.057 max, hence the wider relative bound)
        lim_rel, lim_cos = (5e-2, 0.999) if name == "evaluate" else (1e-2, 0.9999)
        assert rel <= lim_rel and cos > lim_cos, (name, rel, cos)
        torch.testing.assert_close(got, ref, rtol=5e-2, atol=1e-1, msg=name)


def test_export_matches_reference_runner_files(tmp_path, golden_dir):
    from rapid_locomotion_rl_b200.ppo import ActorCritic, export_policy
    g = np.load(os.path.join(golden_dir, "checkpoint.npz"))
    ac = ActorCritic(42, 18, 630, 12, device=DEV)
    ac.load_state_dict(torch.load(_checkpoint_path(), map_location="cpu", weights_only=True))
    files = export_policy(ac, str(tmp_path), iteration=400)
    names = sorted(os.path.basename(f) for f in files)
    assert names == ["ac_weights_000400.pt", "ac_weights_last.pt", "adaptation_module_latest.jit", "body_latest.jit"]
    sd = torch.load(os.path.join(tmp_path, "checkpoints", "ac_weights_last.pt"), map_location="cpu", weights_only=True)
    assert list(sd.keys()) == list(ac.state_dict().keys())
    adaptation = torch.jit.load(os.path.join(tmp_path, "checkpoints", "adaptation_module_latest.jit"))
    body = torch.jit.load(os.path.join(tmp_path, "checkpoints", "body_latest.jit"))
    obs, priv, hist = _inputs()
    with torch.no_grad():                        # scripts/play.py: action = body(cat(obs, adaptation_module(obs_history)))
        latent = adaptation(hist)
        action = body(torch.cat((obs, latent), dim=-1))
    torch.testing.assert_close(latent, torch.from_numpy(g["latent_student"]), rtol=1e-5, atol=1e-5)      # fp32 modules: exact math
    torch.testing.assert_close(action, torch.from_numpy(g["act_student"]), rtol=1e-4, atol=1e-5)
