"""Pins oracle/ppo_oracle.py against tests/golden/learner.npz (reference outputs)."""
import os

import numpy as np
import torch

from oracle import ppo_oracle


def test_gae_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "learner.npz"))
    torch.set_num_threads(1)
    ret, adv = ppo_oracle.compute_returns(torch.from_numpy(g["gae/rewards"]), torch.from_numpy(g["gae/values"]),
                                          torch.from_numpy(g["gae/dones"]), torch.from_numpy(g["gae/last_values"]),
                                          0.99, 0.95)
    assert np.array_equal(ret.numpy(), g["gae/returns"])
    assert np.array_equal(adv.numpy(), g["gae/advantages"])


def learner_case(g):
    """(init state dict, flattened storage dict, permutation) of the golden PPO case as torch tensors."""
    from cases import learner_rollout_inputs, learner_weights
    init = {k: torch.from_numpy(v) for k, v in learner_weights().items()}
    steps, last, perm = learner_rollout_inputs(64, 8)
    storage = {k[len("ppo/storage/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("ppo/storage/")}
    for name in ("obs", "priv", "hist"):
        storage[name] = torch.from_numpy(np.concatenate([s[name] for s in steps], 0))
    return init, storage, torch.from_numpy(perm)


def test_ppo_update_bit_exact(golden_dir):
    """The oracle's PPO.update reproduces the reference's final weights, losses and learning rate."""
    from cases import tensor_digest
    g = np.load(os.path.join(golden_dir, "learner.npz"))
    torch.set_num_threads(1)
    init, storage, perm = learner_case(g)
    o = ppo_oracle.PPOOracle(init)
    res = o.update(storage, perm)
    assert np.array_equal(np.array(res), g["ppo/result"])
    assert o.lr == float(g["ppo/final_lr"])
    for k in ppo_oracle.PARAM_ORDER:
        assert np.array_equal(tensor_digest(o.p[k].detach().numpy()), g["ppo/final_digest/" + k]), k
