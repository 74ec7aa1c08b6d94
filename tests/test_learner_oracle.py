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
