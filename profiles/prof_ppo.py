"""Small driver for ncu: one PPO.update at BASELINE configs[0] shape (4000 envs x 24 steps)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rapid_locomotion_rl_b200.ppo import PPO, ActorCritic, PPO_Args  # noqa: E402

n, T = int(os.environ.get("PPO_ENVS", 4000)), 24
torch.manual_seed(0)
ac = ActorCritic(42, 18, 630, 12, device="cuda:0")
ppo = PPO(ac, device="cuda:0")
ppo.init_storage(n, T, [42], [18], [630], [12])
obs = torch.randn(T + 1, n, 42, device="cuda"); priv = torch.rand(T + 1, n, 18, device="cuda") * 2 - 1
hist = torch.randn(n, 630, device="cuda")
PPO_Args.num_learning_epochs = int(os.environ.get("PPO_EPOCHS", 1))
for it in range(2):
    for t in range(T):
        ppo.act(obs[t], priv[t], hist)
        ppo.process_env_step(torch.randn(n, device="cuda") * 0.05, torch.rand(n, device="cuda") < 0.01, {"env_bins": torch.zeros(n, device="cuda")})
    ppo.compute_returns(obs[T], priv[T])
    print(ppo.update())
torch.cuda.synchronize()
