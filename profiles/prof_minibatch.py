"""ncu driver: ONE PPO minibatch step (all its kernels once) at B = PPO_B rows."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rapid_locomotion_rl_b200.ppo import PPO, ActorCritic  # noqa: E402

B = int(os.environ.get("PPO_B", 196608))
T = 24
n = B * 4 // T
torch.manual_seed(0)
ac = ActorCritic(42, 18, 630, 12, device="cuda:0")
ppo = PPO(ac, device="cuda:0")
ppo.init_storage(n, T, [42], [18], [630], [12])
st = ppo.storage
st.observations.normal_(); st.privileged_observations.uniform_(-1, 1); st.observation_histories.normal_()
st.actions.normal_(); st.values.normal_(); st.returns.normal_(); st.advantages.normal_()
st.actions_log_prob.fill_(-17.0); st.mu.normal_(); st.sigma.fill_(1.0)
idx = torch.randperm(n * T, device="cuda")[:B]
for it in range(int(os.environ.get("PPO_ITERS", 2))):
    ppo.minibatch_step(idx)
torch.cuda.synchronize()
print("ok", ppo._stats.tolist())
