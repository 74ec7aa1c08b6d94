"""ncu driver: the teacher-forward chain alone (B rows, save=True), three launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rapid_locomotion_rl_b200.ppo import ActorCritic  # noqa: E402

B = int(os.environ.get("B", 24000))
torch.manual_seed(0)
ac = ActorCritic(42, 18, 630, 12, device="cuda:0")
w = ac.workspace(B, backward=True)
for k in ("Xp", "Xac"):
    w[k].copy_(torch.randn(w[k].shape, device="cuda") * 0.5)
for _ in range(3):
    ac.forward_teacher(B, save=True)
torch.cuda.synchronize()
print("ok")
