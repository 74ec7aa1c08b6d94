"""ncu driver: Runner iterations at PPO_ENVS envs with the 24-step rollout replayed from CUDA graphs.
`ncu --metrics gpu__time_duration.sum --clock-control none --csv ...` of this script lists every kernel of the
rollout graphs (policy chain, Normal sample, fused env step, history push, storage add) plus GAE + PPO.update."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import build_case  # noqa: E402
from rapid_locomotion_rl_b200.envs import HistoryWrapper, VelocityTrackingEasyEnv  # noqa: E402
from rapid_locomotion_rl_b200.ppo import Runner  # noqa: E402

n = int(os.environ.get("PPO_ENVS", 4000))
iters = int(os.environ.get("PPO_ITERS", 8))
cfg, robot, terrain = build_case("mc_flat", n)
env = HistoryWrapper(VelocityTrackingEasyEnv(sim_device="cuda:0", headless=True, cfg=cfg, terrain=terrain, seed=0))
torch.manual_seed(0)
r = Runner(env, device="cuda:0", graph_rollout=True)
hist = r.learn(iters)
torch.cuda.synchronize()
print("ok", hist[-1]["time_iter"] * 1e3, "ms per iteration;", hist[-1]["mean_value_loss"])
