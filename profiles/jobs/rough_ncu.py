"""A few eager rough-terrain steps at 32768 envs (driver for the ncu capture of the height pre-pass and the 229-column step kernel)."""
import sys, math, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
envs = 32768
reps = bench.build_replicas("mc_rough_full", envs, 3, "cuda:0")
for _ in range(3):
    for env, actions, st in reps:
        env.step(actions)
torch.cuda.synchronize()
print("ok")
