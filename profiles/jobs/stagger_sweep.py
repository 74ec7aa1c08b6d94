"""Sweep of the time-stagger parameters of the fused env step (RL_ENV_STAGGER is read once per process: one subprocess each)."""
import os, subprocess, sys
envs = sys.argv[1] if len(sys.argv) > 1 else "32768"
code = r'''
import sys, math, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
envs = int(sys.argv[1]); steps = 500
import os
case = os.environ.get("CASE", "mc_flat")
reps = bench.build_replicas(case, envs, max(2, math.ceil(2.0 * bench.L2_BYTES / (envs * bench.BYTES_PER_ENV_STEP[case]))), "cuda:0")
g = bench.time_env_steps(reps, steps, 5)
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("%.2f" % (best / steps * 1e3))
'''
for cfg in sys.argv[2:]:
    env = dict(os.environ)
    if cfg != "default":
        env["RL_ENV_STAGGER"] = cfg
    r = subprocess.run([sys.executable, "-c", code, envs], env=env, capture_output=True, text=True)
    print("%s envs %s stagger %-16s %s us/launch" % (os.environ.get("CASE", "mc_flat"), envs, cfg, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]))
