"""Durations of the two launches of a rough-terrain step (height pre-pass, step kernel): CUDA events around each, 32768 and 4000 envs."""
import sys, math, os, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
for envs in (32768, 4000):
    reps = bench.build_replicas("mc_rough_full", envs, max(2, math.ceil(2.0 * bench.L2_BYTES / (envs * 2921))), "cuda:0")
    for env, actions, st in reps:
        for _ in range(2):
            env.step(actions)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            for env, actions, st in reps:
                env.step(actions)
        torch.cuda.synchronize()
    import collections
    t, n = collections.Counter(), collections.Counter()
    for e in prof.events():
        if e.device_type.name == "CUDA" or "cuda" in str(e.device_type).lower():
            t[e.name[:70]] += e.device_time if hasattr(e, "device_time") else e.cuda_time
            n[e.name[:70]] += 1
    print("envs", envs)
    for k, v in t.most_common(6):
        print("   %-72s n=%4d avg %.2f us" % (k, n[k], v / n[k]))
