"""Rough-terrain step (BASELINE configs[2]): time of the two launches (height pre-pass, step kernel) per step, CUDA events."""
import sys, math, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
for envs in (4000, 32768):
    case = "mc_rough_full"
    bpe = bench.BYTES_PER_ENV_STEP[case]
    n_rep = max(2, math.ceil(2.0 * bench.L2_BYTES / (envs * bpe)))
    reps = bench.build_replicas(case, envs, n_rep, "cuda:0")
    steps = 200
    g = bench.time_env_steps(reps, steps, 5)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("rough envs %d: %.2f us/step" % (envs, best / steps * 1e3))
    # the two launches separately (eager, same replica rotation)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        g.replay(); torch.cuda.synchronize()
    import collections
    tot = collections.Counter(); cnt = collections.Counter()
    for e in prof.events():
        if e.device_type.name == "CUDA" and ("kernel" in e.name or "rl::" in e.name):
            tot[e.name[:60]] += e.device_time; cnt[e.name[:60]] += 1
    for k, v in tot.most_common(4):
        print("   %-60s %6.2f us x %d" % (k, v / cnt[k], cnt[k]))
    del reps, g
    torch.cuda.empty_cache()
