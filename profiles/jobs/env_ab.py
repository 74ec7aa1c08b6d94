"""A/B helper: us per launch of the fused env step inside the K-step graph, several sizes (CUDA events, best of 3)."""
import sys, os, math, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for case, envs, steps in (("mc_flat", 4000, 1000), ("mc_flat", 16384, 500), ("mc_flat", 32768, 500), ("go1", 32768, 500), ("mc_flat", 262144, 100)):
    bpe = bench.BYTES_PER_ENV_STEP[case]
    n_rep = max(2, math.ceil(2.0 * bench.L2_BYTES / (envs * bpe)))
    reps = bench.build_replicas(case, envs, n_rep, "cuda:0")
    g = bench.time_env_steps(reps, steps, 5)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%s %s envs %d: %.2f us/launch, %.3e env-steps/s, frac %.3f" % (tag, case, envs, best / steps * 1e3, envs * steps / best * 1e3, envs * bpe / (best / steps * 1e-3) / 1e9 / 6557.1))
    del reps, g
    torch.cuda.empty_cache()
