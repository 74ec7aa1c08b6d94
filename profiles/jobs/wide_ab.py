"""Four-warp vs 16-warp (wide) env-step kernel at small grids: us per launch inside the K-step graph (rl_debug_env_rows 3 / 4)."""
import sys, math, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from rapid_locomotion_rl_b200 import _lib
lib = _lib.lib()
for case, envs in (("mc_flat", 2048), ("mc_flat", 4000), ("go1", 4000), ("mc_flat", 8192), ("mc_flat", 16384), ("mc_flat", 32768)):
    if envs % 32:
        envs = envs // 32 * 32
    for mode, name in ((3, "four warps"), (4, "wide")):
        lib.rl_debug_env_rows(mode)
        bpe = bench.BYTES_PER_ENV_STEP[case]
        reps = bench.build_replicas(case, envs, max(2, math.ceil(2.0 * bench.L2_BYTES / (envs * bpe))), "cuda:0")
        steps = 500
        g = bench.time_env_steps(reps, steps, 5)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("%s envs %d %-10s: %.2f us/launch, %.3e env-steps/s" % (case, envs, name, best / steps * 1e3, envs * steps / best * 1e3))
        del reps, g
        torch.cuda.empty_cache()
lib.rl_debug_env_rows(-1)
