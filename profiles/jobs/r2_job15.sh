set -x
python -m pytest tests/test_env_gpu.py tests/test_plugins_gpu.py -q 2>&1 | grep -v Warning | tail -6
python - <<'PY'
import sys, os, json, math, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
for case, envs, steps in (("mc_rough_full", 4000, 300), ("mc_rough_full", 32768, 200), ("go1", 32768, 500), ("go1", 4000, 500), ("mc_flat", 4000, 1000), ("mc_flat", 32768, 500)):
    bpe = bench.BYTES_PER_ENV_STEP[case]
    n_rep = max(2, math.ceil(2.0 * bench.L2_BYTES / (envs * bpe)))
    reps = bench.build_replicas(case, envs, n_rep, "cuda:0")
    g = bench.time_env_steps(reps, steps, 5)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%s envs %d: %.2f us/launch, %.3e env-steps/s, frac %.3f" % (case, envs, best / steps * 1e3, envs * steps / best * 1e3, envs * bpe / (best / steps * 1e-3) / 1e9 / 6557.1))
    del reps, g
    torch.cuda.empty_cache()
PY
