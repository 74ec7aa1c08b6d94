python -m pytest tests/test_env_gpu.py tests/test_plugins_gpu.py tests/test_gymapi_gpu.py -x -q 2>&1 | grep -v Warning | tail -6
python - <<'PY'
import sys, os, json, math, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
for pre in ("1", "0"):
    os.environ["RL_ENV_HEIGHTS_PREPASS"] = pre
    for case, envs, steps in (("mc_rough_full", 4000, 300), ("mc_rough_full", 32768, 200)):
        bpe = bench.BYTES_PER_ENV_STEP[case]
        n_rep = max(2, math.ceil(2.0 * bench.L2_BYTES / (envs * bpe)))
        reps = bench.build_replicas(case, envs, n_rep, "cuda:0")
        g = bench.time_env_steps(reps, steps, 5)
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("prepass=%s %s envs %d: %.2f us/step, %.3e env-steps/s, frac %.3f" % (pre, case, envs, best / steps * 1e3, envs * steps / best * 1e3, envs * bpe / (best / steps * 1e-3) / 1e9 / 6557.1))
        del reps, g
        torch.cuda.empty_cache()
PY
