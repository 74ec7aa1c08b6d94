timeout 600 python -m pytest tests/test_env_gpu.py -x -q -k "rows_kernel" 2>&1 | grep -v Warning | tail -6
python - <<'PY'
import sys, os, json, math, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from rapid_locomotion_rl_b200 import _lib
lib = _lib.lib()
for mode in (3, 2):
    lib.rl_debug_env_rows(mode)
    for case, envs, steps in (("mc_flat", 32768, 500), ("mc_flat", 262144, 100), ("go1", 32768, 300), ("mc_flat", 4000, 500), ("mc_flat", 16384, 500)):
        bpe = bench.BYTES_PER_ENV_STEP[case]
        n_rep = max(2, math.ceil(2.0 * bench.L2_BYTES / (envs * bpe)))
        reps = bench.build_replicas(case, envs, n_rep, "cuda:0")
        g = bench.time_env_steps(reps, steps, 5)
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("mode=%d %s envs %d: %.2f us/step, %.3e env-steps/s, frac %.3f" % (mode, case, envs, best / steps * 1e3, envs * steps / best * 1e3, envs * bpe / (best / steps * 1e-3) / 1e9 / 6557.1))
        del reps, g
        torch.cuda.empty_cache()
PY
