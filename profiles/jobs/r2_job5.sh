set -x
python -m pytest tests/test_env_gpu.py tests/test_plugins_gpu.py tests/test_runner_gpu.py -q -x 2>&1 | tail -15
python -m pytest tests/test_ppo_gpu.py -q -x 2>&1 | tail -15
for G in 1 2 3 4; do
RL_ROWS_STAGGER=$G python - <<'PY'
import sys, os, json, math, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
for envs, steps in ((32768, 500), (16384, 500)):
    n_rep = max(2, math.ceil(2.0 * bench.L2_BYTES / (envs * 1425)))
    reps = bench.build_replicas("mc_flat", envs, n_rep, "cuda:0")
    g = bench.time_env_steps(reps, steps, 5)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("STAGGER=%s envs %d: %.2f us/launch, frac %.3f" % (os.environ["RL_ROWS_STAGGER"], envs, best / steps * 1e3, envs * 1425 / (best / steps * 1e-3) / 1e9 / 6557.1))
    del reps, g
    torch.cuda.empty_cache()
PY
done
RL_ROWS_STAGGER=2 python profiles/trace_env_rows.py 2>&1 | head -16
python profiles/prof_rollout.py
python - <<'PY'
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
print(bench.runner_bench(4000, "cuda:0", iters=10))
PY
python bench.py --only-ppo | head -c 330
