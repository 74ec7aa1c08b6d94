set -x
python -m pytest tests/test_env_gpu.py -q 2>&1 | tail -8
python -m pytest tests/test_ppo_gpu.py -q -x 2>&1 | tail -5
for pdl in 1 0; do
RL_ENV_PDL=$pdl python - <<'PY'
import sys, os, json, math, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
for envs, steps in ((4000, 1000), (32768, 500), (262144, 100)):
    n_rep = max(2, math.ceil(2.0 * bench.L2_BYTES / (envs * 1425)))
    reps = bench.build_replicas("mc_flat", envs, n_rep, "cuda:0")
    g = bench.time_env_steps(reps, steps, 5)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("PDL=%s envs %d: %.2f us/launch, %.3e env-steps/s, frac %.3f" % (os.environ["RL_ENV_PDL"], envs, best / steps * 1e3, envs * steps / best * 1e3, envs * 1425 / (best / steps * 1e-3) / 1e9 / 6557.1))
    del reps, g
    torch.cuda.empty_cache()
PY
done
python bench.py --only-ppo > gpurun_out/r2_ppo2.json 2> gpurun_out/r2_ppo2.err; cat gpurun_out/r2_ppo2.json | head -c 400
RL_PPO_OVERLAP=0 python bench.py --only-ppo | head -c 300
