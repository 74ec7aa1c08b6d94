python - <<'PY'
import sys, os, json, math, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from rapid_locomotion_rl_b200 import _lib
lib = _lib.lib()
lib.rl_debug_env_rows(3)
case, envs, steps = "mc_flat", 32768, 300
bpe = bench.BYTES_PER_ENV_STEP[case]
n_rep = max(2, math.ceil(2.0 * bench.L2_BYTES / (envs * bpe)))
reps = bench.build_replicas(case, envs, n_rep, "cuda:0")
for cfg in ("0,0,0", "1000,592,0", "300,148,148", "500,148,148", "800,148,148", "500,296,296", "1000,296,296", "1500,296,296", "200,148,148", "100,148,148", "400,74,74", "1000,148,296", "700,296,148", "1000,444,148"):
    os.environ["RL_ENV_STAGGER"] = cfg
    g = bench.time_env_steps(reps, steps, 5)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("stagger=%s: %.2f us/step frac %.3f" % (cfg, best / steps * 1e3, envs * bpe / (best / steps * 1e-3) / 1e9 / 6557.1))
    del g
PY
