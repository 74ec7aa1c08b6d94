"""Per-CTA timeline of env_step_rows_kernel (rl_debug_env_rows_trace): when each CTA entered, had its tile, finished
phase 1 / phase 2, issued its stores and left - per SM residency slot - for one launch inside a K-step graph."""
import ctypes as C
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from rapid_locomotion_rl_b200 import _lib  # noqa: E402

envs = int(os.environ.get("ENVS", 32768))
lib = _lib.lib()
n_cta = envs // 32
n_rep = max(2, math.ceil(2.0 * bench.L2_BYTES / (envs * 1425)))
reps = bench.build_replicas("mc_flat", envs, n_rep, "cuda:0")
_lib.check(lib.rl_debug_env_rows_trace(n_cta, None, 0))
g = bench.time_env_steps(reps, 40, 5)      # the LAST launch of the graph is the one left in the buffer
g.replay(); torch.cuda.synchronize()
buf = (C.c_uint64 * (n_cta * 8))()
_lib.check(lib.rl_debug_env_rows_trace(0, buf, n_cta))
t = np.frombuffer(buf, dtype=np.uint64).reshape(n_cta, 8).astype(np.int64)
smid = t[:, 7]
t0 = t[:, 0].min()
rel = (t[:, :7] - t0) / 1e3          # us since the first CTA entered
names = ["entry", "loads issued", "tile landed", "phase1 done", "phase2 done", "stores issued", "stores read"]
print("envs %d, %d CTAs on %d SMs; kernel span (first entry -> last exit) %.2f us" % (envs, n_cta, len(set(smid.tolist())), rel[:, 6].max()))
for i, n in enumerate(names):
    c = rel[:, i]
    print("  %-14s min %6.2f  p10 %6.2f  median %6.2f  p90 %6.2f  max %6.2f" % (n, c.min(), np.percentile(c, 10), np.median(c), np.percentile(c, 90), c.max()))
d = np.diff(rel, axis=1)
for i in range(6):
    print("  %-14s -> %-14s median %6.2f  p90 %6.2f us" % (names[i], names[i + 1], np.median(d[:, i]), np.percentile(d[:, i], 90)))
# residency order on one SM
sm = int(np.bincount(smid.astype(int)).argmax())
rows = np.nonzero(smid == sm)[0]
rows = rows[np.argsort(rel[rows, 0])]
print("SM %d:" % sm)
for r in rows:
    print("   cta %5d " % r + "  ".join("%6.2f" % v for v in rel[r]))
