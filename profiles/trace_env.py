"""Phase timeline (globaltimer, ns) of one CTA of the fused env-step kernel + CUDA-event kernel time."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from rapid_locomotion_rl_b200 import _lib  # noqa: E402

for envs in (4000, 32768, 262144):
    reps = bench.build_replicas("mc_flat", envs, 3, "cuda:0")
    lib = _lib.lib()
    for i in range(6):
        reps[i % 3][0].step(reps[i % 3][1])
    torch.cuda.synchronize()
    lib.rl_debug_env_trace(1, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps[0][0].step(reps[0][1])
    e1.record()
    torch.cuda.synchronize()
    buf = (C.c_uint64 * 16)()
    lib.rl_debug_env_trace(0, buf)
    t = list(buf)[:9]
    names = ["entry", "SoA loads issued", "staged rows landed", "phase1 start", "phase1 end", "barrier", "outputs written",
             "barrier", "stores issued"]
    print("envs=%d  kernel (events, eager launch) %.1f us; middle CTA timeline (ns since its entry):" % (envs, e0.elapsed_time(e1) * 1e3))
    print("   " + "  ".join("%s +%d" % (n, x - t[0]) for n, x in zip(names, t)))
    del reps
    torch.cuda.empty_cache()
