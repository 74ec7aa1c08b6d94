"""CUDA-event time of the two grouped wgrad launches of a PPO minibatch (main: 10 layers, adaptation: 3)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rapid_locomotion_rl_b200 import _lib  # noqa: E402

lib = _lib.lib()
_lib.check(lib.rl_gemm_init())
for K in (24000, 196608):
    groups = {"main": [(256, 512), (256, 512), (128, 256), (128, 256), (128, 256), (12, 128), (1, 128), (18, 128), (1024, 60), (256, 18)],
              "adapt": [(256, 630), (32, 256), (18, 32)]}
    for name, shapes in groups.items():
        probs, keep = [], []
        for M, N in shapes:
            ldy, ldx = (M + 7) // 8 * 8, (N + 7) // 8 * 8
            dY = torch.randn(K, ldy, device="cuda").to(torch.bfloat16); X = torch.randn(K, ldx, device="cuda").to(torch.bfloat16)
            dW = torch.zeros(M, N, device="cuda"); db = torch.zeros(M, device="cuda")
            q = _lib.RlWgradProblem()
            q.dY, q.X, q.dW, q.db = dY.data_ptr(), X.data_ptr(), dW.data_ptr(), db.data_ptr()
            q.M, q.N, q.K, q.ld_dy, q.ld_x, q.ld_dw, q.split_k = M, N, K, ldy, ldx, N, 0
            probs.append(q); keep.append((dY, X, dW, db))
        arr = (_lib.RlWgradProblem * len(probs))(*probs)
        st = _lib.current_stream()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        ts = []
        for it in range(6):
            flush.zero_()                      # operands not L2-hot
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); _lib.check(lib.rl_wgrad_grouped(arr, len(probs), st)); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        fl = 2.0 * K * sum(m * n for m, n in shapes)
        t = sorted(ts)[len(ts) // 2]
        print("K=%d %s: %.1f us  %.0f TFLOP/s  (RL_WGRAD_KB=%s)" % (K, name, t, fl / t / 1e6, os.environ.get("RL_WGRAD_KB", "auto")))
