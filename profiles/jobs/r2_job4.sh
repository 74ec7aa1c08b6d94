set -x
python profiles/trace_env_rows.py > gpurun_out/r2_env_rows_trace_32k.txt 2>&1; head -60 gpurun_out/r2_env_rows_trace_32k.txt
ENVS=4000 python profiles/trace_env_rows.py > gpurun_out/r2_env_rows_trace_4k.txt 2>&1; head -30 gpurun_out/r2_env_rows_trace_4k.txt
python - <<'PY'
# floor of a kernel node inside a CUDA graph: K dependent launches of an (almost) empty kernel
import torch, time
x = torch.zeros(1024, device="cuda")
g = torch.cuda.CUDAGraph()
torch.cuda.synchronize()
with torch.cuda.graph(g):
    for _ in range(1000):
        x.add_(1.0)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print("empty-ish kernel node in a graph: %.2f us per node" % (e0.elapsed_time(e1)))
PY
