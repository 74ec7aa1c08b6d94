set -x
# host topology of the 8-GPU box, then the e2e section with / without binding every rank to its GPU's NUMA node
nvidia-smi topo -m 2>&1 | head -30
lscpu | grep -i -E "numa|socket|^CPU\(s\)|model name"
nproc
for v in 0 1; do
  RL_BENCH_NUMA=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 300 --warmup 5 --quick > gpurun_out/r2_bench_n8_numa$v.json 2> gpurun_out/r2_bench_n8_numa$v.err
  python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_n8_numa$v.json") if l.startswith("{")][-1])
e=d["e2e"]; print("NUMA=$v", "%.4g"%d["value"], {k:("%.4g"%x if isinstance(x,float) else x) for k,x in e.items() if k!="note"})
P
done
