set -x
python -m pytest tests/test_multigpu.py -x -q 2>&1 | grep -v Warning | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 500 --warmup 5 > gpurun_out/r2_bench_n2_final3.json 2> gpurun_out/r2_bench_n2_final3.err
tail -c 200 gpurun_out/r2_bench_n2_final3.err
