set -x
python -m pytest tests/test_multigpu.py -x -q 2>&1 | grep -v Warning | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 500 --warmup 5 > gpurun_out/r2_bench_n2_final2.json 2> gpurun_out/r2_bench_n2_final2.err
tail -c 300 gpurun_out/r2_bench_n2_final2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 30 --warmup 3 | tail -c 400
