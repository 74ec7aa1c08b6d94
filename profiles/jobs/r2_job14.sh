set -x
python -m pytest tests/test_env_gpu.py -q 2>&1 | grep -v Warning | tail -6
python bench.py --steps 1000 --warmup 5 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; tail -c 3000 gpurun_out/r2_bench3.json; tail -5 gpurun_out/r2_bench3.err
python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/r2_bench3_ref.json 2> gpurun_out/r2_bench3_ref.err; tail -c 700 gpurun_out/r2_bench3_ref.json
