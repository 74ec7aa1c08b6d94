timeout 900 python -m pytest tests/test_env_gpu.py tests/test_plugins_gpu.py tests/test_gymapi_gpu.py tests/test_runner_gpu.py -x -q 2>&1 | grep -v Warning | tail -4
python bench.py --steps 1000 --warmup 5 --quick 2>gpurun_out/r2_bench4q.err | tee gpurun_out/r2_bench4q.json | cut -c1-600
