set -x
python -m pytest tests/test_env_gpu.py -q -x -k "rows_kernel" 2>&1 | tail -15
python -m pytest tests/test_env_gpu.py -q 2>&1 | tail -8
python -m pytest tests/test_ppo_gpu.py tests/test_chain_gpu.py tests/test_runner_gpu.py -q -x 2>&1 | tail -15
python bench.py --steps 500 --warmup 5 --quick > gpurun_out/r2_bench1_quick.json 2> gpurun_out/r2_bench1_quick.err; tail -c 900 gpurun_out/r2_bench1_quick.json
python bench.py --only-ppo > gpurun_out/r2_ppo1.json 2> gpurun_out/r2_ppo1.err; cat gpurun_out/r2_ppo1.json | head -c 700
ncu --set full --clock-control none --import-source on -k regex:env_step_rows -s 70 -c 1 -o gpurun_out/r2_env_rows_32k -f python bench.py --steps 60 --warmup 3 --quick > gpurun_out/r2_ncu_env.log 2>&1
ls -la gpurun_out/
