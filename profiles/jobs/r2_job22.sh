python -m pytest tests/test_chain_gpu.py tests/test_ppo_gpu.py tests/test_runner_gpu.py tests/test_checkpoint_gpu.py -x -q 2>&1 | grep -v Warning | tail -6
python bench.py --only-ppo 2>gpurun_out/r2_ppo3.err | tee gpurun_out/r2_ppo3.json | cut -c1-300
python bench.py --only-ppo --ppo-envs 32768 2>>gpurun_out/r2_ppo3.err | tee -a gpurun_out/r2_ppo3.json | cut -c1-300
