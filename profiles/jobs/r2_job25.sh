python -m pytest tests/test_chain_gpu.py tests/test_ppo_gpu.py tests/test_runner_gpu.py tests/test_checkpoint_gpu.py -x -q 2>&1 | grep -v Warning | tail -4
python bench.py --only-ppo 2>gpurun_out/r2_ppo4.err | tee gpurun_out/r2_ppo4.json | cut -c1-200
RL_PPO_FUSED_ADAM=0 python bench.py --only-ppo 2>>gpurun_out/r2_ppo4.err | cut -c1-200
python bench.py --only-ppo --ppo-envs 32768 2>>gpurun_out/r2_ppo4.err | tee -a gpurun_out/r2_ppo4.json | cut -c1-200
