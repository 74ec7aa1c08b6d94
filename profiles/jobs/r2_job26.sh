RL_PPO_FUSED_ADAM=0 python bench.py --only-ppo --ppo-envs 32768 2>>gpurun_out/r2_ppo4.err | cut -c1-200
python bench.py --only-ppo --ppo-envs 32768 2>>gpurun_out/r2_ppo4.err | cut -c1-200
