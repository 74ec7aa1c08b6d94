set -x
# PPO loss fused into the forward chain's value-output epilogue (rl_chain_set_ppo_loss): parity, then A/B of the update
python -m pytest tests/test_ppo_gpu.py tests/test_chain_gpu.py tests/test_hlp_gpu.py tests/test_runner_gpu.py -x -q 2>&1 | grep -v Warning | tail -6
bash profiles/jobs/ppo_ab.sh RL_PPO_FUSED_LOSS=0 RL_PPO_FUSED_LOSS=1 RL_PPO_FUSED_LOSS=0 RL_PPO_FUSED_LOSS=1
