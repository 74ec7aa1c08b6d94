set -x
RL_PPO_FUSED_LOSS=1 python profiles/prof_timeline.py > gpurun_out/r2_ppo_timeline_fused.txt 2>&1
RL_PPO_FUSED_LOSS=0 python profiles/prof_timeline.py > gpurun_out/r2_ppo_timeline_unfused.txt 2>&1
