set -x
RL_PPO_ADA_FWD_EARLY=1 python profiles/prof_timeline.py > gpurun_out/r2_ppo_timeline_adafwd1.txt 2>&1
RL_PPO_ADA_FWD_EARLY=0 python profiles/prof_timeline.py > gpurun_out/r2_ppo_timeline_adafwd0.txt 2>&1
