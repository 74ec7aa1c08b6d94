set -x
# adaptation forward of minibatch i at the end of call i's side branch: parity (lagged == serial), then A/B
python -m pytest tests/test_ppo_gpu.py tests/test_runner_gpu.py tests/test_hlp_gpu.py -x -q 2>&1 | grep -v Warning | tail -4
bash profiles/jobs/ppo_ab.sh RL_PPO_ADA_FWD_EARLY=0 RL_PPO_ADA_FWD_EARLY=1 RL_PPO_ADA_FWD_EARLY=0 RL_PPO_ADA_FWD_EARLY=1
for v in 0 1; do
RL_PPO_ADA_FWD_EARLY=$v python bench.py --only-ppo --ppo-envs 32768 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('32768 envs ADA_FWD_EARLY=$v', d['ms_per_iteration'], d['ms_per_iteration_all'], round(d['roofline']['frac'],4))"
done
RL_PPO_ADA_FWD_EARLY=1 python profiles/prof_timeline.py > gpurun_out/r2_ppo_timeline_adafwd.txt 2>&1
