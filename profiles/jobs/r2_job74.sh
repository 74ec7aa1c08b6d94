set -x
# all minibatch steps of an update in one CUDA graph: parity, then A/B
python -m pytest tests/test_ppo_gpu.py tests/test_runner_gpu.py tests/test_hlp_gpu.py tests/test_checkpoint_gpu.py -x -q 2>&1 | grep -v Warning | tail -4
bash profiles/jobs/ppo_ab.sh RL_PPO_ONE_GRAPH=0 RL_PPO_ONE_GRAPH=1 RL_PPO_ONE_GRAPH=0 RL_PPO_ONE_GRAPH=1
for v in 0 1; do
RL_PPO_ONE_GRAPH=$v python bench.py --only-ppo --ppo-envs 32768 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('32768 envs ONE_GRAPH=$v', d['ms_per_iteration'], d['ms_per_iteration_all'], round(d['roofline']['frac'],4))"
done
