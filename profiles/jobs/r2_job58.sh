set -x
# four-rows-per-warp policy-row gather: parity, then A/B of the update at both sizes
python -m pytest tests/test_ppo_gpu.py tests/test_chain_gpu.py tests/test_hlp_gpu.py -x -q 2>&1 | grep -v Warning | tail -4
bash profiles/jobs/ppo_ab.sh RL_PPO_GATHER_ROWS=0 RL_PPO_GATHER_ROWS=1 RL_PPO_GATHER_ROWS=0 RL_PPO_GATHER_ROWS=1
for v in 0 1; do
RL_PPO_GATHER_ROWS=$v python bench.py --only-ppo --ppo-envs 32768 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('32768 envs GATHER_ROWS=$v', d['ms_per_iteration'], d['ms_per_iteration_all'], round(d['roofline']['frac'],4))"
done
