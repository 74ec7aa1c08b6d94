set -x
python -m pytest tests/test_ppo_gpu.py tests/test_chain_gpu.py -q -x 2>&1 | tail -6
for c in 1 0; do
RL_PPO_CHUNKS=$c python bench.py --only-ppo | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('CHUNKS=$c 4000 envs', d['ms_per_iteration'], d['roofline']['frac'])"
done
for c in 1 0; do
RL_PPO_CHUNKS=$c python bench.py --only-ppo --ppo-envs 32768 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('CHUNKS=$c 32768 envs', d['ms_per_iteration'], d['roofline']['frac'])"
done
