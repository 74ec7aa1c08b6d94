set -x
python -c "import torch; print(torch.cuda.Stream(priority=-4).priority, torch.cuda.Stream(priority=-1).priority)"
for c in 1 0; do
RL_PPO_CHUNKS=$c python bench.py --only-ppo | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('CHUNKS=$c 4000 envs', d['ms_per_iteration'], d['roofline']['frac'])"
done
RL_PPO_CHUNKS=1 python bench.py --only-ppo --ppo-envs 32768 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('CHUNKS=1 32768 envs', d['ms_per_iteration'], d['roofline']['frac'])"
python -m pytest tests/test_gymapi_gpu.py tests/test_checkpoint_gpu.py tests/test_env_gpu.py tests/test_runner_gpu.py -q 2>&1 | tail -25
