set -x
python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -3
python bench.py --only-ppo | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('4000 envs', d['ms_per_iteration'], d['roofline']['frac'], d['c5_32768_envs']['ms_per_iteration'] if 'c5_32768_envs' in d else '')"
