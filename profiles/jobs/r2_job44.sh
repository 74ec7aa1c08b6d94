python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -4
python bench.py --only-ppo 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_iteration'], d['ms_per_iteration_all'], d['roofline']['frac'])"
python bench.py --only-ppo --ppo-envs 32768 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_iteration'], d['ms_per_iteration_all'], d['roofline']['frac'])"
