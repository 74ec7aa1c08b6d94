for i in 1 2; do
python bench.py --only-ppo 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_iteration'], d['ms_per_iteration_all'], d['clocks'])"
python bench.py --only-ppo --ppo-envs 32768 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_iteration'], d['ms_per_iteration_all'], d['clocks'])"
done
