# PPO update, 4000 envs: env-knob A/B inside one box (ms per iteration, roofline fraction)
for v in "$@"; do
  env $v python bench.py --only-ppo | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['ms_per_iteration'], d['ms_per_iteration_all'], round(d['roofline']['frac'],4))"
done
