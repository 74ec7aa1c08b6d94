set -x
bash profiles/jobs/ppo_ab.sh RL_WGRAD_KB=48 RL_WGRAD_KB=56 RL_WGRAD_KB=64 RL_WGRAD_KB=75 RL_WGRAD_KB=94 RL_WGRAD_KB=48 RL_WGRAD_KB=32
for v in 96 128 192 256; do
RL_WGRAD_KB=$v python bench.py --only-ppo --ppo-envs 32768 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('32768 envs KB=$v', d['ms_per_iteration'], d['ms_per_iteration_all'], round(d['roofline']['frac'],4))"
done
