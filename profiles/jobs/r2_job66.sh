set -x
python -m pytest tests/test_ppo_gpu.py -x -q -k "lagged or fused or gather" 2>&1 | grep -v Warning | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
bash profiles/jobs/ppo_ab.sh RL_WGRAD_KB=24 RL_WGRAD_KB=32 RL_WGRAD_KB=40 RL_WGRAD_KB=48 RL_WGRAD_KB=32
