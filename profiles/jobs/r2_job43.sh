python -m pytest tests/test_chain_gpu.py tests/test_ppo_gpu.py -x -q 2>&1 | grep -v Warning | tail -3
for cfg in "3 1" "3 2"; do set -- $cfg; echo "workers $1 issuers $2"; RL_CHAIN_WORKERS=$1 RL_CHAIN_ISSUERS=$2 CHAIN_ONLY=1 python profiles/time_chain.py 2>&1 | grep chain=True | cut -c1-220; done
