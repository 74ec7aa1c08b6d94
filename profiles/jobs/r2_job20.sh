CHAIN_ONLY=1 python profiles/time_chain.py 2>&1 | grep chain=True
RL_CHAIN_WORKERS=2 CHAIN_ONLY=1 python profiles/time_chain.py 2>&1 | grep chain=True
