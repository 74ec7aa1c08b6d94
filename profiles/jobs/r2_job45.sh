set -x
# A/B of the MMA-issuer rewrite (overlapped barrier tests, op words prefetched one op ahead) against the round's previous library
python -m pytest tests/test_chain_gpu.py -x -q 2>&1 | grep -v Warning | tail -4
CHAIN_ONLY=1 python profiles/time_chain.py 2>&1 | grep chain=True
python bench.py --only-ppo | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NEW 4000 envs', d['ms_per_iteration'], d['roofline']['frac'])"
B=196608 PROG=teacher python profiles/trace_chain.py > gpurun_out/r2_trace_teacher_v5.txt 2>&1; head -3 gpurun_out/r2_trace_teacher_v5.txt
cp rapid_locomotion_rl_b200/librl_b200.so /tmp/new.so
cp rapid_locomotion_rl_b200/librl_b200_base.so rapid_locomotion_rl_b200/librl_b200.so
CHAIN_ONLY=1 python profiles/time_chain.py 2>&1 | grep chain=True
python bench.py --only-ppo | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BASE 4000 envs', d['ms_per_iteration'], d['roofline']['frac'])"
cp /tmp/new.so rapid_locomotion_rl_b200/librl_b200.so
