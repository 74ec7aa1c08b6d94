set -x
python -m pytest tests/test_chain_gpu.py -x -q 2>&1 | grep -v Warning | tail -8
CHAIN_ONLY=1 python profiles/time_chain.py 2>&1 | grep chain=True
RL_CHAIN_WORKERS=2 CHAIN_ONLY=1 python profiles/time_chain.py 2>&1 | grep chain=True
B=196608 PROG=teacher python profiles/trace_chain.py > gpurun_out/r2_trace_teacher_w4.txt 2>&1; head -3 gpurun_out/r2_trace_teacher_w4.txt
