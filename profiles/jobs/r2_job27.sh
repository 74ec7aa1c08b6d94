export RL_NVCC_EXTRA="-DRL_CHAIN_TRACE_WRITE"
RL_CHAIN_WORKERS=4 B=196608 PROG=teacher python profiles/trace_chain.py > gpurun_out/r2_trace_teacher_w4b.txt 2>&1; sed -n 44,70p gpurun_out/r2_trace_teacher_w4b.txt | cut -c1-200
RL_CHAIN_WORKERS=4 BS=196608 CHAIN_ONLY=1 python profiles/time_chain.py 2>&1 | grep chain=True | cut -c1-150
