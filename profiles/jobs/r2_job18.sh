export RL_NVCC_EXTRA="-DRL_CHAIN_TRACE_WRITE"
B=196608 PROG=teacher python profiles/trace_chain.py > gpurun_out/r2_trace_teacher_v4.txt 2>&1; sed -n 56,100p gpurun_out/r2_trace_teacher_v4.txt | cut -c1-200
