RL_CHAIN_WORKERS=3 RL_CHAIN_ISSUERS=2 B=196608 PROG=teacher python profiles/trace_chain.py > gpurun_out/r2_trace_teacher_w3i2.txt 2>&1; head -3 gpurun_out/r2_trace_teacher_w3i2.txt
