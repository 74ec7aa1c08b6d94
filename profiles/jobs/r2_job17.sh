B=196608 PROG=teacher python profiles/trace_chain.py > gpurun_out/r2_trace_teacher_v3.txt 2>&1; head -30 gpurun_out/r2_trace_teacher_v3.txt | cut -c1-200
