set -x
python -m pytest tests/test_chain_gpu.py -x -q 2>&1 | grep -v Warning | tail -8
python profiles/time_chain.py 2>&1 | grep chain=True
B=196608 PROG=teacher python profiles/trace_chain.py > gpurun_out/r2_trace_teacher_v2.txt 2>&1; head -3 gpurun_out/r2_trace_teacher_v2.txt
B=196608 PROG=trunk_backward python profiles/trace_chain.py > gpurun_out/r2_trace_bwd_v2.txt 2>&1; head -3 gpurun_out/r2_trace_bwd_v2.txt
