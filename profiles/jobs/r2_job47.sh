set -x
# env step: per-warp roles rebalanced (frames split over warps 0 / 2, collision + gravity noise on warp 3)
python -m pytest tests/test_env_gpu.py -x -q 2>&1 | grep -v Warning | tail -4
python profiles/jobs/env_ab.py NEW
python profiles/trace_env_rows.py 32768 > gpurun_out/r2_env_rows_trace_32768_v3.txt 2>&1; head -16 gpurun_out/r2_env_rows_trace_32768_v3.txt
python profiles/trace_env_rows.py 4000 > gpurun_out/r2_env_rows_trace_4000_v3.txt 2>&1; head -16 gpurun_out/r2_env_rows_trace_4000_v3.txt
cp rapid_locomotion_rl_b200/librl_b200.so /tmp/new.so
cp rapid_locomotion_rl_b200/librl_b200_base.so rapid_locomotion_rl_b200/librl_b200.so
python profiles/jobs/env_ab.py BASE
cp /tmp/new.so rapid_locomotion_rl_b200/librl_b200.so
