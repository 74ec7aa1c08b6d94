ENVS=32768 python profiles/trace_env_rows.py > gpurun_out/r2_env_rows_trace_32k_stagger.txt 2>&1; cat gpurun_out/r2_env_rows_trace_32k_stagger.txt | tail -30
