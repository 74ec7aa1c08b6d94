set -x
python bench.py --steps 60 --warmup 3 --quick > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:env_step_rows -s 70 -c 1 -o gpurun_out/r2_env_rows_32k_final3 -f python bench.py --steps 60 --warmup 3 --quick > gpurun_out/r2_ncu73.log 2>&1
ncu -i gpurun_out/r2_env_rows_32k_final3.ncu-rep --page raw --csv > gpurun_out/r2_env_rows_32k_final3_raw.csv 2>/dev/null
ENVS=32768 python profiles/trace_env_rows.py > gpurun_out/r2_env_rows_trace_32768_final3.txt 2>&1
head -12 gpurun_out/r2_env_rows_trace_32768_final3.txt
