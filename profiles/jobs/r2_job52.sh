set -x
# final evidence of the round: bench line, launch list, ncu --set full of the shipped env-step kernel (32768 and 4000 envs), per-CTA traces
python bench.py > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err
python bench.py --steps 60 --warmup 3 --quick > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_env_step_launches_final.csv python bench.py --steps 60 --warmup 3 --quick > gpurun_out/r2_ncu52a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_step_rows -s 70 -c 1 -o gpurun_out/r2_env_rows_32k_final -f python bench.py --steps 60 --warmup 3 --quick > gpurun_out/r2_ncu52b.log 2>&1
ENVS=32768 python profiles/trace_env_rows.py > gpurun_out/r2_env_rows_trace_32768_final.txt 2>&1
ENVS=4000 python profiles/trace_env_rows.py > gpurun_out/r2_env_rows_trace_4000_final.txt 2>&1
python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/r2_bench7_ref.json 2> gpurun_out/r2_bench7_ref.err
ls -la gpurun_out | tail -12
