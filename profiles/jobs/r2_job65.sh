set -x
# final code: full GPU suite, default bench line, launch list of the quick bench
python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -3
python bench.py > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err
python bench.py --steps 60 --warmup 3 --quick > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_env_step_launches_final3.csv python bench.py --steps 60 --warmup 3 --quick > gpurun_out/r2_ncu65.log 2>&1
ls -la gpurun_out/r2_env_step_launches_final3.csv
