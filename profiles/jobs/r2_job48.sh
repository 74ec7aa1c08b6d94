set -x
# source-level ncu of the env step at 4000 envs (one CTA per SM: the pure-latency case)
python bench.py --steps 60 --warmup 3 --quick --envs 4000 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:env_step_rows -s 70 -c 1 -o gpurun_out/r2_env_rows_4000_src -f python bench.py --steps 60 --warmup 3 --quick --envs 4000 > gpurun_out/r2_ncu48.log 2>&1
ls -la gpurun_out/*.ncu-rep
