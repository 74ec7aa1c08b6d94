set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python profiles/prof_rollout.py > gpurun_out/r2_rollout_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_rollout_launches_raw.csv python profiles/prof_rollout.py > gpurun_out/r2_rollout_ncu.log 2>&1
tail -2 gpurun_out/r2_rollout_plain.log
python bench.py --steps 500 --warmup 5 > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err; tail -c 600 gpurun_out/r2_bench0.json
