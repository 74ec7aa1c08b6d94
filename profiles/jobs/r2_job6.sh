set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python bench.py --steps 500 --warmup 5 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -c 1500 gpurun_out/r2_bench2.json; tail -5 gpurun_out/r2_bench2.err
