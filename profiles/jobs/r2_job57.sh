set -x
# final evidence refresh: full GPU suite, default bench line, reference arm
python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -3
python bench.py > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err
python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/r2_bench9_ref.json 2> gpurun_out/r2_bench9_ref.err
tail -c 600 gpurun_out/r2_bench9_ref.json
