python bench.py > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; tail -c 600 gpurun_out/r2_bench5.json; tail -3 gpurun_out/r2_bench5.err
python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/r2_bench5_ref.json 2> gpurun_out/r2_bench5_ref.err; tail -c 300 gpurun_out/r2_bench5_ref.json
