set -x
python profiles/jobs/stagger_sweep.py 32768 default 1100,444,296 default 1100,444,296
python -m pytest tests/test_env_gpu.py -x -q 2>&1 | grep -v Warning | tail -3
python bench.py --steps 2000 --warmup 10 --quick > gpurun_out/r2_bench10_quick.json 2> gpurun_out/r2_bench10_quick.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench10_quick.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline'])"
