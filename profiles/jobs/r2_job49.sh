set -x
# env step: Philox blocks + re-draw test hoisted above the tile wait, trace stamps as a template parameter
python -m pytest tests/test_env_gpu.py -x -q 2>&1 | grep -v Warning | tail -4
python profiles/jobs/env_ab.py BUMP
