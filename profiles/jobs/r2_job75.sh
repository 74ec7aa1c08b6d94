set -x
python -m pytest tests/test_ppo_gpu.py tests/test_runner_gpu.py -x -q 2>&1 | grep -v Warning | tail -3
