set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python -m pytest tests/test_runner_gpu.py tests/test_env_gpu.py -q 2>&1 | tail -8
