set -x
python -m pytest tests/test_gymapi_gpu.py tests/test_checkpoint_gpu.py tests/test_env_gpu.py tests/test_runner_gpu.py -q 2>&1 | tail -25
