set -x
python -m pytest tests/test_gymapi_gpu.py tests/test_checkpoint_gpu.py -q -x 2>&1 | grep -v Warning | tail -60
