set -x
# full GPU suite on the code with the packed host boundary
python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -6
