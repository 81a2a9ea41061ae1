timeout 600 python -m pytest tests/test_gpu_step.py -q -m gpu -x -k terrain 2>&1 | grep -E "^E   +Assertion|UserWarning: parity" | cut -c1-600
