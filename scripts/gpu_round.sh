set -x
timeout 600 python -m pytest tests/test_tn_gpu.py -x -q -k "without_parameters or plugin" 2>&1 | tail -8
