set -x
timeout 600 python -m pytest tests/test_tn_gpu.py -x -q -k "plugin or c5" 2>&1 | tail -4
