set -x
timeout 900 python -m pytest tests/test_tn_gpu.py -x -q -k "tree or gradients or plugin" 2>&1 | tail -15
