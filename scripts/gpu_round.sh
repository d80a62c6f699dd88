set -x
timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -k "after_measurement or errors or reference_style" 2>&1 | tail -5
