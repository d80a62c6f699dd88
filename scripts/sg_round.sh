timeout 900 python -m pytest tests/test_tn_gpu.py tests/test_tn_tc_gpu.py -x -q 2>&1 | tail -5
timeout 900 python scripts/c5_groups.py 0 2 3 4 2>&1 | tail -8
