set -x
timeout 600 python -m pytest tests/test_tn_tc_gpu.py -x -q 2>&1 | tail -3
timeout 100 python scripts/tc_gemm_single.py 11 10 13 32 3 | tail -1
timeout 100 python scripts/tc_gemm_single.py 10 9 13 32 3 | tail -1
timeout 100 python scripts/tc_gemm_single.py 12 8 11 32 3 | tail -1
timeout 100 python scripts/tc_gemm_single.py 16 5 9 32 3 | tail -1
timeout 100 python scripts/tc_gemm_single.py 8 13 3 32 3 | tail -1
timeout 300 python scripts/c5_amplitude.py 64 1 2>&1 | sed -n 3,9p
