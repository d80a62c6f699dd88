set -x
timeout 300 python -m pytest tests/test_tn_tc_gpu.py -x -q -k "complex128" 2>&1 | tail -3
timeout 100 python scripts/tc_gemm_single.py 11 10 10 32 3 c128 | tail -1
timeout 100 python scripts/tc_gemm_single.py 12 12 8 32 3 c128 | tail -1
TQ_TN_NO_DMMA=1 timeout 100 python scripts/tc_gemm_single.py 11 10 10 32 3 c128 | tail -1
TQ_TN_NO_DMMA=1 timeout 100 python scripts/tc_gemm_single.py 12 12 8 32 3 c128 | tail -1
timeout 300 python -m pytest tests/test_tn_gpu.py -x -q -k "c5 or complex128" 2>&1 | tail -3
