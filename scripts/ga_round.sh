timeout 600 python -m pytest tests/test_tn_tc_gpu.py -x -q -k "gatherA or batched" 2>&1 | tail -3
for sh in "14 7 9" "15 6 8" "15 8 7" "16 6 6" "17 7 5" "14 5 10"; do
  echo "== $sh images"; TQ_GATHER=0 timeout 120 python scripts/tc_gemm_single.py $sh 32 3 2>&1 | tail -1
  for d in 0 16 14; do echo "== $sh gather debug=$d"; TQ_TC_DEBUG=$d TQ_GATHER=1 timeout 120 python scripts/tc_gemm_single.py $sh 32 3 2>&1 | tail -1; done
done
