set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 1 -c 1 -o gpurun_out/prof_tc_gemm_m11n10k13_v2 python scripts/tc_gemm_single.py 11 10 13 32 2 > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tc_pack -s 2 -c 2 -o gpurun_out/prof_tc_pack_m11n10k13_v2 python scripts/tc_gemm_single.py 11 10 13 32 2 > gpurun_out/ncu_pack.log 2>&1; tail -2 gpurun_out/ncu_pack.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_c5_v5.csv python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/bench_c5_ncu.log 2>&1
timeout 600 python bench.py --workload c5 --steps 5 > gpurun_out/bench_c5_v5.json 2> gpurun_out/bench_c5_v5.err; tail -c 300 gpurun_out/bench_c5_v5.err
