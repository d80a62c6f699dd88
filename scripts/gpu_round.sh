set -x
timeout 600 python -m pytest tests/test_tn_tc_gpu.py -x -q 2>&1 | tail -2
timeout 100 python scripts/tc_gemm_single.py 15 8 7 32 3 | tail -1
timeout 100 python scripts/tc_gemm_single.py 12 13 7 32 3 | tail -1
timeout 100 python scripts/tc_gemm_single.py 11 10 13 32 3 | tail -1
timeout 100 python scripts/tc_gemm_single.py 9 14 9 32 3 | tail -1
timeout 600 python bench.py --workload c5 --steps 5 > gpurun_out/bench_c5_reconf.json 2> gpurun_out/bench_c5_reconf.err; tail -c 300 gpurun_out/bench_c5_reconf.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c5_reconf.json'))
print(d['value'], d['ms_per_step'], d['per_slice_ms_profiled'])
for r in d['step_table'][:8]: print(r.get('ms'), r.get('pack_ms'), r.get('M'), r.get('N'), r.get('K'), r.get('frac'))
PY
