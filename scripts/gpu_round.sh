set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_v9.json 2> gpurun_out/bench_v9.err; tail -c 300 gpurun_out/bench_v9.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v9.json'))
print(d['value'], d['e2e']['value'], d['cpu_baseline']['value'])
for k,v in d['other_configs'].items(): print(k, round(v['value'],2), round(v['ms_per_step'],3), v.get('fwd_only_ms'), v.get('steps_by_kernel'))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tn_gemm_dmma -s 1 -c 1 -o gpurun_out/prof_dmma python scripts/tc_gemm_single.py 11 10 10 32 2 c128 > gpurun_out/ncu_dmma.log 2>&1; tail -1 gpurun_out/ncu_dmma.log
