set -x
timeout 600 python -m pytest tests/test_tn_tc_gpu.py -x -q -k "apply" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_tn_gpu.py tests/test_tn_fused_gpu.py -x -q 2>&1 | tail -3
TQ_BENCH_EXTRAS=c2tn,c4tn,c5,c5s timeout 600 python bench.py --steps 5 > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err; tail -c 300 gpurun_out/bench_v10.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v10.json'))
for k,v in d['other_configs'].items(): print(k, round(v['value'],2), round(v['ms_per_step'],3), v.get('fwd_only_ms'), v.get('steps_by_kernel'))
PY
