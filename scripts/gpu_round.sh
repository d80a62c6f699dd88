set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -c 300 gpurun_out/bench_final2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final2.json'))
print('c2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
for k,v in d['other_configs'].items(): print(k, round(v['value'],2), round(v['ms_per_step'],3), v.get('fwd_only_ms'), v.get('tree_backward_ms_per_step'), (v.get('roofline') or {}).get('dominant_steps_frac_of_complex_gemm_roofline'))
PY
