set -x
timeout 600 python -m pytest tests/test_engine_gpu.py -x -q 2>&1 | tail -2
TQ_BENCH_EXTRAS=c1,c3,c2tn timeout 600 python bench.py --steps 10 > gpurun_out/bench_v13.json 2> gpurun_out/bench_v13.err; tail -c 300 gpurun_out/bench_v13.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v13.json'))
print('c2', d['value'], d['ms_per_step'], d['roofline']['fwd_ms'], d['roofline']['bwd_ms'], d['cpu_baseline']['sample'][-40:])
for k,v in d['other_configs'].items(): print(k, round(v['value'],2), round(v['ms_per_step'],3), v.get('roofline',{}).get('fwd_ms'), v.get('roofline',{}).get('bwd_ms'))
PY
