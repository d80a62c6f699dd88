set -x
TQ_BENCH_EXTRAS=c2tn,c4tn timeout 600 python bench.py --steps 5 > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err; tail -c 300 gpurun_out/bench_v11.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_v11.json'))
for k,v in d['other_configs'].items(): print(k, round(v['value'],2), round(v['ms_per_step'],3), v.get('fwd_only_ms'), v.get('tree_backward_ms_per_step'), v.get('steps_by_kernel'))
PY
