set -x
timeout 600 python -m pytest tests/test_tn_tc_gpu.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --workload c5 --steps 5 > gpurun_out/bench_c5_x.json 2> gpurun_out/bench_c5_x.err; tail -c 200 gpurun_out/bench_c5_x.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c5_x.json'))
print(d['value'], d['ms_per_step'], d['per_slice_ms_profiled'])
for r in d['step_table'][:9]: print(r.get('ms'), r.get('pack_ms'), r.get('M'), r.get('N'), r.get('K'))
PY
