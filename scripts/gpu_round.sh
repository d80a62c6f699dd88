set -x
timeout 600 python -m pytest tests/test_tn_fused_gpu.py tests/test_tn_gpu.py -x -q 2>&1 | tail -3
timeout 300 python scripts/c2tn_run.py
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 10 --csv --log-file gpurun_out/launches_c2tn.csv python scripts/c2tn_run.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_c2tn.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size') if 'Grid Size' in hdr else None
for r in rows[hi+2:hi+12]:
    print(r[ki][:50], r[vi], r[gi] if gi else '')
PY
timeout 300 python scripts/c5_simplified.py 64 1 2>&1 | grep -E "amp|profiled" | tail -2
