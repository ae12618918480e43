#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ours_C2.csv python tools/run_once.py C2 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/launches_ours_C2.csv')))
hi = next(i for i,r in enumerate(rows) if r and r[0]=='ID')
hdr = rows[hi]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
data = [(r[ki], float(r[vi].replace(',',''))/1000) for r in rows[hi+1:] if len(r)>vi]
n = len(data)//3
for name, us in data[-n:]:
    print(f"{us:9.1f} us  {name[:90]}")
PY
