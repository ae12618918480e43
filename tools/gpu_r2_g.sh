#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/g_ncu_step_C2.csv python tools/slab_step.py --config C2 --steps 2 > gpurun_out/g_ncu_step.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(open('gpurun_out/g_ncu_step_C2.csv')))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = {n: i for i, n in enumerate(rows[hi])}
L = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(H): continue
    d = L.setdefault(int(r[H["ID"]]), {"k": r[H["Kernel Name"]][:60]})
    d[r[H["Metric Name"]]] = float(r[H["Metric Value"]].replace(",", "")) * {"usecond": 1e3, "msecond": 1e6, "nsecond": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(r[H["Metric Unit"]], 1.0)
L = list(L.values())
start = max(i for i, d in enumerate(L) if "preprocess_forward" in d["k"])
for d in L[start-3:]:
    t = d.get("gpu__time_duration.sum", 0); b = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    print("%-62s %8.1f us %8.1f MB %7.0f GB/s" % (d["k"], t / 1e3, b / 1e6, b / max(t, 1)))
PY
