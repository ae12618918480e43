#!/bin/bash
# round 2, call A: baseline state of the tree + stale-memory diagnosis
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/a_pytest.log
timeout 300 python tools/diag_poison.py 2 8 > gpurun_out/a_poison.log 2>&1; echo "poison rc=$?"; tail -40 gpurun_out/a_poison.log
timeout 600 compute-sanitizer --tool initcheck --print-limit 40 python tools/diag_poison.py 8 > gpurun_out/a_initcheck.log 2>&1; echo "initcheck rc=$?"
grep -E "Uninitialized|at rgs::|ERROR SUMMARY" gpurun_out/a_initcheck.log | sort | uniq -c | sort -rn | head -20
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench_ours.json 2> gpurun_out/a_bench_ours.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/a_bench_ours.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['value'], d.get('stage_ms'))
PY
