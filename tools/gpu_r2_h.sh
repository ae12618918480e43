#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_fused.py tests/test_gpu_workflow.py -m gpu -q -x > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/h_pytest.log | tail -6
bash tools/gpu_r2_g.sh 2>&1 | tail -14
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/h_bench_ours.json 2> gpurun_out/h_bench_ours.err; echo "bench ours rc=$?"; tail -3 gpurun_out/h_bench_ours.err
timeout 600 python bench.py --steps 10 --warmup 3 --config C3 --no-cpu-baseline > gpurun_out/h_bench_ours_C3.json 2> gpurun_out/h_bench_ours_C3.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/h_bench_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'value %.1f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.1f (%.3f ms)' % (d['e2e']['value'], d['e2e']['ms_per_step']), {k: round(v,3) for k,v in (d.get('stage_ms') or {}).items()}, sep='\n   ')
    except Exception as e:
        print(f, 'ERR', e)
PY
