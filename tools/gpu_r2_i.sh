#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_refbuild.jsonl
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/i_pytest.log | tail -8
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/i_bench_ref.json 2> gpurun_out/i_bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/i_bench_ours.json 2> gpurun_out/i_bench_ours.err; echo "bench ours rc=$?"; tail -3 gpurun_out/i_bench_ours.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/i_bench_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'value %.1f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.1f (%.3f ms)' % (d['e2e']['value'], d['e2e']['ms_per_step']), {k: round(v,3) for k,v in (d.get('stage_ms') or {}).items()}, d['timing'], sep='\n   ')
    except Exception as e:
        print(f, 'ERR', e)
PY
