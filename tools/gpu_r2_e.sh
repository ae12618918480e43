#!/bin/bash
# round 2, call E (1 GPU): full suite again, statistics, ncu artefacts, both bench arms
mkdir -p gpurun_out; rm -f gpurun_out/parity_refbuild.jsonl
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/e_pytest.log | tail -20
timeout 300 python tools/diag_stats.py C2 C3 > gpurun_out/e_stats.txt 2>&1; cat gpurun_out/e_stats.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/e_bench_ref.json 2> gpurun_out/e_bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench_ours.json 2> gpurun_out/e_bench_ours.err; echo "bench ours rc=$?"; tail -3 gpurun_out/e_bench_ours.err
for cfg in C1 C3 C4; do timeout 600 python bench.py --steps 10 --warmup 3 --config $cfg --no-cpu-baseline > gpurun_out/e_bench_ours_$cfg.json 2> gpurun_out/e_bench_ours_$cfg.err; echo "bench $cfg rc=$?"; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/e_bench_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'value %.1f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], d.get('stage_ms'), sep='\n   ')
    except Exception as e:
        print(f, 'ERR', e)
PY
bash tools/gpu_r2_ncu.sh
