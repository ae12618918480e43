#!/bin/bash
# round 2, call C (1 GPU): new tests, bench protocol on both arms, gradient noise A/B
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x  > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/c_pytest.log
for cfg in C1 C2; do timeout 300 python tools/grad_noise.py --cfg $cfg > gpurun_out/c_noise_$cfg.txt 2>&1; cat gpurun_out/c_noise_$cfg.txt | tail -12; done
LD_LIBRARY_PATH=$PWD/build_ab/precdiv timeout 300 python tools/grad_noise.py --cfg C1 > gpurun_out/c_noise_C1_precdiv.txt 2>&1; tail -12 gpurun_out/c_noise_C1_precdiv.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c_bench_ref.json 2> gpurun_out/c_bench_ref.err; echo "bench ref rc=$?"; tail -3 gpurun_out/c_bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c_bench_ours.json 2> gpurun_out/c_bench_ours.err; echo "bench ours rc=$?"; tail -3 gpurun_out/c_bench_ours.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/c_bench_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'value %.1f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], d.get('timing'), d['e2e'].get('timing'), d.get('stage_ms'), d['clocks'], sep='\n   ')
    except Exception as e:
        print(f, 'ERR', e)
PY
grep -l native gpurun_out/c_bench_ref.json >/dev/null; python - <<'PY'
# which of our .so files did the reference arm map?  (must be none)
import subprocess, sys
PY
