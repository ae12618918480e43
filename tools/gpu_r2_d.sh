#!/bin/bash
# round 2, call D (1 GPU): the whole GPU suite (no -x), poison/initcheck after the hitmask memset
mkdir -p gpurun_out; rm -f gpurun_out/parity_refbuild.jsonl
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/d_pytest.log | tail -30
timeout 300 python tools/diag_poison.py 8 > gpurun_out/d_poison.log 2>&1; echo "poison rc=$?"; tail -2 gpurun_out/d_poison.log
timeout 600 compute-sanitizer --tool initcheck --print-limit 10 python tools/diag_poison.py 8 > gpurun_out/d_initcheck.log 2>&1; echo "initcheck rc=$?"
grep -E "ERROR SUMMARY" gpurun_out/d_initcheck.log; grep -E "Device Frame" gpurun_out/d_initcheck.log | sort | uniq -c | head
cat gpurun_out/parity_refbuild.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['cfg'], d['ks'], 'R', d['num_rendered'], 'ncontrib', d['n_contrib_mismatch'], {k: '%.1e' % v for k, v in d['images_max_abs'].items()})
    for k, v in d['grads'].items(): print('    %-10s' % k, {a: '%.2e' % b for a, b in v.items()})
"
