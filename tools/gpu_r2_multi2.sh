#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/multi_mismatch_evidence.txt
for rep in 1 2 3 4 5 6; do
  timeout 600 python -m pytest tests/test_gpu_multi.py -q -k "test_sharded_equals_single and 2-" > gpurun_out/m2_pytest_multi_$rep.log 2>&1; echo "rep $rep rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/m2_pytest_multi_$rep.log | tail -3
done
cat gpurun_out/multi_mismatch_evidence.txt 2>/dev/null | cut -c1-2500
