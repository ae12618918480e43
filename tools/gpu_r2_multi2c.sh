#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/multi_mismatch_evidence.txt
for rep in 1 2; do
  timeout 500 python -m pytest tests/test_gpu_multi.py -q -k "2-" > gpurun_out/m2c_pytest_multi_$rep.log 2>&1; echo "rep $rep rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/m2c_pytest_multi_$rep.log | tail -3
done
cat gpurun_out/multi_mismatch_evidence.txt 2>/dev/null | cut -c1-800
