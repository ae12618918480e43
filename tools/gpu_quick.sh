#!/bin/bash
# quick GPU check: parity tests + per-stage timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
python tools/run_once.py C2 10
timeout 600 python tools/compare_ref.py --cfg C4 --time --iters 3 > gpurun_out/cmp_C4.log 2>&1; echo "C4 rc=$?"; grep -v "^  ref/ref" gpurun_out/cmp_C4.log | cut -c1-170 | tail -45
