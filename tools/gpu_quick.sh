#!/bin/bash
# quick GPU check: parity tests + per-stage timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
python tools/run_once.py C2 10
python tools/run_once.py C3 5
