#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python tools/run_once.py C2 10
python tools/run_once.py C3 5
ncu --set full --clock-control none --import-source on -k regex:render_backward -s 1 -c 1 -o gpurun_out/prof_bwd python tools/run_once.py C2 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_forward -s 1 -c 1 -o gpurun_out/prof_fwd python tools/run_once.py C2 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:preprocess_backward -s 1 -c 1 -o gpurun_out/prof_pbwd python tools/run_once.py C2 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
