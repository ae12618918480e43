#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for mb in 3 4; do
  echo "RGS_PRE_MINBLOCKS=$mb"
  RGS_PRE_MINBLOCKS=$mb python tools/run_once.py C2 20
  RGS_PRE_MINBLOCKS=$mb python tools/run_once.py C3 10
done
