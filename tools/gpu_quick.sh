#!/bin/bash
# quick GPU check: parity tests + per-stage timings + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
python tools/run_once.py C2 10
python tools/run_once.py C3 5
python tools/compare_ref.py --cfg C3 --ks 0.1 2>&1 | grep -B3 -A12 "^gradients" | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench ours rc=$?"; tail -2 gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json
