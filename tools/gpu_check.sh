#!/bin/bash
# GPU-box check used during development: tests, smoke, both bench arms.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python tools/gen_golden.py precomp > gpurun_out/gen_golden_precomp.log 2>&1
cp gpurun_out/golden/precomp.npz tests/golden/ 2>/dev/null
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; tail -2 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 600 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench ours rc=$?"; tail -2 gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json
