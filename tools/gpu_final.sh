#!/bin/bash
# end-of-round artefacts: tests, smoke, both bench arms, launch list (ncu, cold cache) of the bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; tail -2 gpurun_out/bench_ref.err
timeout 600 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench ours rc=$?"; tail -2 gpurun_out/bench_ours.err
cat gpurun_out/bench_ref.json gpurun_out/bench_ours.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu rc=$?"
