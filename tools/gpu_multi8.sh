#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_gpu_multi.py -q -k "2-sparse or 8-auto" > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n$N.err; grep "^{" gpurun_out/bench_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['e2e']['value'], d['stage_ms'])"
