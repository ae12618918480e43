#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_multi.log
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
