#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2 3 4 5; do
  echo "=== launch $rep (torchrun, OMP_NUM_THREADS set by torchrun)"
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$rep tools/diag_multi3.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM_THREADS environment\|^$" | grep -E "differ|mismatch|threads|record" | grep -v ": 0 elements\|: 0 mismatches"
done
for rep in 1 2 3; do
  echo "=== launch $rep with OMP_NUM_THREADS unset-like (64)"
  OMP_NUM_THREADS=64 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$rep tools/diag_multi3.py 2>&1 | grep -E "differ|mismatch|threads|record" | grep -v ": 0 elements\|: 0 mismatches"
done
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" | head
