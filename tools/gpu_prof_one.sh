#!/bin/bash
# usage: gpu_prof_one.sh <kernel-regex> <out-name> [cfg]
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$1 -s 1 -c 1 -o gpurun_out/$2 python tools/run_once.py ${3:-C2} 2 > /dev/null 2>&1
ls -la gpurun_out/$2.ncu-rep
