#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_C2.csv python bench.py --steps 2 --warmup 3 --min-time 0.01 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
for N in 1 2 4 8; do
  timeout 100 ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:rgs:: --csv --log-file gpurun_out/ncu_slab_C2_n$N.csv python tools/slab_step.py --config C2 --world $N --rank $((N/2)) --steps 2 > gpurun_out/ncu_slab_n$N.log 2>&1; echo "slab N=$N rc=$?"
done
python tools/per_n_hbm.py gpurun_out/ncu_slab_C2_n1.csv gpurun_out/ncu_slab_C2_n2.csv gpurun_out/ncu_slab_C2_n4.csv gpurun_out/ncu_slab_C2_n8.csv > gpurun_out/r02_ncu_hbm_per_n_C2.txt; grep "step (" gpurun_out/r02_ncu_hbm_per_n_C2.txt
