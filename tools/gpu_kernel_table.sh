#!/bin/bash
# per-kernel table of one fused training-step slice at C2: duration, DRAM bytes, SM / DRAM utilisation
mkdir -p gpurun_out
python tools/step_all.py C2 10 | tee gpurun_out/step_all_fused.json
python tools/step_all.py C2 10 --eager | tee gpurun_out/step_all_eager.json
timeout 900 ncu --clock-control none --csv --log-file gpurun_out/kernel_table_C2.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active \
  --kernel-name-base demangled -k regex:rgs:: -c 80 python tools/step_all.py C2 1 > gpurun_out/kernel_table.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/kernel_table_C2.csv
