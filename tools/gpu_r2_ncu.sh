#!/bin/bash
# round 2: ncu artefacts (1 GPU).  (a) launch list of the bench command, (b) --set full capture of the dominant kernel, (c) per-N slab shares
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_C2.csv python bench.py --steps 2 --warmup 3 --min-time 0.01 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:render_backward_kernel -s 2 -c 1 -o gpurun_out/r02_render_backward_C2 python tools/slab_step.py --config C2 --steps 3 > gpurun_out/ncu_full_bwd.log 2>&1; echo "full bwd rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:render_forward_kernel -s 2 -c 1 -o gpurun_out/r02_render_forward_C2 python tools/slab_step.py --config C2 --steps 3 > gpurun_out/ncu_full_fwd.log 2>&1; echo "full fwd rc=$?"
for N in 1 2 4 8; do
  timeout 600 ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:rgs:: --csv --log-file gpurun_out/ncu_slab_C2_n$N.csv python tools/slab_step.py --config C2 --world $N --rank $((N/2)) --steps 2 > gpurun_out/ncu_slab_n$N.log 2>&1; echo "slab N=$N rc=$?"
done
python tools/per_n_hbm.py gpurun_out/ncu_slab_C2_n1.csv gpurun_out/ncu_slab_C2_n2.csv gpurun_out/ncu_slab_C2_n4.csv gpurun_out/ncu_slab_C2_n8.csv > gpurun_out/r02_ncu_hbm_per_n_C2.txt; tail -60 gpurun_out/r02_ncu_hbm_per_n_C2.txt
for k in backward forward; do
ncu -i gpurun_out/r02_render_${k}_C2.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys, json
rows = list(csv.reader(sys.stdin))
h, u, v = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum', 'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'smsp__cycles_active.avg', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_pipe_xu.sum', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__thread_inst_executed_per_inst_executed.ratio']
out = {'kernel': v[h.index('Kernel Name')] if 'Kernel Name' in h else ''}
for k in want:
    if k in h:
        i = h.index(k); out[k] = {'value': v[i].replace(',', ''), 'unit': u[i]}
print(json.dumps(out, indent=1))
" > gpurun_out/r02_ncu_render_${k}_C2.json; head -c 1500 gpurun_out/r02_ncu_render_${k}_C2.json
done
