#!/bin/bash
# round 2, call B (2 GPUs): the peer exchange at world 2 + new bench protocol, both arms at N=1, ours at N=2
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "2-" > gpurun_out/b_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -15 gpurun_out/b_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/diag_multi.py > gpurun_out/b_diag_multi.log 2>&1; echo "diag rc=$?"; grep "rank" gpurun_out/b_diag_multi.log | head -30
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/b_bench_ref.json 2> gpurun_out/b_bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/b_bench_ours.json 2> gpurun_out/b_bench_ours.err; echo "bench ours rc=$?"; tail -3 gpurun_out/b_bench_ours.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/b_bench_n$N.json 2> gpurun_out/b_bench_n$N.err; echo "bench N rc=$?"; tail -3 gpurun_out/b_bench_n$N.err
RGS_GRAD_EXCHANGE=dense timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/b_bench_n${N}_dense.json 2> gpurun_out/b_bench_n${N}_dense.err; echo "bench N dense rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/b_bench_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'value %.1f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], d.get('timing'), d.get('stage_ms'), d['clocks'])
    except Exception as e:
        print(f, 'ERR', e)
PY
