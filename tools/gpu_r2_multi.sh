#!/bin/bash
# round 2, final multi-GPU call (N = all GPUs of the box): the multi-GPU tests, bench at N on C2 / C3 / C4
mkdir -p gpurun_out; rm -f gpurun_out/multi_mismatch_evidence.txt
N=$(nvidia-smi -L | wc -l)
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 900 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/m${N}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; grep -E "passed|failed|FAILED|skipped" gpurun_out/m${N}_pytest_multi.log | tail -12
run 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/m${N}_bench_C2.json 2> gpurun_out/m${N}_bench_C2.err; echo "bench C2 rc=$?"; tail -2 gpurun_out/m${N}_bench_C2.err
run 29535 bench.py --gpus $N --steps 10 --warmup 3 --config C3 > gpurun_out/m${N}_bench_C3.json 2> gpurun_out/m${N}_bench_C3.err; echo "bench C3 rc=$?"; tail -2 gpurun_out/m${N}_bench_C3.err
run 29536 bench.py --gpus $N --steps 10 --warmup 3 --config C4 > gpurun_out/m${N}_bench_C4.json 2> gpurun_out/m${N}_bench_C4.err; echo "bench C4 rc=$?"; tail -2 gpurun_out/m${N}_bench_C4.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/m8_bench_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'N', d['n_gpus'], 'value %.1f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], d['config']['parallelism'], {k: round(v,3) for k,v in (d.get('stage_ms') or {}).items()}, sep='\n   ')
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/multi_mismatch_evidence.txt 2>/dev/null | cut -c1-600
