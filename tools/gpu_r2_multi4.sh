#!/bin/bash
# round 2, 4-GPU call: multi tests (repeated: an intermittent few-pixel difference is being hunted), bench at N=2 and N=4
mkdir -p gpurun_out
G=$(nvidia-smi -L | wc -l)
run() { n=$1; port=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port "$@"; }
for rep in 1 2 3; do
  timeout 900 python -m pytest tests/test_gpu_multi.py -q -k "not 8" > gpurun_out/m4_pytest_multi_$rep.log 2>&1; echo "pytest multi rep $rep rc=$?"; grep -E "passed|failed|FAILED|skipped" gpurun_out/m4_pytest_multi_$rep.log | tail -4
  grep -E "evidence|AssertionError" gpurun_out/m4_pytest_multi_$rep.log | head -4 | cut -c1-1500
done
for N in 2 4; do
  run $N 2953$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/m4_bench_C2_n$N.json 2> gpurun_out/m4_bench_C2_n$N.err; echo "bench C2 N=$N rc=$?"; tail -2 gpurun_out/m4_bench_C2_n$N.err
done
run 4 29541 bench.py --gpus 4 --steps 10 --warmup 3 --config C3 > gpurun_out/m4_bench_C3_n4.json 2> gpurun_out/m4_bench_C3_n4.err; echo "bench C3 rc=$?"
run 4 29542 bench.py --gpus 4 --steps 10 --warmup 3 --config C4 > gpurun_out/m4_bench_C4_n4.json 2> gpurun_out/m4_bench_C4_n4.err; echo "bench C4 rc=$?"
RGS_EXCHANGE_WINDOW=symm run 4 29543 bench.py --gpus 4 --steps 10 --warmup 3 --config C3 > gpurun_out/m4_bench_C3_n4_symm.json 2> gpurun_out/m4_bench_C3_n4_symm.err; echo "bench C3 symm rc=$?"; tail -3 gpurun_out/m4_bench_C3_n4_symm.err | cut -c1-300
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/m4_bench_*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, 'N', d['n_gpus'], 'value %.1f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], d['config']['parallelism'], {k: round(v,3) for k,v in (d.get('stage_ms') or {}).items()}, sep='\n   ')
    except Exception as e:
        print(f, 'ERR', e)
PY
