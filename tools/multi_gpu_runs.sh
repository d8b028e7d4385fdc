#!/bin/bash
# Round-2 multi-GPU evidence (one box, N = $1 GPUs): weak and strong scaling of the subject-data-parallel step, fold-parallel CV.
N=${1:-8}
P=${P:-r2c}
run() { local n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n "$@" 2> gpurun_out/${P}_mg.err | grep '^{' ; }
FAST="--no-cpu-baseline --no-gpu-eager-baseline --no-mlp"
run $N --steps 20 --warmup 5 $FAST > gpurun_out/${P}_bench_${N}gpu.json
run $N --steps 20 --warmup 5 --scaling strong $FAST > gpurun_out/${P}_bench_strong_${N}gpu.json
run $N --steps 20 --warmup 5 --folds 3 > gpurun_out/${P}_bench_folds3_${N}gpu.json
for f in gpurun_out/${P}_bench_${N}gpu.json gpurun_out/${P}_bench_strong_${N}gpu.json gpurun_out/${P}_bench_folds3_${N}gpu.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', (d.get('e2e') or {}).get('value'), [ (f['ranks'], round(f['subjects_per_s'])) for f in d.get('folds',[])], d.get('weights_differ_across_folds'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
tail -3 gpurun_out/${P}_mg.err | cut -c1-300
