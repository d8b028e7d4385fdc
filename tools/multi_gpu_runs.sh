#!/bin/bash
# Round-2 multi-GPU evidence (one 8-GPU box): weak and strong scaling of the subject-data-parallel step, fold-parallel CV.
run() { local n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $n "$@" 2> gpurun_out/r2_mg.err | grep '^{' ; }
for n in 2 4 8; do run $n --steps 20 --warmup 5 --scaling strong --no-cpu-baseline > gpurun_out/r2_bench_strong_${n}gpu.json; done
run 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_8gpu.json
run 8 --steps 20 --warmup 5 --folds 3 > gpurun_out/r2_bench_folds3_8gpu.json
run 4 --steps 20 --warmup 5 --folds 3 > gpurun_out/r2_bench_folds3_4gpu.json
for f in gpurun_out/r2_bench_strong_*gpu.json gpurun_out/r2_bench_8gpu.json gpurun_out/r2_bench_folds3_*gpu.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('value'), [ (f['ranks'], round(f['subjects_per_s'])) for f in d.get('folds',[])], d.get('weights_differ_across_folds'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
tail -5 gpurun_out/r2_mg.err
