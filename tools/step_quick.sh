timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo rc=$?; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench_a.json') if x.startswith('{')]
d=json.loads(l[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
for r in [d['roofline']]+d['roofline_other_kernels']:
    print(r['kernel'][:40], 'achieved',r['achieved'],'share',r['share_of_step'], 'us_per_graph', r.get('us_per_graph'))
PY
