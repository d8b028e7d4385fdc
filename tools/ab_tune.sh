# A/B of a kernel switch on one box: bash tools/ab_tune.sh eigh_k2_warps 2 4   (alternating runs, ms per step / eigh us per graph of each)
KEY=$1; shift
for v in "$@" "$@"; do
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline --no-mlp --tune $KEY=$v 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
e=[o for o in d['roofline_other_kernels'] if 'us_per_graph' in o]
print('$KEY=$v', 'ms_per_step', round(d['ms_per_step'],4), 'subjects/s', round(d['value'],1), 'eigh us/graph', round(e[0]['us_per_graph'],4) if e else None)"
done
