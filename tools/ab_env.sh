# A/B of an environment switch on one box: bash tools/ab_env.sh GSR_OVERLAP_W 1 0   (alternating runs, ms per step of each)
VAR=$1; A=$2; B=$3
for v in $A $B $A $B; do
  env $VAR=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-eager-baseline --no-mlp 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$VAR=$v', 'ms_per_step', round(d['ms_per_step'],4), 'subjects/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
done
