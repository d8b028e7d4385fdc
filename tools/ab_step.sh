#!/bin/bash
# A/B of kernel switches on ONE box (boxes differ by a few per cent in clocks): ms per step for each --tune set
for t in "" "--tune tcgemm_tmem_a=1" "--tune tcgemm_mn_partial=0" "--tune tcgemm_ragged_rows=0" "" ; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-mlp $t 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('tune', d['config']['tuning'], 'ms', round(d['ms_per_step'],3), 'gemm TF', round(d['roofline']['achieved'],1) if 'tcgemm' in d['roofline']['kernel'] else [round(r['achieved'],1) for r in d['roofline_other_kernels'] if 'tcgemm' in r['kernel']], 'clock', d['clocks']['sm_mhz'])"
done
