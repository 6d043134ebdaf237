#!/bin/bash
# A/B of tuning variants built by tools/variants.sh (run under gpurun): tools/ab.sh name1 name2 ... (each twice)
for v in "$@" "$@"; do
  SLPR_LIB=vkscanlinepr_b200/variants/libslpr_$v.so python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-radix-leg --no-scenes 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms']
print('$v', round(d['ms_per_step'],4), ' '.join(f'{k}={s[k]:.4f}' for k in ('transform','monotonize','piece_setup','walk','sort_passes','wind_scan','span_emit')))"
done
