for v in pair2 xf pair2 xf; do
  SLPR_LIB=vkscanlinepr_b200/variants/libslpr_$v.so python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-radix-leg --no-scenes 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', round(d['ms_per_step'],4), 'walk', round(d['stage_ms']['walk'],4), 'xf', round(d['stage_ms']['transform'],4))"
done
