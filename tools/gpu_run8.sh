#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in $(ls vkscanlinepr_b200/variants | sed "s/libslpr_//;s/.so//"); do
  export SLPR_LIB=$PWD/vkscanlinepr_b200/variants/libslpr_$v.so
  echo "=== variant $v"
  timeout 120 python tools/prof_frame.py synth_1m_4k 6 2>&1 | tail -1 | cut -c100-
  for wl in synth_1m_4k synth_16k; do timeout 120 python tools/lat_frame.py $wl 30 2>&1 | tail -1; done
done
export SLPR_LIB=$PWD/vkscanlinepr_b200/variants/libslpr_w128k_l16.so
timeout 600 ncu --set full --clock-control none -k regex:"k_walk" -s 1 -c 1 -f -o gpurun_out/prof_walkwin2 python tools/prof_frame.py synth_1m_4k 2 > gpurun_out/prof_walkwin2.log 2>&1
ncu -i gpurun_out/prof_walkwin2.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]
for row in r[2:]:
    g=lambda n: row[h.index(n)] if n in h else '?'
    print(g('Kernel Name')[:40], g('gpu__time_duration.sum'), 'rd', g('dram__bytes_read.sum'), 'wr', g('dram__bytes_write.sum'), 'active avg/max', g('sm__cycles_active.avg'), g('sm__cycles_active.max'))
"
