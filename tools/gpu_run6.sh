#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "=== stages"; timeout 300 python tools/prof_frame.py synth_1m_4k 6 2>&1 | tail -1
for wl in synth_1m_4k tiger@3840x2160 reschart@1920x1080 synth_16k; do timeout 120 python tools/lat_frame.py $wl 30 2>&1 | tail -1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_piece_emit|k_monotonize" -s 3 -c 3 -f -o gpurun_out/prof_walkwin python tools/prof_frame.py synth_1m_4k 2 > gpurun_out/prof_walkwin.log 2>&1
ncu -i gpurun_out/prof_walkwin.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]
for row in r[2:]:
    g=lambda n: row[h.index(n)] if n in h else '?'
    print(g('Kernel Name')[:40], g('gpu__time_duration.sum'), 'rd', g('dram__bytes_read.sum'), 'wr', g('dram__bytes_write.sum'), 'issue', g('smsp__issue_active.avg.pct_of_peak_sustained_active'))
"
