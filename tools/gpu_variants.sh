#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for v in $(ls vkscanlinepr_b200/variants | sed "s/libslpr_//;s/.so//"); do
  echo -n "$v: "; SLPR_LIB=$PWD/vkscanlinepr_b200/variants/libslpr_$v.so timeout 120 python tools/prof_frame.py synth_1m_4k 6 2>&1 | tail -1 | python -c "
import sys,ast
l=sys.stdin.read(); i=l.index('} {')+2; d=ast.literal_eval(l[i:]); print(' '.join('%s=%.3f'%(k[:9],v) for k,v in d.items() if v>0.01), 'total %.3f'%sum(d.values()))"
done
