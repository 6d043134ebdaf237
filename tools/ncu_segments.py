#!/usr/bin/env python
"""Summarise an ncu source page: per-kernel totals and code segments by execution count.
usage: ncu_segments.py report.ncu-rep [kernel-substring]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for si, st in enumerate(starts):
    name = rows[st - 1][1] if st > 0 and len(rows[st - 1]) > 1 else "?"
    if want and want not in name:
        continue
    hdr = rows[st]; end = (starts[si + 1] - 1) if si + 1 < len(starts) else len(rows)
    ia, isrc, ie, it, iss = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "Thread Instructions Executed", "Warp Stall Sampling (All Samples)"))
    data = [(int(r[ia], 16), r[isrc].strip(), int(r[ie]), int(r[it]), int(r[iss]) if r[iss].isdigit() else 0)
            for r in rows[st + 1:end] if len(r) > iss and r[ie].isdigit()]
    if not data:
        continue
    base = data[0][0]; tot = sum(d[2] for d in data); ts = sum(d[4] for d in data)
    print(f"== {name[:70]}  warp-instr={tot} thread-instr={sum(d[3] for d in data)} avg-thr={sum(d[3] for d in data)/max(tot,1):.1f} samples={ts}")
    seg = []; cur = None
    for a, s, e, t, sm in data:
        if cur and abs(e - cur["e"]) <= 0.02 * max(cur["e"], 1) + 2:
            cur["n"] += 1; cur["sum"] += e; cur["tsum"] += t; cur["end"] = a; cur["sm"] += sm
        else:
            if cur: seg.append(cur)
            cur = dict(start=a, end=a, e=e, n=1, sum=e, tsum=t, first=s, sm=sm)
    seg.append(cur)
    for s in seg:
        if s["sum"] > 0.01 * tot or s["sm"] > 0.02 * ts:
            print("  %05x-%05x n=%3d exec=%9d instr%%=%5.1f stall%%=%5.1f thr=%5.1f  %s" % (
                s["start"] - base, s["end"] - base, s["n"], s["e"], 100 * s["sum"] / tot, 100 * s["sm"] / max(ts, 1),
                s["tsum"] / max(s["sum"], 1), s["first"][:44]))
