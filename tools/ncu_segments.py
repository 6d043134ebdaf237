#!/usr/bin/env python
"""Summarise an ncu source page: per-kernel totals and code segments by execution count, with the
dominant stall reasons per segment.   usage: ncu_segments.py report.ncu-rep [kernel-substring]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
seen = set()
for si, st in enumerate(starts):
    name = rows[st - 1][1] if st > 0 and len(rows[st - 1]) > 1 else "?"
    if (want and want not in name):
        continue
    hdr = rows[st]; end = (starts[si + 1] - 1) if si + 1 < len(starts) else len(rows)
    ia, isrc, ie, it, iss = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "Thread Instructions Executed", "Warp Stall Sampling (All Samples)"))
    reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[st + 1:end]:
        if len(r) > iss and r[ie].isdigit():
            rs = {h: int(r[i]) for i, h in reasons if r[i].isdigit() and int(r[i])}
            data.append((int(r[ia], 16), r[isrc].strip(), int(r[ie]), int(r[it]), int(r[iss]) if r[iss].isdigit() else 0, rs))
    if not data:
        continue
    base = data[0][0]; tot = sum(d[2] for d in data); ts = sum(d[4] for d in data)
    key = (name[:40], tot)
    if key in seen: continue
    seen.add(key)
    allr = {}
    for d in data:
        for k, v in d[5].items(): allr[k] = allr.get(k, 0) + v
    print(f"== {name[:70]}  warp-instr={tot} avg-thr={sum(d[3] for d in data)/max(tot,1):.1f} samples={ts}")
    print("   stalls:", ", ".join(f"{k[6:]}={100*v/max(ts,1):.0f}%" for k, v in sorted(allr.items(), key=lambda kv: -kv[1])[:8]))
    seg = []; cur = None
    for a, s, e, t, sm, rs in data:
        if cur and abs(e - cur["e"]) <= 0.02 * max(cur["e"], 1) + 2:
            cur["n"] += 1; cur["sum"] += e; cur["tsum"] += t; cur["end"] = a; cur["sm"] += sm
            for k, v in rs.items(): cur["rs"][k] = cur["rs"].get(k, 0) + v
        else:
            if cur: seg.append(cur)
            cur = dict(start=a, end=a, e=e, n=1, sum=e, tsum=t, first=s, sm=sm, rs=dict(rs))
    seg.append(cur)
    for s in seg:
        if s["sum"] > 0.015 * tot or s["sm"] > 0.03 * ts:
            top = ",".join(f"{k[6:]}:{100*v/max(ts,1):.0f}" for k, v in sorted(s["rs"].items(), key=lambda kv: -kv[1])[:3])
            print("  %05x-%05x n=%3d exec=%9d instr%%=%5.1f stall%%=%5.1f thr=%5.1f [%s] %s" % (
                s["start"] - base, s["end"] - base, s["n"], s["e"], 100 * s["sum"] / tot, 100 * s["sm"] / max(ts, 1),
                s["tsum"] / max(s["sum"], 1), top, s["first"][:36]))
