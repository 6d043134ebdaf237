#!/usr/bin/env python
"""Render a few frames of a bench workload with direct launches — the command that ncu wraps.
usage: python tools/prof_frame.py [workload] [frames] [flags]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vkscanlinepr_b200 as V  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "synth_1m_4k"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
flags = int(sys.argv[3]) if len(sys.argv) > 3 else V.FLAG_NO_GRAPH
sc, rows, W, H = bench.load_workload(wl)
r = V.ScanlineRasterizer(0, flags).initialize(None, W, H)
r.loadVG(sc)
r.setMVP(rows)
if os.environ.get("SLPR_BAND"):  # e.g. SLPR_BAND=3/8: render only band 3 of 8 (independent bands)
    from vkscanlinepr_b200 import parallel as PAR
    b, n = (int(x) for x in os.environ["SLPR_BAND"].split("/"))
    r.set_band(*PAR.band_rows(H, n)[b])
for _ in range(frames):
    r.render()
    r.synchronize()
print(r.counts(), r.stage_ms() if flags & V.FLAG_NO_GRAPH else "")
r.close()
