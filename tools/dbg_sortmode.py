#!/usr/bin/env python
"""Print the sort mode and the path-size distribution of a bench workload. usage: dbg_sortmode.py [workload]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import vkscanlinepr_b200 as V
wl = sys.argv[1] if len(sys.argv) > 1 else "synth_16k"
sc, rows, W, H = bench.load_workload(wl)
r = V.ScanlineRasterizer(0, V.FLAG_NO_GRAPH).initialize(None, W, H)
r.loadVG(sc); r.setMVP(rows); r.render()
print(wl, r.counts(), r.sort_mode(), "pieces", r.n_pieces())
d = np.diff(r.tap("segments"))
print("paths", len(d), "max", d.max(), "mean", d.mean(), ">256:", (d > 256).sum(), ">512:", (d > 512).sum(), ">4096:", (d > 4096).sum())
print("hist", np.histogram(d, bins=[0, 1, 2, 33, 65, 129, 257, 513, 1025, 4097, 1 << 30])[0])
r.render(); r.synchronize(); print(r.stage_ms())
r.close()
