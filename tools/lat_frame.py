#!/usr/bin/env python
"""Graph-replay ms/frame of a workload (SLPR_LIB selects the library build). usage: lat_frame.py workload [frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import vkscanlinepr_b200 as V
wl = sys.argv[1]; frames = int(sys.argv[2]) if len(sys.argv) > 2 else 50
sc, rows, W, H = bench.load_workload(wl)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
r = V.ScanlineRasterizer(0, 0).initialize(None, W, H)
r.set_stream(stream.cuda_stream); r.loadVG(sc); r.setMVP(rows)
for _ in range(5): r.render()
r.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(frames): r.render()
e1.record(stream); torch.cuda.synchronize()
print(f"{wl} {os.path.basename(os.environ.get('SLPR_LIB', 'main'))}: {e0.elapsed_time(e1) / frames:.4f} ms/frame  {r.counts()}")
r.close()
