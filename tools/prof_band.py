#!/usr/bin/env python
"""One row band of synth_16k on one GPU, exchange kernels included (a one-band "world": the context is its own
peer), for `ncu --metrics gpu__time_duration.sum` launch lists:  python tools/prof_band.py <band> <n_bands> [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vkscanlinepr_b200 as V  # noqa: E402
from vkscanlinepr_b200 import parallel as PAR, scene as S  # noqa: E402

band, G = int(sys.argv[1]), int(sys.argv[2])
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 3
W = H = 16384
sc = S.synth_16k()
r = V.ScanlineRasterizer(0, V.FLAG_NO_GRAPH).initialize(None, W, H)
r.loadVG(sc)
r.setMVP(S.identity_rows())
if G > 1:
    r.set_band(*PAR.band_rows(H, G)[band])
    box, _ = r.band_mailbox()
    r.set_band_peers(1, 0, 0, [box])
acc = {}
for i in range(frames):
    if G > 1:
        r.render_band(i + 1)
    else:
        r.render()
    r.synchronize()
    if i:
        for k, v in r.stage_ms().items():
            acc[k] = acc.get(k, 0.0) + v / (frames - 1)
print("band %d/%d fragments %d" % (band, G, r.counts()["n_fragments"]), {k: round(v, 3) for k, v in acc.items()}, "sum %.3f" % sum(acc.values()))
