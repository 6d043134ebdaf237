import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import vkscanlinepr_b200 as V
from vkscanlinepr_b200 import scene as S
from oracle import oracle_py as O
W = H = 512
rows = S.identity_rows()
for seed in range(3, 40):
    sc = S.synth_scene(1024, 512, 512, 6.0, 40.0, seed=seed)
    ref = O.render(sc, rows, W, H)
    res = np.zeros(sc.n_paths, np.int64); np.add.at(res, ref["path"], ref["wind"])
    if not res.any(): break
print("seed", seed, ref["n_fragments"])
for (y0, y1) in [(0, 256), (256, 512), (0, 512)]:
    r = V.ScanlineRasterizer(0, V.FLAG_TAPS | V.FLAG_NO_GRAPH).initialize(None, W, H)
    r.loadVG(sc); r.setMVP(rows); r.set_band(y0, y1); r.render()
    img = r.readback()
    a, b = img[H - y1:H - y0], ref["rgba"][H - y1:H - y0]
    d = np.any(a != b, axis=2)
    ys, xs = np.nonzero(d)
    print("band", y0, y1, "diff px", d.sum(), "scan rows", sorted(set((H - 1 - (ys + H - y1)).tolist()))[:20], r.counts(), "wn_total", r.tap("winding_scan")[-1])
    if d.sum():
        y, x = ys[0], xs[0]
        print("  first diff at img row", y + H - y1, "x", x, a[y, x], b[y, x])
    r.close()
