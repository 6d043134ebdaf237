#!/usr/bin/env python
"""tests/golden/full_<scene>.npz: the shipped scenes the reference's own parser cannot read (car, chord: rational arcs,
per-element transforms, gradients), through the COMPLETE RVG reader (slpr_vg_load_rvg_full) — the flat
scene arrays plus the per-curve arc weights, so that the GPU box (which has no /root/reference) can render them.
  python tools/make_full_golden.py        (needs /root/reference/workdir/input/rvg)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkscanlinepr_b200 as V  # noqa: E402

SRC = "/root/reference/workdir/input/rvg"
for name in ("car", "chord"):
    sc, vp, _ = V.load_rvg(os.path.join(SRC, name + ".rvg"), full=True)
    out = os.path.join(ROOT, "tests", "golden", f"full_{name}.npz")
    np.savez_compressed(out, vp=vp, pos=sc.pos, pos_path=sc.pos_path, curve_pos_map=sc.curve_pos_map, curve_type=sc.curve_type,
                        curve_path=sc.curve_path, fill_rule=sc.fill_rule, fill_info=sc.fill_info, curve_weight=sc.curve_weight)
    print(name, sc.n_paths, "paths", sc.n_curves, "curves", os.path.getsize(out), "bytes")
