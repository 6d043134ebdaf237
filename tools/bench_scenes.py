#!/usr/bin/env python
"""BASELINE cfg1/cfg2: every shipped scene the reference parser accepts, at 1024^2 / 1080p / 4K: ms per frame
(CUDA-graph replay, CUDA events on the context's stream), Mpixel/s, and whether the RGBA8 frame and the counts
are identical to the oracle. Prints a markdown table (tools/gpu: run under gpurun)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import vkscanlinepr_b200 as V
from vkscanlinepr_b200 import scene as S
from oracle import oracle_py as O

scenes = ["test", "tiger", "reschart", "drops", "embrace"]
FLAGS = int(os.environ.get("SLPR_SCENE_FLAGS", "0"))  # e.g. 8 = SLPR_FLAG_RADIX_SORT
sizes = [(1024, 1024), (1920, 1080), (3840, 2160)]
frames = 50
print("| scene | size | curves | fragments | records | ms/frame | Mpixel/s | oracle ms (threads) | RGBA8 vs oracle | sort |")
print("|---|---|---:|---:|---:|---:|---:|---:|---|---|")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
for name in scenes:
    c = S.Container.from_npz(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    sc = V.flatten(c, name)
    for W, H in sizes:
        rows = S.fit_rows(c.vp, W, H, centred=not (name == "test" and W == 1024))
        r = V.ScanlineRasterizer(0, FLAGS).initialize(None, W, H)
        r.set_stream(stream.cuda_stream); r.loadVG(sc); r.setMVP(rows)
        for _ in range(5): r.render()
        r.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(frames): r.render()
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / frames
        img = r.readback(); cnt = r.counts()
        t = time.perf_counter(); ref = O.render(sc, rows, W, H); tor = (time.perf_counter() - t) * 1e3
        same = np.array_equal(img, ref["rgba"]) and cnt["n_fragments"] == ref["n_fragments"]
        diff = int(np.abs(img.astype(int) - ref["rgba"].astype(int)).max())
        print(f"| {name} | {W}x{H} | {sc.n_curves} | {cnt['n_fragments']} | {cnt['n_out_frag'] + cnt['n_span']} | {ms:.3f} | {W*H/ms/1e3:.0f} | "
              f"{tor:.0f} ({O.num_threads()}) | {'identical' if same else 'max diff %d' % diff} | {r.sort_mode()} |")
        r.close()
# the shipped scenes only the complete RVG reader can load (SLPR_FLAG_FULL_RVG), and f-3: 4 samples per pixel
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
for name, extra, label in (("car", V.FLAG_FULL_RVG, "full-RVG"), ("chord", V.FLAG_FULL_RVG, "full-RVG"), ("car", V.FLAG_FULL_RVG | V.FLAG_AA4, "full-RVG + AA4"),
                           ("car", V.FLAG_FULL_RVG | V.FLAG_BLEND, "full-RVG + blend"), ("car", V.FLAG_FULL_RVG | V.FLAG_BLEND | V.FLAG_AA4, "full-RVG + blend + AA4")):
    sc, vp = util.full_golden_scene(name)
    for W, H in sizes[1:]:
        if (extra & V.FLAG_AA4) and W > 1920:
            continue
        rows = S.fit_rows(vp, W, H)
        r = V.ScanlineRasterizer(0, FLAGS | extra).initialize(None, W, H)
        r.set_stream(stream.cuda_stream); r.loadVG(sc); r.setMVP(rows)
        for _ in range(5): r.render()
        r.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(frames): r.render()
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / frames
        img = r.readback(); cnt = r.counts()
        t = time.perf_counter()
        bl = bool(extra & V.FLAG_BLEND)
        ref = O.render_aa4(sc, rows, W, H, full=True, blend=bl) if extra & V.FLAG_AA4 else O.render(sc, rows, W, H, full=True, blend=bl, keep={"rgba"})
        tor = (time.perf_counter() - t) * 1e3
        same = np.array_equal(img, ref["rgba_aa"] if extra & V.FLAG_AA4 else ref["rgba"]) and cnt["n_fragments"] == ref["n_fragments"]
        print(f"| {name} ({label}) | {W}x{H} | {sc.n_curves} | {cnt['n_fragments']} | {cnt['n_out_frag'] + cnt['n_span']} | {ms:.3f} | {W*H/ms/1e3:.0f} | "
              f"{tor:.0f} ({O.num_threads()}) | {'identical' if same else 'DIFFERENT'} | {r.sort_mode()} |")
        r.close()
