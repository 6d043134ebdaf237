#!/usr/bin/env python
"""Differential fuzzing of the CUDA path against the oracle: random small scenes built from the coordinate classes that
make the reference's arithmetic branch (exact grid lines, exact negative integers, +-0, denormals, huge values, repeated
points, cusps and loops, curves far outside the frame), random frame sizes and matrices (affine and projective), every
tap and the RGBA8 frame compared bit for bit (tests/test_parity_gpu.py assert_frame_parity), plus the long-piece walk,
the full-RVG arithmetic and the contracted-FMA policy on their own flags.
usage: python tools/fuzz_parity.py [first_seed] [count]      (a failing seed is printed and can be re-run alone)
       compute-sanitizer python tools/fuzz_parity.py 0 60 --survive     (non-finite input: memory safety only)"""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import vkscanlinepr_b200 as V  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from vkscanlinepr_b200 import scene as S  # noqa: E402


def coord(rng, W, H, nonfinite):
    """One coordinate from a class picked at random."""
    k = rng.integers(0, 12)
    span = float(max(W, H))
    if k == 0: return float(rng.integers(-4, span + 6))                    # exact integer (odd or even)
    if k == 1: return float(2 * rng.integers(-2, span // 2 + 3))           # exact even grid line
    if k == 2: return float(rng.integers(-4, span + 6)) + 0.5              # pixel centre
    if k == 3: return float(rng.choice([0.0, -0.0, 1e-45, -1e-45, 1e-38, -1e-38, 1e-30]))
    if k == 4: return float(rng.choice([-1.0, -2.0, -3.0, span, span + 1, span + 2]))
    if k == 5: return float(rng.normal(span / 2, span))                    # often outside
    if k == 6: return float(rng.choice([1e6, -1e6, 2e7, -2e7, 65536.0, 32767.0, 32768.0]))  # (beyond the int range the reference is undefined)
    if k == 7 and nonfinite: return float(rng.choice([np.inf, -np.inf, np.nan, 3.4e38, -3.4e38, 3e9, -3e9, 2.2e9]))
    return float(rng.uniform(-2.0, span + 2.0))


def random_scene(seed, nonfinite=False, full=False, big=False):
    rng = np.random.default_rng(seed)
    W = int(rng.choice([1, 2, 3, 16, 33, 64, 97, 128, 200, 320]))
    H = int(rng.choice([1, 2, 5, 16, 31, 48, 81, 128, 150, 240]))
    n_paths = int(rng.integers(1, 24))
    if big:  # few curves on a big frame: pieces of hundreds of crossings (the long-piece kernels, warp-wide bisection)
        W, H = [(1024, 768), (1920, 1080), (2049, 1537), (3840, 2160)][int(rng.integers(0, 4))]
        n_paths = int(rng.integers(1, 5))
    pos, pos_path, cpm, ctype, cpath = [], [], [], [], []
    weights = []
    for p in range(n_paths):
        n_curves = int(rng.integers(1, 9))
        cur = (coord(rng, W, H, nonfinite), coord(rng, W, H, nonfinite))
        first = cur
        for c in range(n_curves):
            r = rng.random()
            if full and r < 0.35:
                typ = int(rng.choice([S.QUADRIC, S.ARC]))
            else:
                typ = S.LINE if r < 0.45 else S.CUBIC if r < 0.93 else int(rng.choice([S.QUADRIC, S.ARC, 0x24, 0x12, 0x0B, 0x10]))  # and type values the shaders have no arm for
            npts = max(typ & 7, 1)  # the reference's convention: a type's low three bits are its point count
            pts = [cur]
            shape = rng.integers(0, 6)
            for i in range(1, npts):
                if shape == 0: q = cur                                    # repeated point / zero length
                elif shape == 1: q = (cur[0] + float(rng.integers(-3, 4)) * 2.0, cur[1])   # horizontal on the grid
                elif shape == 2: q = (cur[0], cur[1] + float(rng.integers(-3, 4)) * 2.0)   # vertical on the grid
                else: q = (coord(rng, W, H, nonfinite), coord(rng, W, H, nonfinite))
                pts.append(q)
            if c == n_curves - 1 and rng.random() < 0.7:
                pts[-1] = first                                           # closed
            cpm.append(len(pos)); ctype.append(typ); cpath.append(p)
            weights.append(float(rng.choice([1.0, 0.5, 0.70710678, 2.0, 0.0, -0.5])) if typ == S.ARC else 1.0)
            for q in pts:
                pos.append(q); pos_path.append(p)
            cur = pts[-1]
    rule = rng.integers(0, 2, n_paths).astype(np.uint32)
    col = rng.integers(0, 1 << 32, n_paths, dtype=np.uint64).astype(np.uint32)
    sc = S.Scene(np.array(pos, np.float32).reshape(-1, 2), np.array(pos_path, np.uint32), np.array(cpm, np.uint32),
                 np.array(ctype, np.uint32), np.array(cpath, np.uint32), rule, col, f"fuzz{seed}")
    if full:
        sc.curve_weight = np.array(weights, np.float32)
    m = rng.integers(0, 6)
    rows = S.identity_rows()
    if m == 1:
        rows = rows * np.float32(rng.choice([0.5, 2.0, 0.125, 7.3])); rows[2, 2] = rows[3, 3] = 1
    elif m == 2:
        a = rng.uniform(0, 6.3)
        rows[0, 0], rows[0, 1], rows[1, 0], rows[1, 1] = np.cos(a), -np.sin(a), np.sin(a), np.cos(a)
        rows[0, 3], rows[1, 3] = rng.uniform(-W, W), rng.uniform(-H, H)
    elif m == 3:
        rows[3, 0], rows[3, 1], rows[3, 3] = rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01), rng.choice([1.0, 0.5, 2.0, -1.0])
    elif m == 4:
        rows[0, 0], rows[1, 1] = rng.choice([-1.0, 1.0, 0.0]), rng.choice([-1.0, 1.0, 0.0])
        rows[0, 3], rows[1, 3] = float(W) * (rows[0, 0] < 0), float(H) * (rows[1, 1] < 0)
    return sc, rows.astype(np.float32), W, H


def check_bands(sc, rows, W, H, ref, flags):
    """Exact row bands with the device-side exchange of winding sums: three bands into one frame buffer, two frames."""
    import torch
    from vkscanlinepr_b200 import parallel as PAR
    frame = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    ctxs = []
    for y0, y1 in PAR.band_rows(H, 3):
        c = V.ScanlineRasterizer(0, flags).initialize(None, W, H)
        c.loadVG(sc); c.setMVP(rows); c.set_band(y0, y1); c.set_target(frame.data_ptr(), W * 4)
        ctxs.append(c)
    boxes = [c.band_mailbox()[0] for c in ctxs]
    for g, c in enumerate(ctxs):
        c.set_band_peers(3, g, 0, boxes)
    for seq in (1, 2):
        for c in ctxs:
            c.prepare()
        for c in ctxs:
            c.render_band(seq)
        ctxs[0].band_wait_gather(seq)
        for c in ctxs:
            c.synchronize()
        assert np.array_equal(frame.cpu().numpy(), ref), "exact bands"
    for c in ctxs:
        c.close()


def check(seed, nonfinite):
    import test_parity_gpu as T
    sc, rows, W, H = random_scene(seed, nonfinite)
    T.assert_frame_parity(sc, rows, W, H)
    # the long-piece walk (second frame on), windowed order, contracted FMA, translucent blending: frames only
    ref = O.render(sc, rows, W, H, keep={"rgba"})["rgba"]
    for flags in (V.FLAG_WINDOWED_WALK, V.FLAG_NO_LONG_WALK | V.FLAG_SEGMENTED_SORT):
        r = V.ScanlineRasterizer(0, flags).initialize(None, W, H)
        r.loadVG(sc); r.setMVP(rows)
        for _ in range(3):
            r.render()
            assert np.array_equal(r.readback(), ref), f"flags {flags}"
        r.close()
    for kw, flags in ((dict(fma=True), V.FLAG_CONTRACT_FMA), (dict(blend=True), V.FLAG_BLEND)):
        ref2 = O.render(sc, rows, W, H, keep={"rgba"}, **kw)["rgba"]
        r = V.ScanlineRasterizer(0, flags).initialize(None, W, H)
        r.loadVG(sc); r.setMVP(rows); r.render(); r.render()
        assert np.array_equal(r.readback(), ref2), f"{kw}"
        r.close()
    if seed % 4 == 0 and H >= 6:
        check_bands(sc, rows, W, H, ref, 0)
    # the full-RVG arithmetic on a scene of its own (quadratics and rational arcs with odd weights)
    sc, rows, W, H = random_scene(seed, nonfinite, full=True)
    ref3 = O.render(sc, rows, W, H, full=True)
    r = V.ScanlineRasterizer(0, V.FLAG_FULL_RVG | V.FLAG_TAPS | V.FLAG_NO_GRAPH).initialize(None, W, H)
    r.loadVG(sc); r.setMVP(rows)
    for _ in range(2):
        r.render()
        assert r.counts()["n_fragments"] == ref3["n_fragments"]
        assert np.array_equal(r.tap("intersection"), ref3["inter"]), "full: intersection tap"
        assert np.array_equal(r.readback(), ref3["rgba"]), "full: frame"
    r.close()
    if seed % 4 == 1 and H >= 6:
        check_bands(sc, rows, W, H, ref3["rgba"], V.FLAG_FULL_RVG)
    if seed % 4 == 2:  # four samples per pixel, with and without blending (defined through the reference path at 4x)
        sc, rows, W, H = random_scene(seed, nonfinite)
        for blend in (False, True):
            ref4 = O.render_aa4(sc, rows, W, H, blend=blend)["rgba_aa"]
            r = V.ScanlineRasterizer(0, V.FLAG_AA4 | (V.FLAG_BLEND if blend else 0)).initialize(None, W, H)
            r.loadVG(sc); r.setMVP(rows); r.render(); r.render()
            assert np.array_equal(r.readback(), ref4), f"aa4 blend={blend}"
            r.close()
    if seed % 4 == 3:  # a big frame with few, long curves: every tap, then the frames on which the long-piece walk is on
        for full in (False, True):
            sc, rows, W, H = random_scene(seed, nonfinite, full=full, big=True)
            refb = O.render(sc, rows, W, H, full=full)
            r = V.ScanlineRasterizer(0, V.FLAG_TAPS | V.FLAG_NO_GRAPH | (V.FLAG_FULL_RVG if full else 0)).initialize(None, W, H)
            r.loadVG(sc); r.setMVP(rows)
            for _ in range(2):
                r.render()
                assert np.array_equal(r.tap("intersection"), refb["inter"]), f"big full={full}: intersection tap"
                assert np.array_equal(r.tap("sorted_key"), refb["skey"]), f"big full={full}: sorted keys"
                assert np.array_equal(r.readback(), refb["rgba"]), f"big full={full}: frame"
            r.close()
            g = V.ScanlineRasterizer(0, V.FLAG_FULL_RVG if full else 0).initialize(None, W, H)
            g.loadVG(sc); g.setMVP(rows)
            for _ in range(3):
                g.render()
                assert np.array_equal(g.readback(), refb["rgba"]), f"big full={full}: graph frames"
            g.close()


def survive(seed):
    """Outside the reference's domain (non-finite and near-overflow coordinates: its int() conversions are undefined there
    and the oracle defines nothing): the library must stay memory-safe and return — run under compute-sanitizer."""
    for full in (False, True):
        sc, rows, W, H = random_scene(seed, True, full=full)
        for flags in (0, V.FLAG_RADIX_SORT | V.FLAG_SEPARATE_FILL, V.FLAG_BLEND | V.FLAG_AA4):
            r = V.ScanlineRasterizer(0, flags | (V.FLAG_FULL_RVG if full else 0)).initialize(None, W, H)
            r.loadVG(sc); r.setMVP(rows)
            for _ in range(2):
                r.render()
                assert r.readback().shape == (H, W, 4)
            r.close()


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    first = int(args[0]) if args else 0
    count = int(args[1]) if len(args) > 1 else 100
    nonfinite = "--nonfinite" in sys.argv
    bad = []
    for seed in range(first, first + count):
        try:
            if "--survive" in sys.argv:
                survive(seed)
            else:
                check(seed, nonfinite)
        except Exception as e:  # noqa: BLE001
            bad.append(seed)
            print(f"seed {seed}: {type(e).__name__}: {str(e)[:200]}")
            if os.environ.get("FUZZ_TRACE"):
                traceback.print_exc()
    what = "survived" if "--survive" in sys.argv else "identical to the oracle"
    print(f"fuzz: {count - len(bad)} of {count} seeds {what}" + (f"; failing: {bad}" if bad else ""))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
