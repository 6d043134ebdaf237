#!/usr/bin/env python
"""Run the REFERENCE's shipped SPIR-V compute shaders (workdir/shaders/**/spv/*.comp.spv) through
oracle/spirv_exec.py, following the dispatch sequence, buffer sizes, bindings and push constants of
ScanlineVGRasterizer::drawFrame (VkScanlinePR/src/core/scanline/scanline_rasterizer.cpp:282-608),
and write every buffer of the frame to tests/golden/spirv_<scene>.npz.

These fixtures are outputs of the reference's own binaries (executed by an interpreter instead of
a Vulkan driver, which does not exist here): tests/test_spirv_golden.py pins the C oracle — and on
a GPU the CUDA path — to them bit for bit. Runs only where /root/reference exists (minutes of CPU).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import spirv_exec as SX  # noqa: E402
from vkscanlinepr_b200 import scene as S  # noqa: E402

REF = os.environ.get("SLPR_REFERENCE", "/root/reference")
SH = os.path.join(REF, "workdir", "shaders")
U32 = np.uint32
SLACK = 4096  # the scan shader writes a tail past n+1 (naive_scan.comp:66-69); the reference relies on robust access


def mod(rel):
    return SX.Module(os.path.join(SH, rel))


def divup(a, b):
    return (a + b - 1) // b


def run_frame(sc, rows, W, H, log=print):
    M = {k: mod(v) for k, v in dict(
        tp="scanline/compute/spv/transform_pos.comp.spv", mi0="scanline/compute/spv/make_intersection_0.comp.spv",
        mi1="scanline/compute/spv/make_intersection_1.comp.spv", gf="scanline/compute/spv/gen_fragment.comp.spv",
        shuf="scanline/compute/spv/shuffle_fragment.comp.spv",
        mark="scanline/compute/spv/mark_merged_fragment_and_span.comp.spv",
        gen="scanline/compute/spv/gen_merged_fragment_and_span.comp.spv",
        scan="common/spv/naive_scan.comp.spv", sort="common/spv/naive_seg_sort_pairs.comp.spv").items()}
    npnt, nc, P = sc.n_points, sc.n_curves, sc.n_paths
    rows = np.ascontiguousarray(rows, np.float32).reshape(16)

    def run(name, bindings, push, *grid):
        t = time.time()
        r = SX.Runner(M[name], bindings, push)
        with np.errstate(all="ignore"):
            r.dispatch(*grid)
        log(f"  {name:5s} grid={grid} instr={r.instr_count} {time.time() - t:.1f}s")

    pos = sc.pos.reshape(-1).view(U32).copy()
    pos_path, cpm, ctype, cpath = (a.astype(U32).copy() for a in (sc.pos_path, sc.curve_pos_map, sc.curve_type, sc.curve_path))
    frule, finfo = sc.fill_rule.astype(U32).copy(), sc.fill_info.astype(U32).copy()
    # TransPosIn, std140 (compute_ubo.h:9-13): n_points @0, w @4, h @8 (floats), m0..m3 @16,32,48,64
    ubo_tp = np.zeros(20, U32)
    ubo_tp[0] = npnt
    ubo_tp[1:3] = np.array([W, H], np.float32).view(U32)
    ubo_tp[4:20] = rows.view(U32)
    tpos = np.zeros(2 * npnt + 2, U32)
    pvis = np.zeros(P + 1, U32)  # zeroed per frame
    run("tp", {0: (ubo_tp, 0), 1: (pos, 0), 2: (pos_path, 0), 3: (tpos, 0), 4: (pvis, 0)}, None, divup(npnt, 256))
    # MakeInteIn (compute_ubo.h:15-18)
    ubo_mi = np.array([nc, 0, W, H], U32)
    cut = np.zeros(5 * nc + 5, U32)
    cpc = np.zeros(nc + 1 + SLACK, U32)
    run("mi0", {0: (ubo_mi, 0), 1: (ctype, 0), 2: (cpm, 0), 3: (tpos, 0), 4: (cpath, 0), 5: (pvis, 0), 6: (cut, 0), 7: (cpc, 0)},
        None, divup(nc, 256))
    count = cpc[:nc].copy()
    run("scan", {0: (cpc, 0), 1: (cpc, 0)}, np.array([nc], U32), 1)                    # SR.cpp:338-358
    nf = int(cpc[nc])
    offset = cpc[:nc + 1].copy()
    sf = (nf + 256) & -256                                                                 # SR.cpp:365
    inter = np.zeros(2 * nf + 2, U32)
    fd = np.zeros(8 * sf + 1 + SLACK, U32)
    ubo_mi[1] = nf
    run("mi1", {0: (ubo_mi, 0), 1: (inter, 0), 2: (cut, 0), 3: (cpc, 0), 4: (ctype, 0), 5: (cpm, 0), 6: (tpos, 0), 7: (cpath, 0),
                8: (pvis, 0)}, None, divup(nc, 256))                                       # SR.cpp:374-395
    run("gf", {0: (inter, 0), 1: (cpath, 0), 2: (cpm, 0), 3: (ctype, 0), 4: (tpos, 0), 5: (fd, 0)},
        np.array([P, nc, nf, sf, W, H], U32), divup(nf, 256))                              # SR.cpp:398-422
    key, idx, path, wind = (fd[k * sf:k * sf + nf].copy() for k in (0, 1, 2, 4))
    key_sentinel = int(fd[nf])
    seg = fd[3 * sf:3 * sf + P + 1].copy()
    run("sort", {0: (fd, 0), 1: (fd, sf), 2: (fd, 3 * sf)}, np.array([nf, P], U32), 256, divup(P, 256))  # SR.cpp:425-455
    skey, sidx = fd[0:nf].copy(), fd[sf:sf + nf].copy()
    run("shuf", {0: (fd, 0)}, np.array([nf, sf], U32), divup(nf, 256))                    # SR.cpp:461-476
    swind = fd[3 * sf:3 * sf + nf].copy()
    run("scan", {0: (fd, 3 * sf), 1: (fd, 3 * sf)}, np.array([nf], U32), 1)               # SR.cpp:479-506
    wn = fd[3 * sf:3 * sf + nf + 1].copy()
    run("mark", {0: (frule, 0), 1: (fd, 0)}, np.array([nf, sf, W, H], U32), divup(nf, 256))  # SR.cpp:511-529
    flags = fd[4 * sf:4 * sf + 2 * nf].copy()
    run("scan", {0: (fd, 4 * sf), 1: (fd, 6 * sf)}, np.array([2 * nf], U32), 1)           # SR.cpp:545-573
    scan3 = fd[6 * sf:6 * sf + 2 * nf + 1].copy()
    n_out = int(fd[6 * sf + nf])
    n_span = int(fd[6 * sf + 2 * nf]) - n_out                                              # SR.cpp:578-580
    out = np.zeros(4 * (n_out + n_span) + 4, U32)
    run("gen", {0: (fd, 0), 1: (finfo, 0), 2: (out, 0)}, np.array([nf, sf, W, H, n_out, n_span], U32), divup(nf, 256))
    i32 = lambda a: a.view(np.int32)
    return dict(width=W, height=H, rows=rows, n_fragments=nf, n_out_frag=n_out, n_span=n_span,
                tpos=tpos[:2 * npnt].view(np.float32).reshape(-1, 2), path_visible=i32(pvis[:P]),
                cut_cache=cut[:5 * nc].view(np.float32).reshape(-1, 5), curve_count=i32(count), curve_offset=i32(offset),
                inter=i32(inter[:2 * nf]).reshape(-1, 2), key=i32(key), key_sentinel=np.int32(key_sentinel - (1 << 32) if key_sentinel >> 31 else key_sentinel),
                idx=i32(idx), path=i32(path), wind=i32(wind), seg=i32(seg), skey=i32(skey), sidx=i32(sidx), swind=i32(swind),
                wn=i32(wn), flags=i32(flags), scan3=i32(scan3), records=i32(out[:4 * (n_out + n_span)]).reshape(-1, 4))


def scene_arrays(sc):
    return dict(pos=sc.pos, pos_path=sc.pos_path, curve_pos_map=sc.curve_pos_map, curve_type=sc.curve_type,
                curve_path=sc.curve_path, fill_rule=sc.fill_rule, fill_info=sc.fill_info)


def cases():
    import util
    yield "tiny", util.tiny_scene(), S.identity_rows(), 96, 80
    # a few glyph paths of the shipped test scene (cubics with 1-3 cuts, even-odd), zoomed to fill a small frame
    sc, vp = util.golden_scene("test")
    keep = np.isin(sc.curve_path, [1, 2, 3])
    pk = np.isin(sc.pos_path, [1, 2, 3])
    remap = {1: 0, 2: 1, 3: 2}
    sub = S.Scene(sc.pos[pk], np.array([remap[int(p)] for p in sc.pos_path[pk]], np.uint32),
                  (sc.curve_pos_map[keep] - sc.curve_pos_map[keep][0]).astype(np.uint32), sc.curve_type[keep],
                  np.array([remap[int(p)] for p in sc.curve_path[keep]], np.uint32), sc.fill_rule[[1, 2, 3]],
                  sc.fill_info[[1, 2, 3]], "test_glyphs")
    lo, hi = sub.pos.min(0), sub.pos.max(0)
    yield "glyphs", sub, S.fit_rows([lo[0] - 1, lo[1] - 1, hi[0] + 1, hi[1] + 1], 160, 96), 160, 96
    # a 4-cut cubic (exercises the MI0:340 slip), a quadric-typed curve (TODO arms) and clipping at all four edges
    yield "edge", util.edge_scene(), S.identity_rows(), 64, 48
    # the first paths of the shipped tiger (lines + cubics, nonzero and even-odd fills, unclosed sub-contours
    # that leave a winding residue), rotated and scaled so that curves cross the frame edges
    sc, vp = util.golden_scene("tiger")
    npath = 14
    ck = sc.curve_path < npath
    pk = sc.pos_path < npath
    sub = S.Scene(sc.pos[pk], sc.pos_path[pk], sc.curve_pos_map[ck], sc.curve_type[ck], sc.curve_path[ck],
                  sc.fill_rule[:npath], sc.fill_info[:npath], "tiger_head")
    rows = (S.anim_rows(23, 144, 112).astype(np.float64) @ S.fit_rows(vp, 144, 112).astype(np.float64)).astype(np.float32)
    yield "tiger14", sub, rows, 144, 112
    # synthetic blobs: cubics with 0-4 cuts incl. three, mixed fill rules
    yield "synth48", S.synth_scene(48, 128, 96, 6.0, 22.0, seed=11), S.identity_rows(), 128, 96
    # S-shaped and looping cubics, control points beyond the frame: several curves with four monotonic cuts, i.e.
    # pieces that end below their start (MI0:340) and whose first record is not their start parameter
    yield "loops", util.looping_cubics_scene(14, 96, 72, seed=5), S.identity_rows(), 96, 72
    # a projective matrix (w depends on x and y: the division of TP:53-54 and the left-to-right dot products matter)
    persp = np.array([[1.0, 0.1, 0, 2.0], [-0.05, 1.0, 0, 1.0], [0, 0, 1, 0], [0.0015, 0.002, 0, 1.0]], np.float32)
    yield "persp", S.synth_scene(20, 96, 80, 6.0, 20.0, seed=23), persp, 96, 80
    # odd frame size (the last cell column / row is half outside: clamps of GF:177-178, MARK:45) and a frame smaller
    # than the scene, so that paths are cut by every edge
    yield "odd", S.synth_scene(28, 120, 90, 6.0, 24.0, seed=31), S.identity_rows(), 97, 61


def main():
    only = sys.argv[1:] or None
    for name, sc, rows, W, H in cases():
        if only and name not in only:
            continue
        print(f"{name}: {sc.n_curves} curves {sc.n_paths} paths {W}x{H}")
        t = time.time()
        r = run_frame(sc, rows, W, H)
        print(f"  fragments={r['n_fragments']} records={r['records'].shape[0]} total {time.time() - t:.0f}s")
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"spirv_{name}.npz"), **scene_arrays(sc), **r)


if __name__ == "__main__":
    main()
