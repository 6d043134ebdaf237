"""GPU parity tests: the CUDA path (through the C ABI) against the oracle, stage by stage.

Bar: bit-exact for every integer / index / bit-pattern buffer (intersection counts, t parameters,
sorted fragment order, winding numbers, draw records) and identical RGBA8 (tolerance 0 LSB; the
north star allows 1 LSB, PSNR is reported as inf).
"""
import numpy as np
import pytest

import util
import vkscanlinepr_b200 as V
from oracle import oracle_py as O
from vkscanlinepr_b200 import scene as S

pytestmark = pytest.mark.gpu

INT_TAPS = ["path_visible", "curve_count", "curve_offset", "intersection", "path", "winding", "segments",
            "sorted_key", "sorted_index", "winding_scan", "flags", "flag_scan", "records"]
ORACLE_NAME = dict(path_visible="path_visible", curve_count="curve_count", curve_offset="curve_offset",
                   intersection="inter", path="path", winding="wind", segments="seg", sorted_key="skey",
                   sorted_index="sidx", winding_scan="wn", flags="flags", flag_scan="scan3", records="records")


def render_gpu(sc, rows, W, H, flags):
    r = V.ScanlineRasterizer(0, flags).initialize(None, W, H)
    r.loadVG(sc)
    r.setMVP(rows)
    r.render()
    return r


def assert_frame_parity(sc, rows, W, H):
    ref = O.render(sc, rows, W, H)
    r = render_gpu(sc, rows, W, H, V.FLAG_TAPS | V.FLAG_NO_GRAPH)
    cnt = r.counts()
    assert cnt == {k: ref[k] for k in ("n_fragments", "n_out_frag", "n_span")}
    assert np.array_equal(r.tap("transformed_pos").view(np.uint32), ref["tpos"].view(np.uint32)), "transformed_pos"
    assert np.array_equal(r.tap("cut_cache").view(np.uint32), ref["cut_cache"].view(np.uint32)), "cut_cache"
    key = r.tap("key")
    assert np.array_equal(key[:-1], ref["key"]) and (ref["n_fragments"] == 0 or key[-1] == -1), "key plane"
    for t in INT_TAPS:
        got, exp = r.tap(t), ref[ORACLE_NAME[t]]
        assert got.shape == exp.shape and np.array_equal(got, exp), f"tap {t} differs"
    img = r.readback()
    assert np.array_equal(img, ref["rgba"]), f"RGBA differs, PSNR {util.psnr(img, ref['rgba']):.2f} dB"
    r.close()
    # fast path: no taps, CUDA graph replay, twice (the second frame reuses the captured graph)
    f = render_gpu(sc, rows, W, H, 0)
    assert np.array_equal(f.readback(), ref["rgba"])
    f.render()
    assert np.array_equal(f.readback(), ref["rgba"])
    assert f.counts() == cnt
    with pytest.raises(V.SlprError):  # the fast path does not materialise the draw records unless asked to
        f.tap("records")
    f.close()
    # the other sort: the default picks the segmented sort unless a path is too long for it
    g = render_gpu(sc, rows, W, H, V.FLAG_RADIX_SORT | V.FLAG_TAPS)
    assert g.counts() == cnt and g.sort_mode() == "radix"
    for t in ("sorted_key", "sorted_index", "records"):
        assert np.array_equal(g.tap(t), ref[ORACLE_NAME[t]]), f"radix sort: tap {t} differs"
    assert np.array_equal(g.readback(), ref["rgba"])
    g.close()
    # stage-5 coverage both ways: marked by the span kernel itself (big frames) or by a separate pass (small ones)
    # (fused: cells carry the highest path, records only on request; separate: cells carry the last record)
    for flag, fused in ((V.FLAG_FUSED_FILL, True), (V.FLAG_SEPARATE_FILL, False)):
        for rec in (0, V.FLAG_RECORDS):
            h = render_gpu(sc, rows, W, H, flag | rec)
            assert np.array_equal(h.readback(), ref["rgba"]), f"fill fused={fused} records={bool(rec)}"
            assert h.fill_fused() == fused
            if rec:
                assert np.array_equal(h.tap("records"), ref["records"])
            h.close()
    return ref


@pytest.mark.parametrize("name", util.SHIPPED)
@pytest.mark.parametrize("size", [(1024, 1024), (1920, 1080)])
def test_shipped_scene_parity(name, size):
    sc, vp = util.golden_scene(name)
    W, H = size
    assert_frame_parity(sc, S.fit_rows(vp, W, H), W, H)


def test_paper_config_all_nonzero():
    """BASELINE cfg1: test.rvg stands in for the missing paper-1.rvg, 1024x1024, uniform fit without
    centring, once with the file's fill rules and once forced to nonzero."""
    sc, vp = util.golden_scene("test")
    rows = S.fit_rows(vp, 1024, 1024, centred=False)
    assert_frame_parity(sc, rows, 1024, 1024)
    assert_frame_parity(sc.with_fill_rule(S.NON_ZERO), rows, 1024, 1024)


@pytest.mark.parametrize("name", util.EMPTY)
def test_empty_scenes_render_white(name):
    sc, vp = util.golden_scene(name)
    ref = assert_frame_parity(sc, S.identity_rows(), 320, 200)
    assert np.all(ref["rgba"] == 255)


def test_tiny_scene_and_odd_sizes():
    sc = util.tiny_scene()
    for W, H in [(96, 80), (97, 81), (33, 17), (2, 2), (1, 1)]:
        assert_frame_parity(sc, S.identity_rows(), W, H)


def test_zoomed_and_offscreen_views():
    sc, vp = util.golden_scene("tiger")
    W, H = 800, 600
    zoom = S.fit_rows(vp, W, H)
    zoom[0, 0] *= 9; zoom[1, 1] *= 9; zoom[0, 3] = -2400; zoom[1, 3] = -1800   # deep zoom: long curves, clipping
    assert_frame_parity(sc, zoom, W, H)
    away = S.fit_rows(vp, W, H); away[0, 3] += 5000                              # everything off-screen
    ref = assert_frame_parity(sc, away, W, H)
    assert ref["n_fragments"] == 0
    rot = S.anim_rows(37, W, H)                                                   # rotation + scale (cfg5 matrix)
    assert_frame_parity(sc, (rot.astype(np.float64) @ S.fit_rows(vp, W, H).astype(np.float64)).astype(np.float32), W, H)


def test_reference_hardcoded_matrix():
    """The zoom matrix the reference hard-codes for frame 0 (scanline_rasterizer.cpp:1162-1165)."""
    sc, _ = util.golden_scene("test")
    rows = np.eye(4, dtype=np.float32)
    rows[0, 0] = rows[1, 1] = np.float32(13.190648); rows[0, 3] = np.float32(-1142.983887); rows[1, 3] = np.float32(-6987.709961)
    assert_frame_parity(sc, rows, 1200, 1024)


def test_synthetic_scene_parity():
    sc = S.synth_scene(4096, 1024, 768, 6.0, 30.0, seed=0x5CA71E01)
    assert_frame_parity(sc, S.identity_rows(), 1024, 768)


def test_windowed_walk_order():
    """The walk's piece order is a scheduling choice (csrc/geom.cuh PieceLayout): windows of consecutive curves
    (the default on scenes of more than two million curves, forced here) must give the same buffers."""
    for sc, W, H in ((S.synth_scene(4096, 1024, 768, 6.0, 30.0, seed=0x5CA71E01), 1024, 768),
                     (util.looping_cubics_scene(), 512, 384), (util.golden_scene("tiger")[0], 640, 480)):
        rows = S.identity_rows() if sc.name != "tiger" else S.fit_rows(util.golden_scene("tiger")[1], W, H)
        ref = O.render(sc, rows, W, H)
        for flags in (V.FLAG_WINDOWED_WALK | V.FLAG_TAPS | V.FLAG_NO_GRAPH, V.FLAG_WINDOWED_WALK | V.FLAG_RECORDS):
            r = render_gpu(sc, rows, W, H, flags)
            assert r.counts() == {k: ref[k] for k in ("n_fragments", "n_out_frag", "n_span")}
            if flags & V.FLAG_TAPS:
                for t in ("intersection", "path", "winding", "sorted_key", "sorted_index", "winding_scan"):
                    assert np.array_equal(r.tap(t), ref[ORACLE_NAME[t]]), f"tap {t} differs"
            assert np.array_equal(r.tap("records"), ref["records"])
            assert np.array_equal(r.readback(), ref["rgba"])
            r.close()


def test_cubics_with_cuts_out_of_order():
    """Hundreds of cubics with four monotonic cuts (MI0:340 leaves the fourth unsorted): pieces that end below
    their start, whose boundary fragments k_piece_fix re-emits from the parameters k_walk leaves for it."""
    sc = util.looping_cubics_scene()
    ref = assert_frame_parity(sc, S.identity_rows(), 512, 384)
    ncuts = ref["cut_cache"].reshape(-1, 5)[:, 4].copy().view(np.uint32)
    assert (ncuts == 4).sum() >= 100
    assert_frame_parity(sc, S.anim_rows(29, 512, 384), 512, 384)


def test_sort_modes_by_path_size():
    """Segmented sort: warp network (paths up to 512 fragments), block network (up to 4096), and the
    switch to the radix sort when a path is longer than that or when the host's cost model prefers it.
    Sorted order must equal the oracle's in every mode (naive_seg_sort_pairs.comp:26-97: signed
    (key, index) order inside each path)."""
    W = H = 1024
    small = S.synth_scene(3000, W, H, 10.0, 40.0, seed=0x5E650001)     # glyph-sized paths
    medium = S.synth_scene(300, W, H, 150.0, 400.0, seed=0x5E650002)   # hundreds .. thousands per path
    sc, vp = util.golden_scene("tiger")
    cases = ((small, S.identity_rows(), V.FLAG_SEGMENTED_SORT, "segmented"),
             (medium, S.identity_rows(), V.FLAG_SEGMENTED_SORT, "segmented"),  # exercises k_segsort_block
             (small, S.identity_rows(), 0, "segmented"),                        # many small paths: the model keeps it
             (medium, S.identity_rows(), 0, "radix"),                           # few long paths: the model leaves it
             (sc, S.fit_rows(vp, W, H), 0, "radix"))                            # a path of more than 4096 fragments
    for scn, rows, flags, want in cases:
        ref = O.render(scn, rows, W, H)
        r = render_gpu(scn, rows, W, H, V.FLAG_TAPS | flags)
        for frame in range(2):  # the second frame runs in the mode chosen after the first
            assert r.counts()["n_fragments"] == ref["n_fragments"]  # completes the frame (and a possible fallback)
            assert np.array_equal(r.tap("sorted_key"), ref["skey"]) and np.array_equal(r.tap("sorted_index"), ref["sidx"])
            assert np.array_equal(r.tap("records"), ref["records"])
            assert np.array_equal(r.readback(), ref["rgba"])
            r.render()
        r.synchronize()
        assert r.sort_mode() == want, (int(np.max(np.diff(ref["seg"]))), r.sort_mode(), want)
        r.close()
    # a scene with one path longer than 4096 fragments: first frame falls back, later frames stay on radix
    big = S.synth_scene(6, 4096, 4096, 1800.0, 2000.0, seed=0x5E650003)
    ref = O.render(big, S.identity_rows(), 4096, 4096)
    assert int(np.max(np.diff(ref["seg"]))) > 4096
    r = render_gpu(big, S.identity_rows(), 4096, 4096, V.FLAG_SEGMENTED_SORT | V.FLAG_RECORDS)
    assert r.counts()["n_fragments"] == ref["n_fragments"]
    assert r.sort_mode() == "radix"
    assert np.array_equal(r.readback(), ref["rgba"]) and np.array_equal(r.tap("records"), ref["records"])
    r.render()
    assert np.array_equal(r.readback(), ref["rgba"])
    r.close()


def test_mvp_changes_between_frames_and_capacity_growth():
    """One context, several matrices: the graph is replayed with new FrameParams; a zoom that makes
    far more fragments than the first frame forces the capacity to grow and the frame to be redone."""
    sc, vp = util.golden_scene("tiger")
    W, H = 640, 480
    r = V.ScanlineRasterizer(0, 0).initialize(None, W, H)
    r.loadVG(sc)
    small = S.fit_rows(vp, W, H); small[0, 0] *= 0.05; small[1, 1] *= 0.05
    for rows in (small, S.fit_rows(vp, W, H), S.anim_rows(11, W, H) @ S.fit_rows(vp, W, H), small):
        rows = np.asarray(rows, np.float32)
        r.setMVP(rows)
        r.render()
        ref = O.render(sc, rows, W, H)
        assert np.array_equal(r.readback(), ref["rgba"])
        assert r.counts()["n_fragments"] == ref["n_fragments"]
    r.close()


def test_row_bands_equal_full_frame():
    """rendering even-aligned row bands independently and stacking them equals the full
    frame (exact for scenes whose per-(path,row) winding sums are zero: closed paths)."""
    W = H = 512
    rows = S.identity_rows()
    for seed in range(3, 40):  # first seed without a 4-cut cubic (MI0:340 slip leaves a winding residue)
        sc = S.synth_scene(1024, 512, 512, 6.0, 40.0, seed=seed)
        ref = O.render(sc, rows, W, H)
        res = np.zeros(sc.n_paths, np.int64)
        np.add.at(res, ref["path"], ref["wind"])
        if not res.any():
            break
    assert ref["wn"][-1] == 0 and not res.any()
    full = np.zeros((H, W, 4), np.uint8)
    for G in (2, 4):
        for g in range(G):
            y0, y1 = g * H // G, (g + 1) * H // G
            r = V.ScanlineRasterizer(0, 0).initialize(None, W, H)
            r.loadVG(sc); r.setMVP(rows); r.set_band(y0, y1); r.render()
            img = r.readback()
            full[H - y1:H - y0] = img[H - y1:H - y0]   # image row = H-1-scanline row
            assert r.counts()["n_fragments"] <= ref["n_fragments"]
            r.close()
        assert np.array_equal(full, ref["rgba"]), f"G={G}"
    # the same scene with its points stored curve by curve in reverse order: points are no longer grouped by path, so
    # band mode falls back from the per-path pass (k_band_paths) to the per-point / per-curve passes
    order = np.concatenate([np.arange(m, m + (t & 7)) for m, t in zip(sc.curve_pos_map[::-1], sc.curve_type[::-1])])
    new_map = np.zeros(sc.n_curves, np.uint32)
    new_map[::-1] = np.concatenate([[0], np.cumsum((sc.curve_type[::-1] & 7))[:-1]]).astype(np.uint32)
    shuffled = S.Scene(sc.pos[order], sc.pos_path[order], new_map, sc.curve_type, sc.curve_path, sc.fill_rule, sc.fill_info, "shuffled")
    assert (np.diff(shuffled.pos_path.astype(np.int64)) < 0).any()
    for g in range(4):
        y0, y1 = g * H // 4, (g + 1) * H // 4
        r = V.ScanlineRasterizer(0, 0).initialize(None, W, H)
        r.loadVG(shuffled); r.setMVP(rows); r.set_band(y0, y1); r.render()
        full[H - y1:H - y0] = r.readback()[H - y1:H - y0]
        r.close()
    assert np.array_equal(full, ref["rgba"]), "points not grouped by path"


def test_exact_row_bands_with_exchange():
    """SURVEY §8e with a winding residue: G band contexts (one GPU here, one per GPU in production) render
    the two halves of the frame around an all-gather of their per-path winding sums (csrc/bands.cuh); the
    assembled frame must be the full frame, bit for bit, on scenes where independent bands are NOT exact."""
    import torch
    from vkscanlinepr_b200 import parallel as PAR
    tig, vp = util.golden_scene("tiger")
    cases = [(S.synth_scene(2000, 512, 512, 6.0, 30.0, seed=0x5CA71E01), S.identity_rows(), 512, 512, 3),
             (S.synth_scene(3000, 640, 360, 4.0, 60.0, seed=0x5CA71E02), S.identity_rows(), 640, 360, 8),
             (tig, S.fit_rows(vp, 640, 480), 640, 480, 4)]
    for sc, rows, W, H, G in cases:
        ref = O.render(sc, rows, W, H)
        P = sc.n_paths
        bands = PAR.band_rows(H, G)
        frame = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
        gathered = torch.zeros((G, 3 * P), dtype=torch.int32, device="cuda")
        sums = [torch.zeros(3 * P, dtype=torch.int32, device="cuda") for _ in range(G)]
        ctxs = []
        for g, (y0, y1) in enumerate(bands):
            c = V.ScanlineRasterizer(0, 0).initialize(None, W, H)
            c.loadVG(sc)
            c.setMVP(rows)
            c.set_band(y0, y1)
            c.set_target(frame.data_ptr(), W * 4)
            assert c.band_exchange_ints() == 3 * P
            c.set_band_exchange(sums[g].data_ptr(), gathered.data_ptr(), G, g)
            ctxs.append(c)
        with pytest.raises(RuntimeError):
            ctxs[0].render()  # a configured exchange needs the two-step call
        for _ in range(2):  # the second frame replays the captured graphs
            frame.zero_()
            torch.cuda.synchronize()
            for c in ctxs:
                c.render_band_begin()  # returns with the band's sums complete
            gathered.copy_(torch.stack(sums))  # stands in for the NCCL all-gather
            torch.cuda.synchronize()
            for c in ctxs:
                c.render_band_end()
            for c in ctxs:
                c.synchronize()
            got = frame.cpu().numpy()
            assert np.array_equal(got, ref["rgba"]), f"{int((got != ref['rgba']).any(axis=2).sum())} pixels differ ({G} bands)"
        if ref["wn"][-1] != 0:  # with a residue, independent bands are expected to differ: the exchange is what fixes it
            ctxs[G - 1].set_band_exchange(0, 0, 0, 0)
            ctxs[G - 1].render()
            ctxs[G - 1].synchronize()
        for c in ctxs:
            c.close()


def test_error_paths():
    r = V.ScanlineRasterizer(0, 0).initialize(None, 64, 64)
    with pytest.raises(V.SlprError):
        r.render()                      # no scene
    with pytest.raises(V.SlprError):
        r.set_band(1, 64)               # odd band start
    sc = util.tiny_scene()
    bad = S.Scene(sc.pos, sc.pos_path, sc.curve_pos_map, sc.curve_type, sc.curve_path + 7, sc.fill_rule, sc.fill_info)
    with pytest.raises(V.SlprError):
        r.loadVG(bad)                   # path index out of range
    r.close()
    with pytest.raises(V.SlprError):
        V.ScanlineRasterizer(0, 0).initialize(None, 40000, 64)


def test_pipelined_submit_to_host():
    """slpr_submit_to_host / slpr_wait_host: two framebuffers, copies on a second stream; every frame of a
    sequence with changing matrices must land in its host buffer intact."""
    import torch
    sc, vp = util.golden_scene("tiger")
    W, H = 512, 384
    mats = [np.asarray(S.anim_rows(f, W, H) @ S.fit_rows(vp, W, H), np.float32) for f in (0, 5, 9, 14, 20)]
    refs = [O.render(sc, m, W, H)["rgba"] for m in mats]
    r = V.ScanlineRasterizer(0, 0).initialize(None, W, H)
    r.loadVG(sc)
    bufs = [torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy() for _ in mats]
    for m, b in zip(mats, bufs):
        r.submit_to_host(m, b)
    r.wait_host()
    for i, (b, ref) in enumerate(zip(bufs, refs)):
        assert np.array_equal(b, ref), f"frame {i}"
    r.close()


def test_pipelined_animation_outgrows_buffers():
    """cfg5-style sequence through slpr_submit_to_host whose fragment count grows several-fold between frames
    (zoom out from a few blobs to the whole synthetic scene): frames that outgrow the buffers sized from earlier
    frames are found from their device counters and rendered again — every frame must still land in its host
    buffer identical to the oracle's."""
    import torch
    W, H = 1024, 768
    sc = S.synth_scene(4096, W, H, 6.0, 30.0)

    def zoom(s):
        m = np.eye(4, dtype=np.float32)
        m[0, 0] = m[1, 1] = s
        m[0, 3] = W * 0.5 * (1 - s)
        m[1, 3] = H * 0.5 * (1 - s)
        return m

    mats = [zoom(s) for s in (0.05, 0.06, 1.0, 1.0, 0.4, 1.6, 1.0, 0.05, 1.3)]
    refs = [O.render(sc, m, W, H) for m in mats]
    nf = [r["n_fragments"] for r in refs]
    assert max(nf) > 1.25 * nf[0] + 65536 + 16, nf  # beyond the head-room the first frame leaves (slpr.cu)
    r = V.ScanlineRasterizer(0, 0).initialize(None, W, H)
    r.loadVG(sc)
    bufs = [torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy() for _ in mats]
    for m, b in zip(mats, bufs):
        r.submit_to_host(m, b)
    r.wait_host()
    for i, (b, ref) in enumerate(zip(bufs, refs)):
        assert np.array_equal(b, ref["rgba"]), f"frame {i}"
    assert r.pipeline_redone() >= 1
    r.close()


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json's full sizes against the oracle (cfg2 at 4K, cfg4 and cfg5 had no oracle comparison
# under pytest). The oracle keeps only what is compared (counts, draw records, RGBA8, sorted order).
def assert_big_frame(sc, rows, W, H, flags=0, keep=("records", "rgba")):
    ref = O.render(sc, rows, W, H, keep=set(keep))
    r = render_gpu(sc, rows, W, H, flags | V.FLAG_RECORDS)
    cnt = r.counts()
    assert cnt == {k: ref[k] for k in ("n_fragments", "n_out_frag", "n_span")}, (cnt, ref["n_fragments"])
    got = r.tap("records")
    assert got.shape == ref["records"].shape and np.array_equal(got, ref["records"]), "draw records differ"
    del got
    img = r.readback()
    assert np.array_equal(img, ref["rgba"]), f"{int((img != ref['rgba']).any(axis=2).sum())} pixels differ"
    r.close()
    f = render_gpu(sc, rows, W, H, flags)  # the default path: no records, graph replay, second frame too
    f.render()
    assert np.array_equal(f.readback(), ref["rgba"])
    info = dict(sort=f.sort_mode(), fused=f.fill_fused(), n_fragments=cnt["n_fragments"])
    f.close()
    return ref, info


@pytest.mark.parametrize("name", util.SHIPPED)
def test_shipped_scene_parity_4k(name):
    """BASELINE cfg2 at 3840x2160: every shipped scene the reference's parser accepts, frame identical to the oracle."""
    sc, vp = util.golden_scene(name)
    assert_big_frame(sc, S.fit_rows(vp, 3840, 2160), 3840, 2160)


@pytest.mark.parametrize("frame", [0, 37, 128, 255])
def test_cfg5_animation_frames_of_synth_1m_4k(frame):
    """BASELINE cfg5: frames of the 256-frame animation of synth_1m_4k (rotation about the centre, scale 1 +- 0.5)."""
    sc = S.synth_1m_4k()
    _, info = assert_big_frame(sc, S.anim_rows(frame, 3840, 2160), 3840, 2160)
    assert info["n_fragments"] > 5_000_000


def test_default_windowed_walk_beyond_two_million_curves():
    """More than 2 M curves: the walk's piece layout switches to windows of consecutive curves by itself (slpr.cu,
    WALK_WINDOWED_CURVES; VERDICT r1 only had the forced flag on 4 096-path scenes). Every tap of the front half
    and the sorted order against the oracle, then records and pixels."""
    sc = S.synth_scene(560_000, 4096, 4096, 3.0, 9.0, seed=0x5CA71E03, name="synth_2m2")
    assert sc.n_curves > 2_000_000
    W = H = 4096
    ref = O.render(sc, S.identity_rows(), W, H, keep={"inter", "path", "wind", "skey", "sidx", "wn", "records", "rgba"})
    r = render_gpu(sc, S.identity_rows(), W, H, V.FLAG_TAPS)
    assert r.counts() == {k: ref[k] for k in ("n_fragments", "n_out_frag", "n_span")}
    for t in ("intersection", "path", "winding", "sorted_key", "sorted_index", "winding_scan", "records"):
        assert np.array_equal(r.tap(t), ref[ORACLE_NAME[t]]), f"tap {t} differs"
    assert np.array_equal(r.readback(), ref["rgba"])
    r.close()
    f = render_gpu(sc, S.identity_rows(), W, H, 0)
    assert np.array_equal(f.readback(), ref["rgba"])
    f.close()


def test_cfg4_synth_16k_full_frame_vs_oracle():
    """BASELINE cfg4 on one GPU: the 16384 x 16384 frame of 4 194 304 curves (133 M fragments: the default windowed
    walk, the 8- and 16-register sort networks, 48-bit keys) against the oracle — counts, all 146 M draw records,
    all 268 M pixels."""
    sc = S.synth_16k()
    ref, info = assert_big_frame(sc, S.identity_rows(), 16384, 16384)
    assert info["n_fragments"] > 100_000_000 and info["sort"] == "segmented" and info["fused"]


def test_long_path_with_default_flags_first_frame():
    """a path of more than 4096 fragments on a context with DEFAULT flags — the first frame starts with
    the segmented sort, which gives up; the back half must not run on the half-sorted buffers (frame_void), and the
    host renders the frame again with the radix sort. A full-frame rectangle at 4K plus small blobs."""
    W, H = 3840, 2160
    base = S.synth_scene(2000, W, H, 6.0, 30.0, seed=0x5CA71E04)
    rect = [(1.5, 1.25), (W - 2.5, 1.25), (W - 2.5, H - 2.75), (1.5, H - 2.75)]
    pos = [p for i in range(4) for p in (rect[i], rect[(i + 1) % 4])]
    P = base.n_paths
    sc = S.Scene(np.concatenate([np.array(pos, np.float32), base.pos]),
                 np.concatenate([np.zeros(8, np.uint32), base.pos_path + 1]),
                 np.concatenate([np.arange(4, dtype=np.uint32) * 2, base.curve_pos_map + 8]),
                 np.concatenate([np.full(4, S.LINE, np.uint32), base.curve_type]),
                 np.concatenate([np.zeros(4, np.uint32), base.curve_path + 1]),
                 np.concatenate([[0], base.fill_rule]).astype(np.uint32),
                 np.concatenate([[0xFF203040], base.fill_info]).astype(np.uint32), "rect_plus_blobs")
    assert sc.n_paths == P + 1
    ref = O.render(sc, S.identity_rows(), W, H, keep={"seg", "rgba"})
    assert int(np.max(np.diff(ref["seg"]))) > 4096
    for flags in (0, V.FLAG_NO_GRAPH):
        r = render_gpu(sc, S.identity_rows(), W, H, flags)
        assert np.array_equal(r.readback(), ref["rgba"])
        assert r.sort_mode() == "radix"
        r.render()
        assert np.array_equal(r.readback(), ref["rgba"])
        r.close()
    # and through the pipelined host path, whose first frame also starts segmented
    import torch
    r = V.ScanlineRasterizer(0, 0).initialize(None, W, H)
    r.loadVG(sc)
    bufs = [torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(3)]
    for b in bufs:
        r.submit_to_host(S.identity_rows(), b)
    r.wait_host()
    for b in bufs:
        assert np.array_equal(b, ref["rgba"])
    r.close()


def _render_peer_bands(ctxs, seq0, root_waits=True, max_tries=4):
    """One frame on all band contexts with the device-side exchange; handles SLPR_ERR_RETRY the way parallel.py does
    across ranks (everybody renders the frame again with a new frame_seq). Returns the next unused seq."""
    seq = seq0
    for _ in range(max_tries):
        for c in ctxs:   # contexts of ONE process: nothing that blocks the device (sizing, allocation, graph capture) may
            c.prepare()  # happen while another context's exchange kernel waits for this one (slpr_prepare)
        for c in ctxs:
            c.render_band(seq)
        if root_waits:
            ctxs[0].band_wait_gather(seq)
        retry = False
        for c in ctxs:
            try:
                c.synchronize()
            except V.SlprRetry:
                retry = True
        seq += 1
        if not retry:
            return seq
    raise AssertionError("the band frame stayed void")


def test_exact_row_bands_device_side_exchange():
    """Round 2: the same exactness as test_exact_row_bands_with_exchange with NO host hand-over — sparse per-path sums
    stored into the other bands' mailboxes, flags, a merged break-point table (csrc/bands.cuh second half). G band
    contexts on one GPU stand in for G GPUs (their mailboxes are plain pointers here, CUDA IPC mappings in
    production: tests/test_ipc_gpu.py); they all render into ONE frame buffer, like bands storing into the root's."""
    import torch
    from vkscanlinepr_b200 import parallel as PAR
    tig, vp = util.golden_scene("tiger")
    cases = [(S.synth_scene(2000, 512, 512, 6.0, 30.0, seed=0x5CA71E01), S.identity_rows(), 512, 512, 3),
             (S.synth_scene(3000, 640, 360, 4.0, 60.0, seed=0x5CA71E02), S.identity_rows(), 640, 360, 8),
             (util.looping_cubics_scene(), S.identity_rows(), 512, 384, 4),      # hundreds of residue paths
             (tig, S.fit_rows(vp, 640, 480), 640, 480, 4)]                       # a path too long for the segmented sort: retry
    for sc, rows, W, H, G in cases:
        ref = O.render(sc, rows, W, H, keep={"rgba", "wn"})
        bands = PAR.band_rows(H, G)
        frame = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
        ctxs = []
        for g, (y0, y1) in enumerate(bands):
            c = V.ScanlineRasterizer(0, 0).initialize(None, W, H)
            c.loadVG(sc)
            c.setMVP(rows)
            c.set_band(y0, y1)
            c.set_target(frame.data_ptr(), W * 4)
            ctxs.append(c)
        boxes = [c.band_mailbox()[0] for c in ctxs]
        for g, c in enumerate(ctxs):
            c.set_band_peers(G, g, 0, boxes)
        seq = 7
        for _ in range(3):  # later frames replay the captured graph; mailbox slots alternate
            frame.zero_()
            torch.cuda.synchronize()
            seq = _render_peer_bands(ctxs, seq)
            got = frame.cpu().numpy()
            assert np.array_equal(got, ref["rgba"]), f"{sc.name}: {int((got != ref['rgba']).any(axis=2).sum())} pixels differ ({G} bands)"
        for c in ctxs:
            c.close()


def test_long_pieces_walked_chain_by_chain():
    """Round 2 (csrc/walk.cuh): on frames with few long monotone pieces the context switches, after the first frame, to
    walking them as two independent chains (x crossings, y crossings; lines all at once) merged by rank. Every tap of
    the SECOND frame — intersection records with their tag bits, fragments, sorted order, draw records, pixels — must
    still equal the oracle's: shipped scenes at large frames, cubics whose cuts are out of order, a deep zoom, and the
    full-RVG arcs."""
    tig, vp = util.golden_scene("tiger")
    tst, tvp = util.golden_scene("test")
    car, cvp = util.full_golden_scene("car")
    zoom = S.fit_rows(vp, 1600, 1200)
    zoom[0, 0] *= 6; zoom[1, 1] *= 6; zoom[0, 3] = -3000; zoom[1, 3] = -2500
    cases = [(tig, S.fit_rows(vp, 2560, 1440), 2560, 1440, 0), (tst, S.fit_rows(tvp, 3840, 2160), 3840, 2160, 0),
             (util.looping_cubics_scene(), S.identity_rows() * np.float32(4) + np.diag([0, 0, -3, -3]).astype(np.float32), 2048, 1536, 0),
             (tig, zoom, 1600, 1200, 0), (car, S.fit_rows(cvp, 1800, 1200), 1800, 1200, V.FLAG_FULL_RVG),
             (S.synth_scene(40, 2048, 2048, 300.0, 900.0, seed=0x5E650009), S.identity_rows(), 2048, 2048, 0)]
    for sc, rows, W, H, extra in cases:
        ref = O.render(sc, rows, W, H, full=bool(extra))
        r = render_gpu(sc, rows, W, H, V.FLAG_TAPS | V.FLAG_NO_GRAPH | extra)
        assert r.counts()["n_fragments"] == ref["n_fragments"]
        on, n_long = r.long_walk_info()
        assert on and n_long > 0, (sc.name, on, n_long)
        r.render()  # this frame uses the long-piece walk
        assert r.counts() == {k: ref[k] for k in ("n_fragments", "n_out_frag", "n_span")}
        for t in INT_TAPS:
            assert np.array_equal(r.tap(t), ref[ORACLE_NAME[t]]), f"{sc.name}: tap {t} differs"
        key = r.tap("key")
        assert np.array_equal(key[:-1], ref["key"])
        assert np.array_equal(r.readback(), ref["rgba"])
        r.close()
        g = render_gpu(sc, rows, W, H, extra)          # graph replay: frames 2 and 3 in long mode
        for _ in range(3):
            assert np.array_equal(g.readback(), ref["rgba"])
            g.render()
        assert g.long_walk_info()[0]
        g.close()
        h = render_gpu(sc, rows, W, H, extra | V.FLAG_NO_LONG_WALK)
        h.render()
        assert np.array_equal(h.readback(), ref["rgba"]) and not h.long_walk_info()[0]
        h.close()


def test_contracted_fma_mode():
    """SURVEY App. D.1 / SLPR_FLAG_CONTRACT_FMA: the other legitimate reading of the shaders (a*b + c fused, as a Vulkan
    driver's compiler may). Nothing of the reference's pins it; the policy is defined in the oracle
    (orc_set_contract_fma) and the CUDA path must reproduce it bit for bit, tap by tap — and differ from the default
    reading only by rounding (a handful of fragments, no visible change)."""
    tig, vp = util.golden_scene("tiger")
    cases = [(tig, S.fit_rows(vp, 1024, 768), 1024, 768, False), (S.synth_scene(4096, 1024, 768, 6.0, 30.0), S.identity_rows(), 1024, 768, False),
             (util.looping_cubics_scene(), S.anim_rows(29, 512, 384), 512, 384, False), (util.quad_arc_scene(300, 640, 480), S.identity_rows(), 640, 480, True)]
    for sc, rows, W, H, full in cases:
        ref = O.render(sc, rows, W, H, full=full, fma=True)
        extra = V.FLAG_CONTRACT_FMA | (V.FLAG_FULL_RVG if full else 0)
        r = render_gpu(sc, rows, W, H, V.FLAG_TAPS | V.FLAG_NO_GRAPH | extra)
        assert r.counts() == {k: ref[k] for k in ("n_fragments", "n_out_frag", "n_span")}
        assert np.array_equal(r.tap("transformed_pos").view(np.uint32), ref["tpos"].view(np.uint32))
        assert np.array_equal(r.tap("cut_cache").view(np.uint32), ref["cut_cache"].view(np.uint32))
        for t in INT_TAPS:
            assert np.array_equal(r.tap(t), ref[ORACLE_NAME[t]]), f"{sc.name}: tap {t} differs"
        assert np.array_equal(r.readback(), ref["rgba"])
        r.render()   # second frame: long-piece walk where the scene has long pieces
        for t in ("intersection", "records"):
            assert np.array_equal(r.tap(t), ref[ORACLE_NAME[t]]), f"{sc.name}: tap {t} differs (frame 2)"
        r.close()
        f = render_gpu(sc, rows, W, H, extra)
        f.render()
        assert np.array_equal(f.readback(), ref["rgba"])
        f.close()
        plain = O.render(sc, rows, W, H, full=full, keep={"rgba", "tpos"})
        assert (plain["rgba"] != ref["rgba"]).any(axis=2).mean() < 0.002
