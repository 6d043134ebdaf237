"""SURVEY section 8 f-3 (beyond the reference, which composites opaque: blendEnable = VK_FALSE, scanline_rasterizer.cpp:893-895):
SLPR_FLAG_BLEND = fills with 0 < alpha < 255 composited "source over" in path order, in the integer arithmetic the
oracle defines next to orc_fill (orc_set_blend). The CUDA result is checked bit for bit against it."""
import dataclasses

import numpy as np
import pytest
import torch

import util
import vkscanlinepr_b200 as V
from oracle import oracle_py as O
from vkscanlinepr_b200 import parallel as PAR
from vkscanlinepr_b200 import scene as S


def with_alphas(sc, seed=5, opaque_share=0.3, clear_share=0.05):
    """The scene with every path's alpha byte redrawn: some opaque, a few fully transparent, the rest 1..254."""
    rng = np.random.default_rng(seed)
    n = len(sc.fill_info)
    a = rng.integers(1, 255, n).astype(np.uint32)
    u = rng.random(n)
    a[u < opaque_share] = 255
    a[u > 1.0 - clear_share] = 0
    return dataclasses.replace(sc, fill_info=(sc.fill_info & np.uint32(0x00FFFFFF)) | (a << np.uint32(24)), name=sc.name + "_alpha")


def test_blend_arithmetic_by_hand():
    """Two overlapping records on a 4 x 2 frame: white <- (200,100,0,a=128) <- (0,0,255,a=64), rounded as documented."""
    def px(r, g, b, a):
        return np.int32(np.uint32(r | (g << 8) | (b << 16) | (a << 24)).view(np.int32))
    rec = np.array([[0, 4, px(200, 100, 0, 128), 0], [2, 2, px(0, 0, 255, 64), 0], [0, 2, px(9, 9, 9, 0), 0]], np.int32)
    img = O.fill(rec, 4, 2, blend=True)
    def over(s, d, a):
        return (s * a + d * (255 - a) + 127) // 255
    first = [over(200, 255, 128), over(100, 255, 128), over(0, 255, 128)]
    second = [over(0, first[0], 64), over(0, first[1], 64), over(255, first[2], 64)]
    assert img[0, 0].tolist() == first + [255] and img[1, 1].tolist() == first + [255]
    assert img[0, 2].tolist() == second + [255] and img[1, 3].tolist() == second + [255]
    # without the flag the later record simply overwrites, alpha byte and all (the reference's behaviour)
    plain = O.fill(rec, 4, 2)
    assert plain[0, 2].tolist() == [0, 0, 255, 64] and plain[0, 0].tolist() == [9, 9, 9, 0]


def test_blend_equals_overwrite_when_every_fill_is_opaque():
    sc = S.synth_scene(300, 256, 192, 6.0, 40.0)
    a = O.render(sc, S.identity_rows(), 256, 192, keep={"rgba"})["rgba"]
    b = O.render(sc, S.identity_rows(), 256, 192, blend=True, keep={"rgba"})["rgba"]
    assert np.array_equal(a, b)


@pytest.mark.gpu
def test_gpu_blend_equals_the_oracle():
    car, cvp = util.full_golden_scene("car")   # 82 distinct alphas, most paths translucent
    tig, tvp = util.golden_scene("tiger")
    cases = [(with_alphas(S.synth_scene(3000, 512, 384, 6.0, 40.0)), S.identity_rows(), 512, 384, 0),
             (car, S.fit_rows(cvp, 900, 600), 900, 600, V.FLAG_FULL_RVG),
             (with_alphas(tig, 9), S.fit_rows(tvp, 640, 480), 640, 480, 0),
             (with_alphas(util.looping_cubics_scene(120, 320, 240), 2), S.identity_rows(), 320, 240, 0),
             (with_alphas(util.tiny_scene(), 3, 0.0, 0.0), S.identity_rows(), 97, 81, 0)]
    for sc, rows, W, H, extra in cases:
        ref = O.render(sc, rows, W, H, full=bool(extra), blend=True, keep={"rgba"})
        plain = O.render(sc, rows, W, H, full=bool(extra), keep={"rgba"})["rgba"]
        for flags in (0, V.FLAG_SEPARATE_FILL, V.FLAG_FUSED_FILL | V.FLAG_RADIX_SORT, V.FLAG_FUSED_FILL | V.FLAG_RECORDS, V.FLAG_TAPS | V.FLAG_NO_GRAPH):
            r = V.ScanlineRasterizer(0, flags | extra | V.FLAG_BLEND).initialize(None, W, H)
            r.loadVG(sc); r.setMVP(rows)
            for _ in range(2):   # the second frame finds the lists and the cell grid re-zeroed
                r.render()
                img = r.readback()
                assert np.array_equal(img, ref["rgba"]), f"{sc.name} flags {flags}: {int((img != ref['rgba']).any(axis=2).sum())} pixels differ"
            r.close()
        if len(sc.fill_info) > 3:
            assert (plain != ref["rgba"]).any(), sc.name   # the case does exercise blending
        # without the flag the alpha bytes change nothing but the framebuffer's alpha channel
        r = V.ScanlineRasterizer(0, extra).initialize(None, W, H)
        r.loadVG(sc); r.setMVP(rows); r.render()
        assert np.array_equal(r.readback(), plain)
        r.close()


@pytest.mark.gpu
def test_gpu_blend_with_four_samples_per_pixel():
    car, cvp = util.full_golden_scene("car")
    for sc, rows, W, H, extra in [(car, S.fit_rows(cvp, 450, 300), 450, 300, V.FLAG_FULL_RVG),
                                  (with_alphas(S.synth_scene(1500, 256, 192, 6.0, 40.0)), S.identity_rows(), 256, 192, 0)]:
        ref = O.render_aa4(sc, rows, W, H, full=bool(extra), blend=True)["rgba_aa"]
        for flags in (0, V.FLAG_SEPARATE_FILL):
            r = V.ScanlineRasterizer(0, flags | extra | V.FLAG_BLEND | V.FLAG_AA4).initialize(None, W, H)
            r.loadVG(sc); r.setMVP(rows); r.render(); r.render()
            assert np.array_equal(r.readback(), ref), (sc.name, flags)
            r.close()


@pytest.mark.gpu
def test_gpu_blend_node_pool_grows():
    """Sixty translucent layers over most of a 1024 x 768 frame ask for far more nodes than the initial pool (2^20): the
    frame is rendered again with a pool that fits, also on the pipelined host path, and stays exact."""
    W, H = 1024, 768
    sc = with_alphas(S.synth_scene(60, W, H, 300.0, 420.0, seed=4), 11, 0.05, 0.0)
    ref = O.render(sc, S.identity_rows(), W, H, blend=True, keep={"rgba"})["rgba"]
    r = V.ScanlineRasterizer(0, V.FLAG_BLEND).initialize(None, W, H)
    r.loadVG(sc); r.setMVP(S.identity_rows()); r.render()
    assert np.array_equal(r.readback(), ref)
    r.close()
    r = V.ScanlineRasterizer(0, V.FLAG_BLEND).initialize(None, W, H)
    r.loadVG(sc)
    outs = [np.zeros((H, W, 4), np.uint8) for _ in range(3)]
    for o in outs:
        r.submit_to_host(S.identity_rows(), o)
    r.wait_host()
    for o in outs:
        assert np.array_equal(o, ref)
    r.close()


@pytest.mark.gpu
def test_gpu_blend_in_exact_row_bands():
    W, H, G = 320, 240, 3
    sc = with_alphas(util.looping_cubics_scene(120, W, H), 2)
    ref = O.render(sc, S.identity_rows(), W, H, blend=True, keep={"rgba"})["rgba"]
    frame = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    ctxs = []
    for y0, y1 in PAR.band_rows(H, G):
        c = V.ScanlineRasterizer(0, V.FLAG_BLEND).initialize(None, W, H)
        c.loadVG(sc); c.setMVP(S.identity_rows()); c.set_band(y0, y1); c.set_target(frame.data_ptr(), W * 4)
        ctxs.append(c)
    boxes = [c.band_mailbox()[0] for c in ctxs]
    for g, c in enumerate(ctxs):
        c.set_band_peers(G, g, 0, boxes)
    for seq in (1, 2):
        for c in ctxs:
            c.prepare()
        for c in ctxs:
            c.render_band(seq)
        ctxs[0].band_wait_gather(seq)
        for c in ctxs:
            c.synchronize()
        assert np.array_equal(frame.cpu().numpy(), ref)
    for c in ctxs:
        c.close()
