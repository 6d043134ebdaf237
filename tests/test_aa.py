"""SURVEY section 8 f-3 (beyond the reference, which has no antialiasing): SLPR_FLAG_AA4 = four coverage samples per pixel, ordered
opaque compositing per sample, box filter. Defined through the reference path itself — the frame at 4x, averaged
(oracle_py.render_aa4) — so the CUDA result is checked bit for bit; quality is checked against an independent
renderer supersampled the same way."""
import numpy as np
import pytest

import util
import vkscanlinepr_b200 as V
from oracle import oracle_py as O
from vkscanlinepr_b200 import scene as S


def test_box4_definition():
    hi = np.zeros((8, 8, 4), np.uint8)
    hi[0:2, 0:2] = 255      # one cell of pixel (0, 0)
    hi[4:8, 4:8] = 100      # all four cells of pixel (1, 1)
    out = O.box4(hi)
    assert out.shape == (2, 2, 4) and out[0, 0, 0] == 64 and out[1, 1, 0] == 100 and out[0, 1, 0] == 0


def test_aa4_oracle_is_closer_to_a_supersampled_independent_render():
    """Quality: against the independent pixel-centre renderer run at 4x and box-filtered over 4 x 4 (16 samples), the
    AA4 frame must beat the plain frame by at least 5 dB PSNR on an edge-dense scene."""
    sc = util.quad_arc_scene(60, 320, 240)
    W, H = 160, 120
    rows = S.identity_rows() * np.float32(0.5); rows[2, 2] = rows[3, 3] = 1
    r4 = rows.copy(); r4[0] *= 4; r4[1] *= 4
    ind = util.point_sampled_fill(sc, r4, 4 * W, 4 * H).astype(np.float64).reshape(H, 4, W, 4, 4).mean(axis=(1, 3))
    aa = O.render_aa4(sc, rows, W, H, full=True)["rgba_aa"]
    plain = O.render(sc, rows, W, H, full=True, keep={"rgba"})["rgba"]
    p_aa, p_plain = util.psnr(aa, ind), util.psnr(plain, ind)
    assert p_aa > p_plain + 5.0 and p_aa > 24.0, (p_aa, p_plain)


@pytest.mark.gpu
def test_gpu_aa4_equals_the_4x_oracle_averaged():
    tig, vp = util.golden_scene("tiger")
    car, cvp = util.full_golden_scene("car")
    cases = [(tig, S.fit_rows(vp, 640, 480), 640, 480, 0), (S.synth_scene(3000, 512, 384, 6.0, 30.0), S.identity_rows(), 512, 384, 0),
             (car, S.fit_rows(cvp, 450, 300), 450, 300, V.FLAG_FULL_RVG), (util.tiny_scene(), S.identity_rows(), 97, 81, 0)]
    for sc, rows, W, H, extra in cases:
        ref = O.render_aa4(sc, rows, W, H, full=bool(extra))
        for flags in (V.FLAG_AA4, V.FLAG_AA4 | V.FLAG_SEPARATE_FILL, V.FLAG_AA4 | V.FLAG_FUSED_FILL | V.FLAG_RADIX_SORT):
            r = V.ScanlineRasterizer(0, flags | extra).initialize(None, W, H)
            r.loadVG(sc); r.setMVP(rows); r.render()
            img = r.readback()
            assert img.shape == (H, W, 4)
            assert np.array_equal(img, ref["rgba_aa"]), f"{sc.name}: {int((img != ref['rgba_aa']).any(axis=2).sum())} pixels differ"
            assert r.counts()["n_fragments"] == ref["n_fragments"]
            r.render()
            assert np.array_equal(r.readback(), ref["rgba_aa"])
            r.close()
    # the antialiased frame has intermediate levels, the plain one does not
    plain = O.render(tig, S.fit_rows(vp, 640, 480), 640, 480, keep={"rgba"})["rgba"]
    aa = O.render_aa4(tig, S.fit_rows(vp, 640, 480), 640, 480)["rgba_aa"]
    assert len(np.unique(aa.reshape(-1, 4), axis=0)) > 3 * len(np.unique(plain.reshape(-1, 4), axis=0))


@pytest.mark.gpu
def test_gpu_aa4_row_bands_and_limits():
    W = H = 256
    rows = S.identity_rows() * np.float32(0.5); rows[2, 2] = rows[3, 3] = 1
    r4 = rows.copy(); r4[0] *= 4; r4[1] *= 4
    for seed in range(3, 60):  # independent bands (no exchange) are exact only without a winding residue: find such a scene
        sc = S.synth_scene(1024, 512, 512, 6.0, 40.0, seed=seed)
        hi = O.render(sc, r4, 4 * W, 4 * H, do_fill=False, keep={"path", "wind", "wn"})
        res = np.zeros(sc.n_paths, np.int64)
        np.add.at(res, hi["path"], hi["wind"])
        if hi["wn"][-1] == 0 and not res.any():
            break
    ref = O.render_aa4(sc, rows, W, H)["rgba_aa"]
    full = np.zeros((H, W, 4), np.uint8)
    for g in range(4):
        y0, y1 = g * H // 4, (g + 1) * H // 4
        r = V.ScanlineRasterizer(0, V.FLAG_AA4).initialize(None, W, H)
        r.loadVG(sc); r.setMVP(rows); r.set_band(y0, y1); r.render()
        full[H - y1:H - y0] = r.readback()[H - y1:H - y0]
        r.close()
    assert np.array_equal(full, ref)
    with pytest.raises(V.SlprError):
        V.ScanlineRasterizer(0, V.FLAG_AA4).initialize(None, 9000, 100)
