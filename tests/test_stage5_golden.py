"""Stage 5 pinned to the reference's own shaders, executed.

tests/golden/stage5_{1,3}.npz hold what the reference's shipped scanlinepr.vert.spv / scanlinepr.frag.spv produce —
run by oracle/spirv_exec.py (tools/make_stage5_golden.py) — for every vertex of the reference's own draw-record dumps
(workdir/test_data.csv, test_data3.csv = tests/golden/ref_records_*.npz; a real run at the 1200x1024 viewport the vertex
shader hard-codes, VERT:39). Here the fixed-function rest of the Vulkan pipeline is applied to those shader outputs —
viewport transform (SR:636), LINE_LIST with lineWidth 2 (SR:884,890), non-strict wide lines and the diamond-exit rule,
flat shading from the provoking vertex, float -> UNORM8 — and the frame must equal what the oracle's stage 5 (orc_fill)
makes of the same records. What remains a reading of the specification is that fixed-function part alone; it is written
here a second time, independently of oracle.c, for arbitrary segments."""
import os

import numpy as np
import pytest

from oracle import oracle_py as O
import util

W, H = 1200, 1024  # the viewport of the run the records were dumped from == the constants in the vertex shader
SUBPIXEL = 256.0   # vertex positions are snapped to a sub-pixel grid before rasterization (Vulkan: >= 4 bits; 8 here)


def load(tag):
    rec = np.load(os.path.join(util.GOLDEN, f"ref_records_{tag}.npz"))["records"]
    g = np.load(os.path.join(util.GOLDEN, f"stage5_{tag}.npz"))
    pos = g["position"].view(np.float32).reshape(-1, 4)
    col = g["color"].view(np.float32).reshape(-1, 4)
    return rec, pos, col, g["frag_pos"], int(g["sample_mask"])


def window_coords(pos):
    """Clip coordinates -> framebuffer coordinates (Vulkan 'Controlling the Viewport': x_f = p_x/2 * x_d + o_x with
    o_x = x + width/2), in float32 like the hardware path, then the sub-pixel snap."""
    ndc = pos[:, :2] / pos[:, 3:4]
    half = np.array([W / 2, H / 2], np.float32)
    xy = ndc * half + half
    return np.round(xy.astype(np.float64) * SUBPIXEL) / SUBPIXEL


def diamond_exit(pa, pb):
    """Pixels (i, j) whose diamond |x - i - 1/2| + |y - j - 1/2| < 1/2 the segment pa -> pb passes through and does not
    end in (Vulkan 'Basic Line Segment Rasterization'). Exact for coordinates on the sub-pixel grid (float64)."""
    (xa, ya), (xb, yb) = pa, pb
    i0, i1 = int(np.floor(min(xa, xb))) - 1, int(np.floor(max(xa, xb))) + 1
    j0, j1 = int(np.floor(min(ya, yb))) - 1, int(np.floor(max(ya, yb))) + 1
    ii, jj = np.meshgrid(np.arange(i0, i1 + 1), np.arange(j0, j1 + 1), indexing="ij")
    cx, cy = ii + 0.5, jj + 0.5
    dx, dy = xb - xa, yb - ya
    lo = np.zeros(ii.shape)
    hi = np.ones(ii.shape)
    ok = np.ones(ii.shape, bool)
    for sx in (1.0, -1.0):  # the open diamond is the intersection of four half planes sx*(x-cx) + sy*(y-cy) < 1/2
        for sy in (1.0, -1.0):
            a = sx * dx + sy * dy                         # along the segment: f(t) = f0 + a t < 1/2
            f0 = sx * (xa - cx) + sy * (ya - cy)
            if a == 0.0:
                ok &= f0 < 0.5
            elif a > 0.0:
                hi = np.minimum(hi, (0.5 - f0) / a)
            else:
                lo = np.maximum(lo, (0.5 - f0) / a)
    inside = ok & (lo < hi)                               # a piece of the segment lies strictly inside the diamond
    end_in = (np.abs(xb - cx) + np.abs(yb - cy)) < 0.5    # ... and the segment does not end there
    sel = inside & ~end_in
    return ii[sel], jj[sel]


def rasterize(rec, pos, col):
    """LINE_LIST, lineWidth = 2, non-strict: an x-major line is moved by -(w - 1)/2 in y, rasterized as a thin line,
    and every fragment is repeated w times towards +y ('Wide Lines'); later primitives overwrite (blendEnable = VK_FALSE,
    SR:893-895); colour attachment cleared to white (SR:622)."""
    fb = np.full((H, W), 0xFFFFFFFF, np.uint32)
    xy = window_coords(pos)
    # float -> UNORM8 (round to nearest), channel order as the oracle stores it (R in the low byte)
    c8 = np.clip(np.round(col.astype(np.float64) * 255.0), 0, 255).astype(np.uint32)
    word = c8[:, 0] | (c8[:, 1] << 8) | (c8[:, 2] << 16) | (c8[:, 3] << 24)
    width = 2
    for k in range(rec.shape[0]):
        pa, pb = xy[2 * k].copy(), xy[2 * k + 1].copy()
        assert abs(pb[0] - pa[0]) >= abs(pb[1] - pa[1]), "the records are horizontal spans: x-major"
        if pa[0] == pb[0]:
            continue  # zero-length line: no fragment
        pa[1] -= (width - 1) / 2
        pb[1] -= (width - 1) / 2
        ii, jj = diamond_exit(pa, pb)
        for r in range(width):
            j = jj + r
            keep = (ii >= 0) & (ii < W) & (j >= 0) & (j < H)
            fb[j[keep], ii[keep]] = word[k]
    return fb


@pytest.mark.parametrize("tag", ["1", "3"])
def test_vertex_shader_decodes_records_like_the_oracle(tag):
    rec, pos, col, frag_pos, mask = load(tag)
    n = rec.shape[0]
    assert pos.shape == (2 * n, 4) and col.shape == (n, 4) and mask == 0xFFFFFFFF
    X, Y, wd = rec[:, 0] & 0xFFFF, rec[:, 0] >> 16, rec[:, 1]   # orc_fill's reading of VERT:27-33
    assert np.array_equal(frag_pos[:, 0], X) and np.array_equal(frag_pos[:, 1], Y)
    xy = window_coords(pos)
    assert np.array_equal(xy[0::2, 0], X) and np.array_equal(xy[1::2, 0], X + wd), "span [X, X + width) on the pixel grid"
    assert np.array_equal(xy[0::2, 1], H - (Y + 1)) and np.array_equal(xy[1::2, 1], H - (Y + 1)), "line at y = Y + 1, flipped (VERT:35,44)"
    assert np.all(pos[:, 2] == 0) and np.all(pos[:, 3] == 1)
    # colour: bytes of fill_info / 255 (VERT:8-15), which UNORM8 conversion maps back onto the bytes
    c8 = np.round(col.astype(np.float64) * 255.0).astype(np.uint32)
    word = c8[:, 0] | (c8[:, 1] << 8) | (c8[:, 2] << 16) | (c8[:, 3] << 24)
    assert np.array_equal(word, rec[:, 2].view(np.uint32))


@pytest.mark.parametrize("tag", ["1", "3"])
def test_oracle_stage5_equals_the_executed_shaders_rasterized(tag):
    rec, pos, col, _, _ = load(tag)
    want = rasterize(rec, pos, col)
    got = O.fill(rec, W, H).view(np.uint32).reshape(H, W)
    assert np.count_nonzero(want != 0xFFFFFFFF) > 1000
    assert np.array_equal(got, want)


def test_diamond_exit_on_sloped_segments():
    """The rasterizer above is not special-cased to horizontal spans: a 45-degree and a shallow segment give the
    pixel sets the rule prescribes (one fragment per column for an x-major line, end pixel excluded)."""
    i, j = diamond_exit((0.5, 0.5), (4.5, 4.5))
    assert sorted(zip(i.tolist(), j.tolist())) == [(0, 0), (1, 1), (2, 2), (3, 3)]
    i, j = diamond_exit((0.5, 0.5), (8.5, 2.25))
    assert sorted(i.tolist()) == list(range(8)) and j.min() == 0 and j.max() == 2
    i, j = diamond_exit((3.0, 1.5), (7.0, 1.5))  # integer ends on the row's centre line: [3, 7)
    assert sorted(i.tolist()) == [3, 4, 5, 6] and set(j.tolist()) == {1}


def test_all_eleven_shaders_end_to_end_at_the_native_viewport():
    """tests/golden/e2e1200.npz: one small scene through ALL of the reference's shipped shaders at its native
    1200x1024 viewport — the nine compute shaders along drawFrame's dispatch sequence, then scanlinepr.vert / .frag
    over the records they produced (tools/make_stage5_golden.py end_to_end). The oracle reproduces every compute
    buffer bit for bit and its frame equals the executed stage-5 shaders' lines under the fixed-function rules."""
    from vkscanlinepr_b200 import scene as S
    import test_spirv_golden as T
    z = np.load(os.path.join(util.GOLDEN, "e2e1200.npz"))
    assert int(z["width"]) == W and int(z["height"]) == H
    sc = S.Scene(z["pos"], z["pos_path"], z["curve_pos_map"], z["curve_type"], z["curve_path"], z["fill_rule"], z["fill_info"], "e2e1200")
    r = O.render(sc, z["rows"], W, H)
    for k in ("n_fragments", "n_out_frag", "n_span"):
        assert int(z[k]) == r[k], k
    T.check_cut_cache(z["cut_cache"], r["cut_cache"])
    for k in T.BUFFERS:
        assert z[k].shape == r[k].shape and np.array_equal(T.bits(z[k]), T.bits(r[k])), f"buffer {k}"
    pos = z["s5_position"].view(np.float32).reshape(-1, 4)
    col = z["s5_color"].view(np.float32).reshape(-1, 4)
    want = rasterize(z["records"], pos, col)
    assert np.count_nonzero(want != 0xFFFFFFFF) > 1000
    assert np.array_equal(r["rgba"].view(np.uint32).reshape(H, W), want)


@pytest.mark.skipif(not os.path.isdir("/root/reference/workdir/shaders"), reason="reference shaders not present")
def test_interpreter_still_reproduces_the_stage5_fixture():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(util.GOLDEN), "..", "tools"))
    import make_stage5_golden as M5
    rec, pos, col, frag_pos, _ = load("1")
    sub = np.concatenate([rec[:40], rec[-40:]])
    out = M5.run_records(sub, log=lambda *a: None)
    assert np.array_equal(out["position"].view(np.float32).reshape(-1, 4), np.concatenate([pos[:80], pos[-80:]]))
    assert np.array_equal(out["color"].view(np.float32).reshape(-1, 4), np.concatenate([col[:40], col[-40:]]))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["1", "3"])
def test_cuda_stage5_equals_the_executed_shaders_rasterized(tag):
    """The CUDA stage 5 (k_fill_cells + k_resolve through slpr_draw_records) draws the reference's own record dumps into
    the frame the reference's own shaders + the fixed-function rules give."""
    import vkscanlinepr_b200 as V
    rec, pos, col, _, _ = load(tag)
    want = rasterize(rec, pos, col)
    r = V.ScanlineRasterizer(0, 0).initialize(None, W, H)
    try:
        r.draw_records(rec)
        got = r.readback().view(np.uint32).reshape(H, W)
        assert np.array_equal(got, want)
        r.draw_records(rec[:0])  # an empty list clears the frame
        assert np.all(r.readback().view(np.uint32) == 0xFFFFFFFF)
        with pytest.raises(V.SlprError):
            r.draw_records(np.array([[3, 2, -1, 0]], np.int32))  # x = 3: off the 2 x 2 fragment grid
    finally:
        r.close()


@pytest.mark.gpu
def test_cuda_draw_records_between_rendered_frames():
    """slpr_draw_records leaves a context with a scene as it was: the next rendered frame is the scene's."""
    import vkscanlinepr_b200 as V
    from vkscanlinepr_b200 import scene as S
    sc, rows = S.synth_scene(300, 512, 512), S.identity_rows()
    r = V.ScanlineRasterizer(0, 0).initialize(None, 512, 512)
    try:
        r.loadVG(sc)
        r.setMVP(rows)
        r.render()
        a = r.readback().copy()
        r.draw_records(np.array([[(10 << 16) | 4, 20, 0x7F112233, 0]], np.int32))
        b = r.readback().view(np.uint32).reshape(512, 512)
        assert np.count_nonzero(b != 0xFFFFFFFF) == 40 and np.all(b[512 - 1 - 11:512 - 1 - 9, 4:24] == 0x7F112233)
        r.render()
        assert np.array_equal(r.readback(), a)
    finally:
        r.close()
