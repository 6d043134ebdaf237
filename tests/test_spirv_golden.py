"""Parity against the REFERENCE'S OWN SHIPPED BINARIES. tests/golden/spirv_*.npz hold every buffer of a
frame produced by executing workdir/shaders/**/spv/*.comp.spv (the reference's prebuilt SPIR-V) with
oracle/spirv_exec.py along the dispatch sequence of drawFrame (tools/make_spirv_golden.py). The C oracle
must reproduce all of them bit for bit (CPU), and so must the CUDA path (GPU)."""
import glob
import os

import numpy as np
import pytest

import util
from oracle import oracle_py as O
from vkscanlinepr_b200 import scene as S

CASES = sorted(os.path.basename(p)[6:-4] for p in glob.glob(os.path.join(util.GOLDEN, "spirv_*.npz")))
BUFFERS = ["tpos", "path_visible", "curve_count", "curve_offset", "inter", "key", "idx", "path", "wind", "seg",
           "skey", "sidx", "swind", "wn", "flags", "scan3", "records"]


def load_case(name):
    z = np.load(os.path.join(util.GOLDEN, f"spirv_{name}.npz"))
    sc = S.Scene(z["pos"], z["pos_path"], z["curve_pos_map"], z["curve_type"], z["curve_path"], z["fill_rule"],
                 z["fill_info"], name)
    return z, sc, int(z["width"]), int(z["height"])


def bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


def check_cut_cache(gold, got):
    """Slots >= n_cuts are uninitialised shared memory in the reference (MI0:310-313): compare the rest."""
    g, o = gold.view(np.uint32), got.view(np.uint32)
    assert np.array_equal(g[:, 4], o[:, 4]), "n_cuts"
    for c in range(g.shape[0]):
        n = int(g[c, 4])
        assert np.array_equal(g[c, :n], o[c, :n]), f"cut parameters of curve {c}"


def test_fixtures_exist():
    assert {"tiny", "glyphs", "edge"} <= set(CASES)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_shipped_spirv(name):
    z, sc, W, H = load_case(name)
    r = O.render(sc, z["rows"], W, H)
    for k in ("n_fragments", "n_out_frag", "n_span"):
        assert int(z[k]) == r[k], k
    assert int(z["key_sentinel"]) == -1                      # gen_fragment.comp:240
    check_cut_cache(z["cut_cache"], r["cut_cache"])
    for k in BUFFERS:
        assert z[k].shape == r[k].shape and np.array_equal(bits(z[k]), bits(r[k])), f"{name}: buffer {k}"
    util.check_record_invariants(z["records"], width=W)


def test_edge_fixture_covers_the_corner_cases():
    z, sc, W, H = load_case("edge")
    ncuts = z["cut_cache"][:, 4].view(np.uint32)
    assert (ncuts == 4).any(), "a 4-cut cubic exercises the MI0:340 slip"
    assert (sc.curve_type == S.QUADRIC).any(), "QUADRIC TODO arms"
    assert (z["curve_count"][sc.curve_path == 4] == 0).all(), "invisible path"
    assert (z["key"] == np.int32(-65538)).any(), "invalid keys (0xFFFEFFFE)"


def test_interpreter_still_reproduces_a_fixture():
    """Re-run the first two shaders of the reference through the interpreter (fast) when the reference is here."""
    spv = "/root/reference/workdir/shaders/scanline/compute/spv/transform_pos.comp.spv"
    if not os.path.exists(spv):
        pytest.skip("reference binaries are not present on this box")
    from oracle import spirv_exec as SX
    z, sc, W, H = load_case("tiny")
    U32 = np.uint32
    ubo = np.zeros(20, U32)
    ubo[0] = sc.n_points
    ubo[1:3] = np.array([W, H], np.float32).view(U32)
    ubo[4:20] = np.ascontiguousarray(z["rows"], np.float32).reshape(16).view(U32)
    tpos = np.zeros(2 * sc.n_points, U32)
    pvis = np.zeros(sc.n_paths, U32)
    r = SX.Runner(SX.Module(spv), {0: (ubo, 0), 1: (sc.pos.reshape(-1).view(U32).copy(), 0), 2: (sc.pos_path.copy(), 0),
                                   3: (tpos, 0), 4: (pvis, 0)})
    with np.errstate(all="ignore"):
        r.dispatch((sc.n_points + 255) // 256)
    assert np.array_equal(tpos, z["tpos"].reshape(-1).view(U32))
    assert np.array_equal(pvis.view(np.int32), z["path_visible"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_shipped_spirv(name):
    import vkscanlinepr_b200 as V
    z, sc, W, H = load_case(name)
    r = V.ScanlineRasterizer(0, V.FLAG_TAPS | V.FLAG_NO_GRAPH).initialize(None, W, H)
    r.loadVG(sc)
    r.setMVP(z["rows"])
    r.render()
    cnt = r.counts()
    assert cnt == {k: int(z[k]) for k in ("n_fragments", "n_out_frag", "n_span")}
    check_cut_cache(z["cut_cache"], r.tap("cut_cache"))
    tapname = dict(tpos="transformed_pos", path_visible="path_visible", curve_count="curve_count", curve_offset="curve_offset",
                   inter="intersection", path="path", wind="winding", seg="segments", skey="sorted_key", sidx="sorted_index",
                   wn="winding_scan", flags="flags", scan3="flag_scan", records="records")
    for k, t in tapname.items():
        got = r.tap(t)
        assert got.shape == z[k].shape and np.array_equal(bits(got), bits(z[k])), f"{name}: tap {t}"
    key = r.tap("key")
    assert np.array_equal(key[:-1], z["key"]) and key[-1] == -1
    r.close()
