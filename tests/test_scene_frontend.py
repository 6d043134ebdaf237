"""Host scene front end (libslpr.so, slpr_vg_*): pinned against the REFERENCE's own RVG parser
(golden containers written by tools/make_golden.py from oracle/_ref/rvg_dump) and against a numpy
restatement of loadVG's flattening (scanline_rasterizer.cpp:67-118)."""
import os

import numpy as np
import pytest

import util
import vkscanlinepr_b200 as V
from oracle import oracle_py as O
from vkscanlinepr_b200 import scene as S

FIELDS = ("pos", "curve_pos", "curve_type", "path_curve", "fill_rule", "fill_color", "fill_opacity")
COUNTS = {"test": (864, 264, 17), "tiger": (60060, 28042, 302), "reschart": (17602, 8705, 747),
          "drops": (3348, 859, 81), "embrace": (14718, 3687, 184), "car": (0, 0, 1), "chord": (0, 0, 1),
          "chord-black": (0, 0, 1)}


@pytest.mark.parametrize("name", list(COUNTS))
def test_golden_counts_match_survey(name):
    c = util.golden_container(name)
    assert (c.pos.shape[0], len(c.curve_pos), len(c.path_curve)) == COUNTS[name]


@pytest.mark.parametrize("name", list(COUNTS))
def test_rvg_parser_matches_reference_parser(name):
    path = os.path.join(util.REF_RVG_DIR, name + ".rvg")
    if not os.path.exists(path):
        pytest.skip("reference scenes are not present on this box")
    sc, vp, cont = V.load_rvg(path)
    ref = util.golden_container(name)
    for f in FIELDS:
        assert np.array_equal(getattr(cont, f), getattr(ref, f)), f
    assert np.array_equal(vp, ref.vp)
    exp = S.flatten_reference(ref, name)
    for a, b in zip(sc.arrays(), exp.arrays()):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", util.SHIPPED + util.EMPTY)
def test_flatten_matches_loadvg_restatement(name):
    c = util.golden_container(name)
    got = V.flatten(c, name)
    exp = S.flatten_reference(c, name)
    for a, b in zip(got.arrays(), exp.arrays()):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    # colour words agree with the oracle's quantiser (SR.cpp:107-118)
    for i in range(min(len(c.path_curve), 50)):
        assert int(got.fill_info[i]) == O.quantise_colour(c.fill_color[i], float(c.fill_opacity[i]))


def test_missing_file_reports_like_the_reference():
    with pytest.raises(V.SlprError, match="can't open file"):
        V.load_rvg("/nonexistent/scene.rvg")


def test_parser_quirks(tmp_path):
    """rvg.cpp behaviour: implicit close to the LAST M contour; Z sets closed; lower-case
    commands end the command loop; a non-solid paint keeps (0,0,0,1) with alpha = opacity."""
    txt = """viewport 0,0 100,100
window 0,0 100,100
scene dyn_identity
 1 element nzfill dyn_concrete 0,0 0,0 0,0: M 0,0 L 10,0 L 10,10 M 20,20 L 30,20 L 30,30 dyn_identity dyn_paint 0.5 solid rgba(1,0.5,0,1)
 1 element ofill dyn_concrete 0,0 0,0 0,0: M 1,1 C 2,2 3,3 4,4 L 1,1 Z dyn_identity dyn_paint 1 solid rgb(0,1,0)
"""
    p = tmp_path / "q.rvg"
    p.write_text(txt)
    sc, vp, c = V.load_rvg(str(p))
    assert sc.n_paths == 2
    # path 0: 4 lines + closing line of the second contour only (30,30)->(20,20)
    assert list(c.curve_type[:5]) == [2, 2, 2, 2, 2] and np.allclose(c.pos[8:10], [[30, 30], [20, 20]])
    # path 1: cubic + line, closed by Z (no extra curve)
    assert list(c.curve_type[5:]) == [4, 2]
    assert sc.fill_rule.tolist() == [0, 1]
    assert sc.fill_info[0] == (127 << 24) | (0 << 16) | (127 << 8) | 255
    assert sc.fill_info[1] == 0xFF00FF00
