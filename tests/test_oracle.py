"""CPU tests of the oracle itself (the checker): pinned against the reference's only golden data
for this path (workdir/test_data*.csv, committed as tests/golden/ref_records_*.npz by
tools/make_golden.py) and against size-independent properties."""
import os

import numpy as np
import pytest

from oracle import oracle_py as O
from vkscanlinepr_b200 import scene as S
import util


@pytest.mark.parametrize("tag,n", [("1", 4020), ("3", 27073)])
def test_reference_csv_invariants(tag, n):
    rec = np.load(os.path.join(util.GOLDEN, f"ref_records_{tag}.npz"))["records"]
    assert rec.shape == (n, 4)
    nfrag, nspan = util.check_record_invariants(rec, width=1200)
    assert nfrag + nspan == n
    # the reference's stage 5 (restated in orc_fill) must accept its own dumps: 1200x1024 viewport (VERT:39)
    img = O.fill(rec, 1200, 1024)
    assert img.shape == (1024, 1200, 4)
    cols = {int(c) for c in np.unique(rec[:, 2])}
    seen = {int(v) for v in np.unique(img.view(np.uint32))}
    assert seen <= ({c & 0xFFFFFFFF for c in cols} | {0xFFFFFFFF})


def test_reference_csv3_row_order_is_signed_key_order():
    """test_data3.csv rows run y=2,4,.. then y=0 within a path: the signed int32 key order of
    naive_seg_sort_pairs.comp:69. The oracle's comparator must reproduce it."""
    rec = np.load(os.path.join(util.GOLDEN, "ref_records_3.npz"))["records"]
    frag = rec[rec[:, 3] != 0]
    same_col = frag[frag[:, 2] == frag[0, 2]]
    y = same_col[:, 0] >> 16
    first_path = y[: np.argmax(np.diff(y) < 0) + 1] if np.any(np.diff(y) < 0) else y
    assert first_path[0] != 0 or len(first_path) == 1
    # encode as the reference key and check the oracle's sort keeps this order
    x = same_col[:, 0] & 0xFFFF
    key = (((y + 0x7FFF) << 16) | ((x + 0x7FFF) & 0xFFFF)).astype(np.uint32).view(np.int32)
    k0 = key[: len(first_path)]
    seg = np.array([0, len(k0)], np.int32)
    ks, _ = O.seg_sort(seg, k0, np.arange(len(k0), dtype=np.int32))
    assert np.array_equal(ks, k0), "CSV order within a path == oracle signed (key,index) order"


def test_sort_matches_literal_odd_even_network():
    rng = np.random.default_rng(7)
    n = 3000
    seg = np.array([0, 1, 1, 700, 701, 2500, n], np.int32)
    key = rng.integers(-2**31, 2**31 - 1, n, dtype=np.int64).astype(np.int32)
    key[100:400] = key[100]  # ties broken by index
    idx = np.arange(n, dtype=np.int32)
    a = O.seg_sort(seg, key, idx)
    b = O.seg_sort(seg, key, idx, literal=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_scan_semantics():
    rng = np.random.default_rng(1)
    a = rng.integers(-5, 9, 10001).astype(np.int32)
    out = O.exclusive_scan(a)
    assert out[0] == 0 and np.array_equal(out[1:], np.cumsum(a, dtype=np.int64).astype(np.int32))


def test_colour_quantisation():
    assert O.quantise_colour([1, 1, 1, 1], 1.0) == 0xFFFFFFFF
    assert O.quantise_colour([1, 0, 0, 1], 0.5) == ((127 << 24) | 0xFF)
    assert O.quantise_colour([1, 1, 1, 1], 0.0) == 0          # alpha byte 0 -> whole word 0 (SR.cpp:117)
    assert O.quantise_colour([0.5, 0.25, 0, 1], 1.0) == (0xFF000000 | (63 << 8) | 127)


@pytest.mark.parametrize("name", ["test", "tiger", "reschart"])
def test_frame_properties_on_shipped_scenes(name):
    sc, vp = util.golden_scene(name)
    W = H = 512
    r = O.render(sc, S.fit_rows(vp, W, H), W, H)
    nf = r["n_fragments"]
    assert nf == r["curve_offset"][-1] == int(r["curve_count"].sum())
    # every curve's records are its own, contiguous, in walk order
    assert np.array_equal(r["inter"][:, 0], np.repeat(np.arange(sc.n_curves), r["curve_count"]))
    # sorted planes are a permutation that never crosses a path boundary and is ordered by (key, idx)
    assert np.array_equal(np.sort(r["sidx"]), np.arange(nf))
    assert np.array_equal(r["path"][r["sidx"]], r["path"])
    for p in range(sc.n_paths):
        b, e = r["seg"][p], r["seg"][p + 1]
        k, i = r["skey"][b:e].astype(np.int64), r["sidx"][b:e].astype(np.int64)
        assert np.all((k[1:] > k[:-1]) | ((k[1:] == k[:-1]) & (i[1:] > i[:-1])))
    # winding scan is the global exclusive prefix of the shuffled deltas
    assert np.array_equal(r["wn"][1:], np.cumsum(r["swind"]))
    util.check_record_invariants(r["records"], width=W)
    assert r["records"].shape[0] == r["n_out_frag"] + r["n_span"]
    # stage 5 is idempotent w.r.t. re-running on its own records
    assert np.array_equal(O.fill(r["records"], W, H), r["rgba"])


def test_empty_scene_is_white():
    sc, vp = util.golden_scene("car")
    r = O.render(sc, S.identity_rows(), 64, 48)
    assert r["n_fragments"] == 0 and r["records"].shape[0] == 0
    assert np.all(r["rgba"] == 255)


def test_invisible_path_emits_nothing():
    sc = util.tiny_scene()
    r = O.render(sc, S.identity_rows(), 96, 80)
    off_curves = np.nonzero(sc.curve_path == 2)[0]
    assert np.all(r["curve_count"][off_curves] == 0)
    assert r["n_fragments"] > 0
    # closed paths: winding deltas sum to zero per (path,row)
    assert r["wn"][-1] == 0
