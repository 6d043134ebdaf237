"""Shared helpers for the tests: golden scenes, matrices, oracle access (tests may use oracle/)."""
import os

import numpy as np

from vkscanlinepr_b200 import scene as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SHIPPED = ["test", "tiger", "reschart", "drops", "embrace"]
EMPTY = ["car", "chord", "chord-black"]
REF_RVG_DIR = "/root/reference/workdir/input/rvg"


def golden_container(name):
    return S.Container.from_npz(os.path.join(GOLDEN, name + ".npz"))


def golden_scene(name):
    c = golden_container(name)
    return S.flatten_reference(c, name), c.vp


def tiny_scene():
    """Two overlapping closed paths (a square of lines, a blob of cubics) + one off-screen path."""
    pos, pos_path, cpm, ctype, cpath = [], [], [], [], []

    def add(curve_pts, path):
        cpm.append(len(pos)); ctype.append(S.LINE if len(curve_pts) == 2 else S.CUBIC); cpath.append(path)
        for p in curve_pts:
            pos.append(p); pos_path.append(path)

    sq = [(10.5, 8.25), (50.0, 8.25), (50.0, 40.0), (10.5, 40.0)]
    for i in range(4):
        add([sq[i], sq[(i + 1) % 4]], 0)
    blob = [(30, 20), (45, 5), (75, 15), (70, 35), (65, 60), (40, 62), (25, 45), (20, 30), (22, 22), (30, 20)]
    add(blob[0:4], 1); add(blob[3:7], 1); add(blob[6:10], 1)
    off = [(-50, -50), (-40, -50), (-40, -40)]
    for i in range(3):
        add([off[i], off[(i + 1) % 3]], 2)
    return S.Scene(np.array(pos, np.float32), np.array(pos_path, np.uint32), np.array(cpm, np.uint32),
                   np.array(ctype, np.uint32), np.array(cpath, np.uint32), np.array([0, 1, 0], np.uint32),
                   np.array([0xFF0000FF, 0xFF00FF00, 0xFFFF0000], np.uint32), "tiny")


def psnr(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = float(np.mean(d * d))
    return float("inf") if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def check_record_invariants(rec, width=None):
    """Invariants of the reference's output_buf (gen_merged_fragment_and_span.comp:62-102) that hold
    for test_data.csv / test_data3.csv and must hold for every record list we produce."""
    rec = np.asarray(rec).reshape(-1, 4)
    yx, w, fi = rec[:, 0], rec[:, 1], rec[:, 3]
    x, y = yx & 0xFFFF, yx >> 16
    is_frag = fi != 0
    assert np.all(w[is_frag] == 2), "fragment records have width 2"
    assert np.all((x % 2 == 0) & (y % 2 == 0)), "records sit on the even 2x2 grid"
    assert np.all(w % 2 == 0) and np.all(w > 0)
    assert np.array_equal(fi[is_frag], np.arange(1, is_frag.sum() + 1)), "frag_index counts 1..N in order"
    # Emit order (GEN:64-66,83): fragment i at slot oi, then its span (from the previous fragment's
    # x+2 up to fragment i's x) at oi+frag_flag. So a span record that directly follows a fragment
    # record on its row either ENDS at that fragment (its own fragment), or STARTS right after it
    # (its own fragment lies beyond the right edge and got no record).
    sp = np.nonzero(~is_frag)[0]
    sp = sp[sp > 0]
    prev = sp - 1
    m = is_frag[prev] & (y[sp] == y[prev])
    own = x[sp[m]] + w[sp[m]] == x[prev[m]]
    after = x[sp[m]] == x[prev[m]] + 2
    assert np.all(own | after)
    if width is not None:
        assert np.all((x[sp[m]] + w[sp[m]])[~own] >= width)
    return int(is_frag.sum()), int((~is_frag).sum())


def edge_scene():
    """Corner cases for a 64x48 frame: a cubic with four monotonic cuts (the MI0:340 slip leaves them out of
    order), a curve typed QUADRIC (the shaders' TODO arms), edges on exact grid lines and exact negative
    integers (float2int_rd vs floor), curves leaving the frame on all four sides, a path entirely outside."""
    pos, pos_path, cpm, ctype, cpath = [], [], [], [], []

    def add(pts, path, typ=None):
        cpm.append(len(pos)); ctype.append(typ if typ is not None else (S.LINE if len(pts) == 2 else S.CUBIC)); cpath.append(path)
        for p in pts:
            pos.append(p); pos_path.append(path)

    # path 0: closed shape whose first segment is an S-shaped cubic with 2 x-extrema and 2 y-extrema
    add([(8, 8), (60, 4), (-6, 40), (50, 30)], 0)
    add([(50, 30), (52, 44)], 0)
    add([(52, 44), (6, 42)], 0)
    add([(6, 42), (8, 8)], 0)
    # path 1: axis-aligned rectangle on exact even grid lines, partly left of and below the frame (exact negative integers)
    rect = [(-4.0, -2.0), (20.0, -2.0), (20.0, 10.0), (-4.0, 10.0)]
    for i in range(4):
        add([rect[i], rect[(i + 1) % 4]], 1)
    # path 2: big triangle leaving the frame on the right and the top
    tri = [(30.5, 20.25), (90.0, 25.0), (40.0, 70.0)]
    for i in range(3):
        add([tri[i], tri[(i + 1) % 3]], 2)
    # path 3: contains a QUADRIC-typed curve (3 points; the reference's parser never emits one) between two lines
    add([(12, 30), (20, 36)], 3)
    add([(20, 36), (26, 28), (30, 34)], 3, typ=S.QUADRIC)
    add([(30, 34), (12, 30)], 3)
    # path 4: entirely outside (invisible)
    out = [(100, 100), (120, 100), (110, 120)]
    for i in range(3):
        add([out[i], out[(i + 1) % 3]], 4)
    # path 5: a closed blob of two cubics (1-2 cuts each), even-odd
    add([(34, 6), (48, 2), (58, 14), (44, 18)], 5)
    add([(44, 18), (30, 22), (24, 10), (34, 6)], 5)
    return S.Scene(np.array(pos, np.float32), np.array(pos_path, np.uint32), np.array(cpm, np.uint32),
                   np.array(ctype, np.uint32), np.array(cpath, np.uint32), np.array([0, 1, 0, 0, 0, 1], np.uint32),
                   np.array([0xFF0000FF, 0xFF00FF00, 0xFFFF0000, 0xFF00FFFF, 0xFFFFFFFF, 0xFF808080], np.uint32), "edge")


def looping_cubics_scene(n_paths=600, width=512, height=384, seed=7):
    """Paths of one S-shaped or looping cubic closed by a line, control points scattered over (and beyond) the
    frame: many cubics with three and four monotonic cuts, i.e. plenty of pieces left out of order by the
    MI0:340 slip — the case in which a piece's first record differs from its start parameter."""
    rng = np.random.default_rng(seed)
    pos, pos_path, cpm, ctype, cpath = [], [], [], [], []
    for p in range(n_paths):
        cx, cy = rng.uniform(0, width), rng.uniform(0, height)
        r = rng.uniform(10, 90)
        pts = [(cx + rng.uniform(-r, r), cy + rng.uniform(-r, r)) for _ in range(4)]
        if p % 3 == 0:  # force the S shape of edge_scene(): x and y both turn twice
            pts = [(cx - r, cy - r), (cx + 1.6 * r, cy - 1.2 * r), (cx - 1.5 * r, cy + 1.1 * r), (cx + r, cy + 0.7 * r)]
        cpm.append(len(pos)); ctype.append(S.CUBIC); cpath.append(p)
        for q in pts:
            pos.append(q); pos_path.append(p)
        cpm.append(len(pos)); ctype.append(S.LINE); cpath.append(p)
        for q in (pts[3], pts[0]):
            pos.append(q); pos_path.append(p)
    col = (0xFF000000 | rng.integers(0, 1 << 24, n_paths)).astype(np.uint32)
    return S.Scene(np.array(pos, np.float32), np.array(pos_path, np.uint32), np.array(cpm, np.uint32),
                   np.array(ctype, np.uint32), np.array(cpath, np.uint32), (np.arange(n_paths) & 1).astype(np.uint32),
                   col, "looping_cubics")


def write_rvg(container, path):
    """RVG text (the subset the reference's parser reads: M / L / C / Z, solid paints) of a golden container: one
    `element` per path, a new `M` wherever a curve does not start where the previous one ended. Numbers are printed
    with 9 significant digits so that sscanf("%f") reads back the same floats."""
    c = container
    n_paths, n_curves = len(c.path_curve), len(c.curve_type)
    fmt = lambda p: f"{float(p[0]):.9g},{float(p[1]):.9g}"
    with open(path, "w") as f:
        f.write(f"viewport {c.vp[0]:.9g},{c.vp[1]:.9g} {c.vp[2]:.9g},{c.vp[3]:.9g}\n")
        f.write(f"window {c.vp[0]:.9g},{c.vp[1]:.9g} {c.vp[2]:.9g},{c.vp[3]:.9g}\nscene dyn_identity\n")
        for p in range(n_paths):
            c0 = int(c.path_curve[p])
            c1 = int(c.path_curve[p + 1]) if p + 1 < n_paths else n_curves
            if c1 <= c0:
                continue
            cmds, last = [], None
            for k in range(c0, c1):
                t = int(c.curve_type[k])
                npt = 2 if t == S.LINE else 4 if t == S.CUBIC else 0
                if npt == 0:
                    continue
                pts = c.pos[int(c.curve_pos[k]):int(c.curve_pos[k]) + npt]
                if last is None or tuple(pts[0]) != last:
                    cmds.append("M " + fmt(pts[0]))
                cmds.append(("L " + fmt(pts[1])) if npt == 2 else ("C " + " ".join(fmt(q) for q in pts[1:])))
                last = tuple(pts[-1])
            rule = "ofill" if int(c.fill_rule[p]) == S.EVEN_ODD else "nzfill"
            col = ",".join(f"{float(v):.9g}" for v in c.fill_color[p])
            f.write(f"  1 element {rule} dyn_concrete 0,0 0,0 0,0: {' '.join(cmds)} dyn_identity dyn_paint "
                    f"{float(c.fill_opacity[p]) if float(c.fill_opacity[p]) > 0 else 1:.9g} solid rgba({col})\n")


# ---------------------------------------------------------------------------------------------------------------
# SURVEY section 8 f-1 (full RVG: quadratics and rational arcs)
def full_golden_scene(name):
    """tests/golden/full_<name>.npz (tools/make_full_golden.py): a shipped scene through the complete RVG reader."""
    z = np.load(os.path.join(GOLDEN, f"full_{name}.npz"))
    sc = S.Scene(z["pos"], z["pos_path"], z["curve_pos_map"], z["curve_type"], z["curve_path"], z["fill_rule"], z["fill_info"], "full_" + name)
    sc.curve_weight = z["curve_weight"].astype(np.float32)
    return sc, z["vp"]


def quad_arc_scene(n_paths=300, width=640, height=480, seed=5):
    """Closed paths that mix all four curve types: line, quadratic, rational arc (weights 0.2 .. 3), cubic."""
    rng = np.random.default_rng(seed)
    pos, pos_path, cpm, ctype, cpath, wts = [], [], [], [], [], []
    for p in range(n_paths):
        cx, cy = rng.uniform(-20, width + 20), rng.uniform(-20, height + 20)
        r = rng.uniform(8, 70)
        k = int(rng.integers(3, 7))
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        vert = [(cx + r * np.cos(a) * rng.uniform(0.6, 1.0), cy + r * np.sin(a) * rng.uniform(0.6, 1.0)) for a in ang]
        for i in range(k):
            a, b = vert[i], vert[(i + 1) % k]
            mid = ((a[0] + b[0]) / 2, (a[1] + b[1]) / 2)
            out = (mid[0] + (mid[0] - cx) * rng.uniform(-0.4, 1.2), mid[1] + (mid[1] - cy) * rng.uniform(-0.4, 1.2))
            kind = int(rng.integers(0, 4))
            cpm.append(len(pos)); cpath.append(p)
            if kind == 0:
                pts, t, w = [a, b], S.LINE, 1.0
            elif kind == 1:
                pts, t, w = [a, out, b], S.QUADRIC, 1.0
            elif kind == 2:
                pts, t, w = [a, out, b], S.ARC, float(rng.uniform(0.2, 3.0))
            else:
                o2 = (out[0] + rng.uniform(-r, r) * 0.5, out[1] + rng.uniform(-r, r) * 0.5)
                pts, t, w = [a, out, o2, b], S.CUBIC, 1.0
            ctype.append(t); wts.append(w)
            for q in pts:
                pos.append(q); pos_path.append(p)
    col = (0xFF000000 | rng.integers(0, 1 << 24, n_paths)).astype(np.uint32)
    sc = S.Scene(np.array(pos, np.float32), np.array(pos_path, np.uint32), np.array(cpm, np.uint32), np.array(ctype, np.uint32),
                 np.array(cpath, np.uint32), (np.arange(n_paths) % 2).astype(np.uint32), col, "quad_arc")
    sc.curve_weight = np.array(wts, np.float32)
    return sc


def point_sampled_fill(sc, rows, W, H, steps=48):
    """An independent renderer for the full-RVG mode: every curve flattened in float64 (lines, quadratics, rational
    arcs, cubics), affine `rows` applied, each path filled by its rule at PIXEL CENTRES with a plain scanline crossing
    test, painted opaque in path order on white, image row = H-1-y like the pipeline. No 2x2 cells, no bisection."""
    m = np.asarray(rows, np.float64)
    img = np.full((H, W, 4), 255, np.uint8)
    t = np.linspace(0.0, 1.0, steps + 1)[:, None]
    w_all = sc.curve_weight if sc.curve_weight is not None else np.ones(sc.n_curves, np.float32)
    first_curve = np.searchsorted(sc.curve_path, np.arange(sc.n_paths + 1))
    yc = np.arange(H) + 0.5
    for p in range(sc.n_paths):
        segs = []
        for c in range(first_curve[p], first_curve[p + 1]):
            ty, po = int(sc.curve_type[c]), int(sc.curve_pos_map[c])
            P = sc.pos[po:po + (ty & 7)].astype(np.float64)
            if ty == S.LINE:
                pts = P
            elif ty == S.QUADRIC:
                pts = (1 - t) ** 2 * P[0] + 2 * t * (1 - t) * P[1] + t ** 2 * P[2]
            elif ty == S.ARC:
                w = float(w_all[c])
                b0, b1, b2 = (1 - t) ** 2, 2 * t * (1 - t) * w, t ** 2
                pts = (b0 * P[0] + b1 * P[1] + b2 * P[2]) / (b0 + b1 + b2)
            else:
                pts = (1 - t) ** 3 * P[0] + 3 * t * (1 - t) ** 2 * P[1] + 3 * t ** 2 * (1 - t) * P[2] + t ** 3 * P[3]
            x = m[0, 0] * pts[:, 0] + m[0, 1] * pts[:, 1] + m[0, 3]
            y = m[1, 0] * pts[:, 0] + m[1, 1] * pts[:, 1] + m[1, 3]
            segs.append(np.stack([x[:-1], y[:-1], x[1:], y[1:]], 1))
        if not segs:
            continue
        e = np.concatenate(segs)
        e = e[e[:, 1] != e[:, 3]]
        if not len(e):
            continue
        lo, hi = int(max(0, np.floor(e[:, [1, 3]].min()))), int(min(H, np.ceil(e[:, [1, 3]].max())))
        rgba = np.array([(int(sc.fill_info[p]) >> s) & 255 for s in (0, 8, 16, 24)], np.uint8)
        for row in range(lo, hi):
            y = yc[row]
            hit = ((e[:, 1] <= y) & (y < e[:, 3])) | ((e[:, 3] <= y) & (y < e[:, 1]))
            if not hit.any():
                continue
            h = e[hit]
            xs = h[:, 0] + (y - h[:, 1]) * (h[:, 2] - h[:, 0]) / (h[:, 3] - h[:, 1])
            d = np.where(h[:, 3] > h[:, 1], 1, -1)
            o = np.argsort(xs)
            xs, wn = xs[o], np.cumsum(d[o])
            inside = (wn % 2 != 0) if sc.fill_rule[p] == 1 else (wn != 0)
            for k in np.nonzero(inside[:-1])[0]:
                a, b = int(np.ceil(xs[k] - 0.5)), int(np.ceil(xs[k + 1] - 0.5))
                a, b = min(max(a, 0), W), min(max(b, 0), W)
                if b > a:
                    img[H - 1 - row, a:b] = rgba
    return img
