"""Scene containers, binary scene files and the synthetic scene generators of BASELINE.json.

`Scene` holds exactly the seven flat arrays that the reference's
ScanlineVGRasterizer::loadVG produces and uploads
(VkScanlinePR/src/core/scanline/scanline_rasterizer.cpp:67-171): they are what crosses the
C-ABI boundary in slpr_load_scene(). `Container` mirrors Galaxysailing::VGContainer
(VkScanlinePR/src/core/vg/vg_container.h:21-87), the input of loadVG.
"""
import hashlib
import struct
from dataclasses import dataclass

import numpy as np

LINE, QUADRIC, CUBIC, ARC = 0x02, 0x03, 0x04, 0x13
NON_ZERO, EVEN_ODD = 0, 1


@dataclass
class Scene:
    pos: np.ndarray            # float32 [n_points, 2]
    pos_path: np.ndarray       # uint32  [n_points]
    curve_pos_map: np.ndarray  # uint32  [n_curves] first point of the curve
    curve_type: np.ndarray     # uint32  [n_curves] 2 = line, 4 = cubic
    curve_path: np.ndarray     # uint32  [n_curves]
    fill_rule: np.ndarray      # uint32  [n_paths] 0 = nonzero, 1 = even-odd
    fill_info: np.ndarray      # uint32  [n_paths] RGBA8, R in the low byte
    name: str = "scene"
    curve_weight = None        # float32 [n_curves] or None: middle weight of ARC curves (full RVG loader)

    def __post_init__(self):
        self.pos = np.ascontiguousarray(self.pos, dtype=np.float32).reshape(-1, 2)
        for f in ("pos_path", "curve_pos_map", "curve_type", "curve_path", "fill_rule", "fill_info"):
            setattr(self, f, np.ascontiguousarray(getattr(self, f), dtype=np.uint32).reshape(-1))

    n_points = property(lambda s: int(s.pos.shape[0]))
    n_curves = property(lambda s: int(s.curve_type.shape[0]))
    n_paths = property(lambda s: int(s.fill_rule.shape[0]))

    def arrays(self):
        return (self.pos, self.pos_path, self.curve_pos_map, self.curve_type, self.curve_path,
                self.fill_rule, self.fill_info)

    def sha256(self):
        h = hashlib.sha256()
        h.update(struct.pack("<III", self.n_points, self.n_curves, self.n_paths))
        for a in self.arrays():
            h.update(a.tobytes())
        return h.hexdigest()

    def save(self, path):
        with open(path, "wb") as f:
            f.write(b"SLPR1\0\0\0")
            f.write(struct.pack("<III", self.n_points, self.n_curves, self.n_paths))
            for a in self.arrays():
                f.write(a.tobytes())

    @staticmethod
    def load(path, name=None):
        with open(path, "rb") as f:
            assert f.read(8) == b"SLPR1\0\0\0", "not a .slpr scene"
            npnt, nc, npath = struct.unpack("<III", f.read(12))
            rd = lambda n, dt: np.frombuffer(f.read(n * np.dtype(dt).itemsize), dtype=dt).copy()
            return Scene(rd(2 * npnt, np.float32), rd(npnt, np.uint32), rd(nc, np.uint32),
                         rd(nc, np.uint32), rd(nc, np.uint32), rd(npath, np.uint32),
                         rd(npath, np.uint32), name or path)

    def with_fill_rule(self, rule):
        s = Scene(*[a.copy() for a in self.arrays()], name=self.name + f"-rule{rule}")
        s.fill_rule[:] = rule
        return s


@dataclass
class Container:
    """Galaxysailing::VGContainer as plain arrays (vg_container.h:21-87)."""
    vp: np.ndarray           # float32[4]
    win: np.ndarray          # float32[4]
    pos: np.ndarray          # float32[n_points,2]
    curve_pos: np.ndarray    # uint32[n_curves]  CurveData::posIndices
    curve_type: np.ndarray   # uint32[n_curves]
    path_curve: np.ndarray   # uint32[n_paths]   PathData::curveIndices
    fill_rule: np.ndarray    # uint32[n_paths]
    fill_color: np.ndarray   # float32[n_paths,4]
    fill_opacity: np.ndarray  # float32[n_paths]

    @staticmethod
    def from_vgc(path):
        """Read the dump written by oracle/_ref/rvg_dump (the reference's own parser)."""
        with open(path, "rb") as f:
            assert f.read(4) == b"VGC1"
            npnt, nc, npath = struct.unpack("<III", f.read(12))
            rd = lambda n, dt: np.frombuffer(f.read(n * np.dtype(dt).itemsize), dtype=dt).copy()
            vp = rd(4, np.float32); win = rd(4, np.float32)
            return Container(vp, win, rd(2 * npnt, np.float32).reshape(-1, 2), rd(nc, np.uint32),
                             rd(nc, np.uint32), rd(npath, np.uint32), rd(npath, np.uint32),
                             rd(4 * npath, np.float32).reshape(-1, 4), rd(npath, np.float32))

    def to_npz(self, path):
        np.savez_compressed(path, vp=self.vp, win=self.win, pos=self.pos, curve_pos=self.curve_pos,
                            curve_type=self.curve_type, path_curve=self.path_curve,
                            fill_rule=self.fill_rule, fill_color=self.fill_color,
                            fill_opacity=self.fill_opacity)

    @staticmethod
    def from_npz(path):
        z = np.load(path)
        return Container(*[z[k] for k in ("vp", "win", "pos", "curve_pos", "curve_type", "path_curve",
                                          "fill_rule", "fill_color", "fill_opacity")])


# --------------------------------------------------------------------------- matrices
def identity_rows():
    return np.eye(4, dtype=np.float32)


def fit_rows(vp, width, height, centred=True):
    """Uniform fit of the scene's `viewport` box to the frame. Rows m0..m3
    as TransPosIn expects them (compute_ubo.h:9-13): x' = dot((x,y,0,1), m0)."""
    # every step in fp32, in the order tools/slpr_render.cpp uses, so that Python and C++ drivers agree bit for bit
    f = np.float32
    x0, y0, x1, y1 = [f(v) for v in vp]
    s = min(f(width) / (x1 - x0), f(height) / (y1 - y0))
    tx = ty = f(0)
    if centred:
        tx = (f(width) - s * (x1 - x0)) * f(0.5) - s * x0
        ty = (f(height) - s * (y1 - y0)) * f(0.5) - s * y0
    m = np.eye(4, dtype=np.float32)
    m[0, 0] = s; m[0, 3] = tx
    m[1, 1] = s; m[1, 3] = ty
    return m


def anim_rows(frame, width, height, n_frames=256):
    """cfg5: M_f = T(c) R(2*pi*f/256) S(1+0.5 sin(2*pi*f/64)) T(-c), fp64 rounded once to fp32."""
    c = np.array([width * 0.5, height * 0.5])
    a = 2 * np.pi * frame / n_frames
    s = 1 + 0.5 * np.sin(2 * np.pi * frame / 64.0)
    R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]) * s
    t = c - R @ c
    m = np.eye(4, dtype=np.float64)
    m[0, 0], m[0, 1], m[0, 3] = R[0, 0], R[0, 1], t[0]
    m[1, 0], m[1, 1], m[1, 3] = R[1, 0], R[1, 1], t[1]
    return m.astype(np.float32)


# --------------------------------------------------------------------------- synthetic scenes
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _u01(seed, path, j):
    """Deterministic U[0,1) stream: 24 high bits of splitmix64(seed, path, j)."""
    with np.errstate(over="ignore"):
        k = np.uint64(seed) * np.uint64(0x100000001B3) + path.astype(np.uint64) * np.uint64(64) + np.uint64(j)
        return (_splitmix64(k) >> np.uint64(40)).astype(np.float64) / float(1 << 24)


def _hash32(x):
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        return (_splitmix64(x ^ np.uint64(0xC0FFEE)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def synth_scene(n_paths, width, height, rmin=6.0, rmax=30.0, seed=0x5CA71E01, name=None):
    """`n_paths` closed 4-segment paths; segments alternate cubic / quadratic,
    the quadratics degree-elevated to cubics on the host in fp32 (the reference has no quadratic
    arithmetic, so elevation defines the expected result); fill rule = path & 1; opaque hashed colour."""
    P = int(n_paths)
    p = np.arange(P, dtype=np.uint64)
    cx = _u01(seed, p, 0) * width
    cy = _u01(seed, p, 1) * height
    r = rmin + (rmax - rmin) * _u01(seed, p, 2)
    phase = _u01(seed, p, 3) * 2 * np.pi
    ang = np.empty((P, 4)); rad = np.empty((P, 4))
    for k in range(4):
        ang[:, k] = phase + (k + 0.5 + 0.6 * (_u01(seed, p, 4 + k) - 0.5)) * (np.pi / 2)
        rad[:, k] = r * (0.6 + 0.4 * _u01(seed, p, 8 + k))
    vx = (cx[:, None] + rad * np.cos(ang)).astype(np.float32)
    vy = (cy[:, None] + rad * np.sin(ang)).astype(np.float32)
    pts = np.empty((P, 4, 4, 2), dtype=np.float32)  # path, segment, control point, xy
    for k in range(4):
        k1 = (k + 1) % 4
        a0, a1 = ang[:, k], ang[:, k1] + (2 * np.pi if k1 == 0 else 0)
        p0 = np.stack([vx[:, k], vy[:, k]], -1)
        p3 = np.stack([vx[:, k1], vy[:, k1]], -1)
        if k % 2 == 0:  # cubic with tangential handles
            h0 = 0.25 + 0.4 * _u01(seed, p, 12 + k)
            h1 = 0.25 + 0.4 * _u01(seed, p, 16 + k)
            c1 = np.stack([vx[:, k] + r * h0 * -np.sin(a0), vy[:, k] + r * h0 * np.cos(a0)], -1).astype(np.float32)
            c2 = np.stack([vx[:, k1] - r * h1 * -np.sin(a1), vy[:, k1] - r * h1 * np.cos(a1)], -1).astype(np.float32)
        else:  # quadratic, control point pushed outward at the mid angle; elevate in fp32
            am = 0.5 * (a0 + a1)
            rq = r * (0.9 + 0.6 * _u01(seed, p, 12 + k))
            q = np.stack([cx + rq * np.cos(am), cy + rq * np.sin(am)], -1).astype(np.float32)
            two3 = np.float32(2.0) / np.float32(3.0)
            c1 = (p0 + two3 * (q - p0)).astype(np.float32)
            c2 = (p3 + two3 * (q - p3)).astype(np.float32)
        pts[:, k, 0], pts[:, k, 1], pts[:, k, 2], pts[:, k, 3] = p0, c1, c2, p3
    pos = pts.reshape(-1, 2)
    pos_path = np.repeat(np.arange(P, dtype=np.uint32), 16)
    nc = 4 * P
    curve_pos_map = (np.arange(nc, dtype=np.uint32) * 4)
    curve_type = np.full(nc, CUBIC, dtype=np.uint32)
    curve_path = np.repeat(np.arange(P, dtype=np.uint32), 4)
    fill_rule = (np.arange(P, dtype=np.uint32) & 1)
    fill_info = np.uint32(0xFF000000) | (_hash32(np.arange(P)) & np.uint32(0xFFFFFF))
    return Scene(pos, pos_path, curve_pos_map, curve_type, curve_path, fill_rule, fill_info,
                 name or f"synth_{P}p_{width}x{height}")


def synth_1m_4k():
    return synth_scene(262144, 3840, 2160, 6.0, 30.0, 0x5CA71E01, "synth_1m_4k")


def synth_16k(n_paths=1048576):
    return synth_scene(n_paths, 16384, 16384, 8.0, 64.0, 0x5CA71E02, "synth_16k")


def flatten_reference(c: Container, name="scene"):
    """numpy restatement of ScanlineVGRasterizer::loadVG (SR.cpp:67-171) used by tests to pin the
    C++ flattening in csrc/host_scene.cpp. Colour: a*=opacity; *255; truncate to u8; whole word 0 if
    the alpha byte is 0 (SR.cpp:107-118)."""
    P = len(c.path_curve); nc = len(c.curve_pos); npnt = c.pos.shape[0]
    col = c.fill_color.astype(np.float32).copy()
    col[:, 3] = col[:, 3] * c.fill_opacity.astype(np.float32)
    col = col * np.float32(255.0)
    b = np.clip(np.trunc(col), 0, 255).astype(np.uint32) if P else np.zeros((0, 4), np.uint32)
    word = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16) | (b[:, 3] << 24)
    word = np.where(word & np.uint32(0xFF000000), word, 0).astype(np.uint32)
    curve_end = np.append(c.path_curve[1:], nc).astype(np.int64)
    curve_path = np.zeros(nc, dtype=np.uint32)
    for pi in range(P):
        curve_path[int(c.path_curve[pi]):int(curve_end[pi])] = pi
    point_end = np.append(c.curve_pos[1:], npnt).astype(np.int64)
    pos_path = np.zeros(npnt, dtype=np.uint32)
    for ci in range(nc):
        pos_path[int(c.curve_pos[ci]):int(point_end[ci])] = curve_path[ci]
    return Scene(c.pos, pos_path, c.curve_pos, c.curve_type, curve_path, c.fill_rule, word, name)
