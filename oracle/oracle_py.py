"""ctypes binding of oracle/liboracle.so — the CPU restatement of the reference shaders.

TEST INFRASTRUCTURE ONLY (parity status: see oracle/oracle.h).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module. The product (vkscanlinepr_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def build_ref():
    """Compile the reference's own RVG parser (oracle/_ref/rvg_dump) when /root/reference exists."""
    subprocess.check_call(["make", "-C", _HERE, "ref", "ref_driver"], stdout=subprocess.DEVNULL)
    p = os.path.join(_HERE, "_ref", "rvg_dump")
    return p if os.path.exists(p) else None


def ref_driver():
    """oracle/_ref/slpr_render_ref: the headless driver compiled against the reference's own headers and parser."""
    p = os.path.join(_HERE, "_ref", "slpr_render_ref")
    return p if os.path.exists(p) else None


class _Frame(C.Structure):
    _fields_ = [
        ("n_fragments", C.c_int32), ("n_out_frag", C.c_int32), ("n_span", C.c_int32),
        ("n_points", C.c_uint32), ("n_curves", C.c_uint32), ("n_paths", C.c_uint32),
        ("width", C.c_int), ("height", C.c_int),
        ("tpos", C.POINTER(C.c_float)), ("path_visible", C.POINTER(C.c_int32)),
        ("cut_cache", C.POINTER(C.c_float)), ("curve_count", C.POINTER(C.c_int32)),
        ("curve_offset", C.POINTER(C.c_int32)), ("inter", C.POINTER(C.c_int32)),
        ("key", C.POINTER(C.c_int32)), ("idx", C.POINTER(C.c_int32)),
        ("path", C.POINTER(C.c_int32)), ("wind", C.POINTER(C.c_int32)),
        ("seg", C.POINTER(C.c_int32)),
        ("skey", C.POINTER(C.c_int32)), ("sidx", C.POINTER(C.c_int32)),
        ("swind", C.POINTER(C.c_int32)), ("wn", C.POINTER(C.c_int32)),
        ("flags", C.POINTER(C.c_int32)), ("scan3", C.POINTER(C.c_int32)),
        ("records", C.POINTER(C.c_int32)), ("rgba", C.POINTER(C.c_uint8)),
        ("ms", C.c_double * 8),
    ]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_render.restype = C.POINTER(_Frame)
        L.orc_frame_free.argtypes = [C.POINTER(_Frame)]
        L.orc_quantise_colour.restype = C.c_uint32
        L.orc_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _arr(ptr, n, dtype):
    if n <= 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


STAGE_NAMES = ("transform", "monotonize", "scan1", "intersect", "gen_fragment", "sort", "spans", "fill")


def render(scene, rows, width, height, do_fill=True, threads=None, keep=None, full=False, fma=False, blend=False):
    """Run the whole oracle frame. `scene` has the 7 flat loadVG arrays (see scene.Scene).
    Returns a dict of every intermediate buffer (numpy copies); `keep` (a set of names) limits the copies to
    those buffers — on frames of 10^8 fragments the full set is tens of gigabytes."""
    L = lib()
    if threads:
        L.orc_set_num_threads(int(threads))
    rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(16)
    s = scene
    weights = None
    if full:  # real QUADRIC / ARC arithmetic instead of the reference's TODO arms (oracle.c, orc_set_full_rvg)
        w = getattr(s, "curve_weight", None)
        weights = np.ascontiguousarray(w if w is not None else np.ones(s.n_curves), dtype=np.float32)
        L.orc_set_full_rvg(C.c_int(1), _p(weights))
    if fma:  # the contracted reading of the shaders (oracle.c, orc_set_contract_fma)
        L.orc_set_contract_fma(C.c_int(1))
    if blend:  # source-over compositing of translucent fills in path order (oracle.c, orc_set_blend)
        L.orc_set_blend(C.c_int(1))
    try:
        fp = _call_render(L, s, rows, width, height, do_fill)
    finally:
        if full:
            L.orc_set_full_rvg(C.c_int(0), None)
        if fma:
            L.orc_set_contract_fma(C.c_int(0))
        if blend:
            L.orc_set_blend(C.c_int(0))
    return _collect(L, fp, s, width, height, do_fill, keep)


def box4(rgba_hi):
    """(4H, 4W, 4) uint8 -> (H, W, 4): every pixel the rounded average of its 2 x 2 coverage cells (each cell is 2 x 2 equal
    pixels of the 4x frame), (sum + 2) >> 2 per channel — the resolve of SLPR_FLAG_AA4."""
    h, w = rgba_hi.shape[0] // 4, rgba_hi.shape[1] // 4
    cells = rgba_hi[::2, ::2].astype(np.uint32).reshape(h, 2, w, 2, 4)
    return ((cells.sum(axis=(1, 3)) + 2) >> 2).astype(np.uint8)


def render_aa4(scene, rows, width, height, full=False, blend=False):
    """the definition of SLPR_FLAG_AA4 in terms of the reference path — the frame rendered at four times the
    size (matrix rows 0 and 1 scaled by 4, exactly) and box-filtered. Returns the dict of the 4x frame plus "rgba_aa"."""
    r4 = np.array(rows, dtype=np.float32).reshape(4, 4).copy()
    r4[0] *= np.float32(4.0)
    r4[1] *= np.float32(4.0)
    out = render(scene, r4, 4 * width, 4 * height, full=full, blend=blend, keep={"rgba"})
    out["rgba_aa"] = box4(out["rgba"])
    return out


def _call_render(L, s, rows, width, height, do_fill):
    return L.orc_render(C.c_uint32(s.n_points), _p(s.pos), _p(s.pos_path),
                      C.c_uint32(s.n_curves), _p(s.curve_pos_map), _p(s.curve_type), _p(s.curve_path),
                      C.c_uint32(s.n_paths), _p(s.fill_rule), _p(s.fill_info),
                      _p(rows), C.c_int(width), C.c_int(height), C.c_int(1 if do_fill else 0))


def _collect(L, fp, s, width, height, do_fill, keep):
    f = fp.contents
    nf = f.n_fragments
    no = f.n_out_frag + f.n_span
    out = dict(n_fragments=nf, n_out_frag=f.n_out_frag, n_span=f.n_span, ms=dict(zip(STAGE_NAMES, list(f.ms))))
    buffers = dict(
        tpos=lambda: _arr(f.tpos, 2 * s.n_points, np.float32).reshape(-1, 2),
        path_visible=lambda: _arr(f.path_visible, s.n_paths, np.int32),
        cut_cache=lambda: _arr(f.cut_cache, 5 * s.n_curves, np.float32).reshape(-1, 5),
        curve_count=lambda: _arr(f.curve_count, s.n_curves, np.int32),
        curve_offset=lambda: _arr(f.curve_offset, s.n_curves + 1, np.int32),
        inter=lambda: _arr(f.inter, 2 * nf, np.int32).reshape(-1, 2),
        key=lambda: _arr(f.key, nf, np.int32), idx=lambda: _arr(f.idx, nf, np.int32),
        path=lambda: _arr(f.path, nf, np.int32), wind=lambda: _arr(f.wind, nf, np.int32),
        seg=lambda: _arr(f.seg, s.n_paths + 1, np.int32),
        skey=lambda: _arr(f.skey, nf, np.int32), sidx=lambda: _arr(f.sidx, nf, np.int32),
        swind=lambda: _arr(f.swind, nf, np.int32), wn=lambda: _arr(f.wn, nf + 1, np.int32),
        flags=lambda: _arr(f.flags, 2 * nf, np.int32), scan3=lambda: _arr(f.scan3, 2 * nf + 1, np.int32),
        records=lambda: _arr(f.records, 4 * no, np.int32).reshape(-1, 4),
    )
    if do_fill:
        buffers["rgba"] = lambda: _arr(f.rgba, 4 * width * height, np.uint8).reshape(height, width, 4)
    for name, get in buffers.items():
        if keep is None or name in keep:
            out[name] = get()
    L.orc_frame_free(fp)
    return out


def fill(records, width, height, blend=False):
    L = lib()
    rec = np.ascontiguousarray(records, dtype=np.int32).reshape(-1, 4)
    rgba = np.empty((height, width, 4), dtype=np.uint8)
    if blend:
        L.orc_set_blend(C.c_int(1))
    try:
        L.orc_fill(C.c_int64(rec.shape[0]), _p(rec), C.c_int(width), C.c_int(height), _p(rgba))
    finally:
        if blend:
            L.orc_set_blend(C.c_int(0))
    return rgba


def seg_sort(seg, key, idx, literal=False):
    L = lib()
    seg = np.ascontiguousarray(seg, dtype=np.int32)
    key = np.array(key, dtype=np.int32, copy=True)
    idx = np.array(idx, dtype=np.int32, copy=True)
    fn = L.orc_seg_sort_literal if literal else L.orc_seg_sort
    fn(C.c_uint32(len(seg) - 1), _p(seg), _p(key), _p(idx))
    return key, idx


def exclusive_scan(a):
    L = lib()
    a = np.ascontiguousarray(a, dtype=np.int32)
    out = np.empty(a.shape[0] + 1, dtype=np.int32)
    L.orc_exclusive_scan(C.c_int64(a.shape[0]), _p(a), _p(out))
    return out


def quantise_colour(rgba, opacity):
    v = np.ascontiguousarray(rgba, dtype=np.float32)
    return int(lib().orc_quantise_colour(_p(v), C.c_float(opacity)))


def num_threads():
    return int(lib().orc_num_threads())
