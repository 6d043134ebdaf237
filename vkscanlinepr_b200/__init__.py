"""vkscanlinepr_b200 — B200-native scanline path renderer behind the reference's renderer interface.

The product is libslpr.so (hand-written sm_100a CUDA kernels + a C ABI, include/slpr.h). This
package is only the Python plumbing over that ABI, used by the tests and bench.py:
`ScanlineRasterizer` mirrors Galaxysailing::VGRasterizer (VkScanlinePR/src/core/rasterizer.h:9-20:
initialize / loadVG / setMVP / render) plus the headless readback. The C++ mirror of the same
interface is include/slpr_rasterizer.hpp.

There is no CPU fallback: if libslpr.so is missing or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import scene as scene  # noqa: F401
from .scene import Scene, Container

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SLPR_LIB") or os.path.join(_HERE, "libslpr.so")  # SLPR_LIB: tuning variants
_LIB = None

FLAG_TAPS, FLAG_CONTRACT_FMA, FLAG_NO_GRAPH, FLAG_RADIX_SORT, FLAG_SEGMENTED_SORT = 1, 2, 4, 8, 16
FLAG_FUSED_FILL, FLAG_SEPARATE_FILL, FLAG_WINDOWED_WALK, FLAG_RECORDS, FLAG_FULL_RVG, FLAG_AA4, FLAG_NO_LONG_WALK = 32, 64, 128, 256, 512, 1024, 2048
FLAG_BLEND = 1 << 12          # f-3: translucent fills composited source-over in path order (integer arithmetic of the oracle)
TAPS = dict(transformed_pos=0, path_visible=1, cut_cache=2, curve_count=3, curve_offset=4, intersection=5,
            key=6, path=7, winding=8, sorted_key=9, sorted_index=10, winding_scan=11, flags=12,
            flag_scan=13, records=14, segments=15)
STAGES = ("transform", "monotonize", "scan1", "piece_setup", "walk", "sort_hist", "sort_passes",
          "wind_scan", "span_emit", "fill_cells", "resolve")


class SlprError(RuntimeError):
    pass


class SlprRetry(SlprError):
    """SLPR_ERR_RETRY: an exact-band frame was void on some band; render it again on every band (new frame_seq)."""


def build(force=False, verbose=False):
    """Compile libslpr.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh", ".cpp"))]
    srcs.append(os.path.join(_HERE, "..", "include", "slpr.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        before = os.path.getmtime(LIB_PATH) if os.path.exists(LIB_PATH) else 0.0
        r = subprocess.run(["make", "-C", src_dir] + (["-B"] if force else []) + ["../libslpr.so"], capture_output=True, text=True)
        if r.returncode != 0:
            raise SlprError("building libslpr.so failed:\n" + r.stdout + r.stderr)
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) <= before:
            raise SlprError("libslpr.so is older than its sources but `make` did not rebuild it "
                            "(a source file missing from csrc/Makefile's dependencies?):\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stdout)
    return LIB_PATH


def lib():
    """Load libslpr.so (never builds implicitly on a GPU box: the .so travels with the repo)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise SlprError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                            "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.slpr_last_error.restype = C.c_char_p
        L.slpr_version.restype = C.c_char_p
        L.slpr_create.restype = C.c_void_p
        L.slpr_create.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
        L.slpr_destroy.argtypes = [C.c_void_p]
        L.slpr_launch_count.restype = C.c_uint64
        L.slpr_launch_count.argtypes = [C.c_void_p]
        L.slpr_pipeline_redone.restype = C.c_uint64
        L.slpr_pipeline_redone.argtypes = [C.c_void_p]
        L.slpr_vg_load_rvg.restype = C.c_void_p
        L.slpr_vg_load_rvg.argtypes = [C.c_char_p]
        L.slpr_vg_load_rvg_full.restype = C.c_void_p
        L.slpr_vg_load_rvg_full.argtypes = [C.c_char_p]
        L.slpr_vg_from_arrays.restype = C.c_void_p
        L.slpr_vg_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise (SlprRetry if rc == 6 else SlprError)(f"slpr error {rc}: {lib().slpr_last_error().decode()}")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class _SceneView(C.Structure):
    _fields_ = [("pos_xy", C.POINTER(C.c_float)), ("pos_path", C.POINTER(C.c_uint32)), ("n_points", C.c_uint32),
                ("curve_pos_map", C.POINTER(C.c_uint32)), ("curve_type", C.POINTER(C.c_uint32)),
                ("curve_path", C.POINTER(C.c_uint32)), ("n_curves", C.c_uint32),
                ("fill_rule", C.POINTER(C.c_uint32)), ("fill_rgba8", C.POINTER(C.c_uint32)), ("n_paths", C.c_uint32),
                ("viewport", C.c_float * 4), ("window", C.c_float * 4), ("curve_weight", C.POINTER(C.c_float))]


def _np_from(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def _flatten_handle(h, name):
    L = lib()
    v = _SceneView()
    _check(L.slpr_vg_flatten(C.c_void_p(h), C.byref(v)))
    sc = Scene(_np_from(v.pos_xy, 2 * v.n_points, np.float32), _np_from(v.pos_path, v.n_points, np.uint32),
               _np_from(v.curve_pos_map, v.n_curves, np.uint32), _np_from(v.curve_type, v.n_curves, np.uint32),
               _np_from(v.curve_path, v.n_curves, np.uint32), _np_from(v.fill_rule, v.n_paths, np.uint32),
               _np_from(v.fill_rgba8, v.n_paths, np.uint32), name)
    sc.curve_weight = _np_from(v.curve_weight, v.n_curves, np.float32) if v.curve_weight else None
    return sc, np.array(list(v.viewport), dtype=np.float32)


def container_from_handle(h):
    L = lib()
    pos = C.POINTER(C.c_float)(); cpos = C.POINTER(C.c_uint32)(); ctype = C.POINTER(C.c_uint32)()
    pcur = C.POINTER(C.c_uint32)(); frule = C.POINTER(C.c_uint32)(); fcol = C.POINTER(C.c_float)()
    fop = C.POINTER(C.c_float)()
    n_pts = C.c_uint32(); n_cur = C.c_uint32(); n_path = C.c_uint32()
    _check(L.slpr_vg_container(C.c_void_p(h), C.byref(pos), C.byref(n_pts), C.byref(cpos), C.byref(ctype), C.byref(n_cur),
                               C.byref(pcur), C.byref(frule), C.byref(fcol), C.byref(fop), C.byref(n_path)))
    return Container(np.zeros(4, np.float32), np.zeros(4, np.float32),
                     _np_from(pos, 2 * n_pts.value, np.float32).reshape(-1, 2),
                     _np_from(cpos, n_cur.value, np.uint32), _np_from(ctype, n_cur.value, np.uint32),
                     _np_from(pcur, n_path.value, np.uint32), _np_from(frule, n_path.value, np.uint32),
                     _np_from(fcol, 4 * n_path.value, np.float32).reshape(-1, 4),
                     _np_from(fop, n_path.value, np.float32))


def load_rvg(path, name=None, full=False):
    """RVG file -> (Scene, viewport, Container) through the library's parser: the reference parser's behaviour
    (rvg.cpp:9-255), or with full=True the complete reader (slpr_vg_load_rvg_full: arcs, quadratics, transforms;
    the scene then carries curve_weight and needs FLAG_FULL_RVG)."""
    L = lib()
    h = (L.slpr_vg_load_rvg_full if full else L.slpr_vg_load_rvg)(os.fsencode(path))
    if not h:
        raise SlprError(L.slpr_last_error().decode())
    try:
        sc, vp = _flatten_handle(h, name or os.path.splitext(os.path.basename(path))[0])
        cont = container_from_handle(h)
        cont.vp = vp
    finally:
        L.slpr_vg_free(C.c_void_p(h))
    return sc, vp, cont


def flatten(cont: Container, name="scene"):
    """Container -> Scene through the library's loadVG flattening (scanline_rasterizer.cpp:67-118)."""
    L = lib()
    pos = np.ascontiguousarray(cont.pos, np.float32)
    cp = np.ascontiguousarray(cont.curve_pos, np.uint32); ct = np.ascontiguousarray(cont.curve_type, np.uint32)
    pc = np.ascontiguousarray(cont.path_curve, np.uint32); fr = np.ascontiguousarray(cont.fill_rule, np.uint32)
    fc = np.ascontiguousarray(cont.fill_color, np.float32); fo = np.ascontiguousarray(cont.fill_opacity, np.float32)
    h = L.slpr_vg_from_arrays(_p(pos), C.c_uint32(pos.shape[0]), _p(cp), _p(ct), C.c_uint32(len(cp)),
                              _p(pc), _p(fr), _p(fc), _p(fo), C.c_uint32(len(pc)))
    if not h:
        raise SlprError(L.slpr_last_error().decode())
    try:
        sc, _ = _flatten_handle(h, name)
    finally:
        L.slpr_vg_free(C.c_void_p(h))
    return sc


class ScanlineRasterizer:
    """Mirror of Galaxysailing::ScanlineVGRasterizer's public interface over the C ABI.

    initialize(window, w, h) / loadVG(scene) / setMVP(rows) / render() as in
    VkScanlinePR/src/core/rasterizer.h:9-20 and app/vg_app.cpp:146-165, plus readback().
    """

    def __init__(self, device=0, flags=0):
        self._h = None
        self._device = device
        self._flags = flags
        self._stream_ptr = 0  # 0: the context's own stream
        self.width = self.height = 0

    # -- VGRasterizer ------------------------------------------------------------------------
    def initialize(self, window, w, h):
        assert window is None, "headless: pass window=None"
        self.close()
        h_ = lib().slpr_create(self._device, w, h, self._flags)
        if not h_:
            raise SlprError(lib().slpr_last_error().decode())
        self._h = C.c_void_p(h_)
        self.width, self.height = int(w), int(h)
        return self

    def loadVG(self, sc):
        if isinstance(sc, Container):
            sc = flatten(sc)
        self._scene = sc  # keep the arrays alive during the call
        _check(lib().slpr_load_scene(self._h, _p(sc.pos), _p(sc.pos_path), C.c_uint32(sc.n_points),
                                     _p(sc.curve_pos_map), _p(sc.curve_type), _p(sc.curve_path), C.c_uint32(sc.n_curves),
                                     _p(sc.fill_rule), _p(sc.fill_info), C.c_uint32(sc.n_paths)))
        w = getattr(sc, "curve_weight", None)
        if (self._flags & FLAG_FULL_RVG) and w is not None:
            w = np.ascontiguousarray(w, dtype=np.float32)
            _check(lib().slpr_set_curve_weights(self._h, _p(w), C.c_uint32(sc.n_curves)))

    def setMVP(self, rows):
        r = np.ascontiguousarray(rows, dtype=np.float32).reshape(16)
        _check(lib().slpr_set_mvp(self._h, _p(r)))

    def render(self):
        _check(lib().slpr_render(self._h))

    def prepare(self):
        _check(lib().slpr_prepare(self._h))

    # -- headless additions ------------------------------------------------------------------
    def readback(self, out=None):
        if out is None:
            out = np.empty((self.height, self.width, 4), dtype=np.uint8)
        _check(lib().slpr_readback(self._h, _p(out), C.c_size_t(out.strides[0])))
        return out

    def draw_records(self, records):
        """Stage 5 alone (slpr_draw_records): draw int32 [n, 4] records in the reference's output_buf format."""
        r = np.ascontiguousarray(records, dtype=np.int32).reshape(-1, 4)
        _check(lib().slpr_draw_records(self._h, _p(r), C.c_uint64(r.shape[0])))

    def render_to_host(self, rows, out):
        r = np.ascontiguousarray(rows, dtype=np.float32).reshape(16)
        _check(lib().slpr_render_to_host(self._h, _p(r), _p(out), C.c_size_t(out.strides[0])))
        return out

    def submit_to_host(self, rows, out):
        """Pipelined render + readback: returns once enqueued; `out` is valid after wait_host()."""
        r = np.ascontiguousarray(rows, dtype=np.float32).reshape(16)
        _check(lib().slpr_submit_to_host(self._h, _p(r), _p(out), C.c_size_t(out.strides[0])))

    def wait_host(self):
        _check(lib().slpr_wait_host(self._h))

    def pipeline_redone(self):
        """Frames the pipelined path rendered twice (they outgrew buffers sized from earlier frames)."""
        return int(lib().slpr_pipeline_redone(self._h))

    def set_band(self, y0, y1):
        _check(lib().slpr_set_band(self._h, C.c_uint32(y0), C.c_uint32(y1)))

    def set_stream(self, cuda_stream_ptr):
        _check(lib().slpr_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))
        self._stream_ptr = int(cuda_stream_ptr or 0)

    def set_target(self, dev_ptr, stride_bytes):
        _check(lib().slpr_set_target(self._h, C.c_void_p(dev_ptr), C.c_size_t(stride_bytes)))

    def band_exchange_ints(self):
        n = C.c_size_t()
        _check(lib().slpr_band_exchange_ints(self._h, C.byref(n)))
        return int(n.value)

    def set_band_exchange(self, dev_sums, dev_gathered, n_bands, band):
        """Exact row bands (include/slpr.h): device pointers of the int32 buffers [3P] and [n_bands][3P]."""
        _check(lib().slpr_set_band_exchange(self._h, C.c_void_p(dev_sums), C.c_void_p(dev_gathered), int(n_bands), int(band)))

    def render_band_begin(self):
        _check(lib().slpr_render_band_begin(self._h))

    def render_band_end(self):
        _check(lib().slpr_render_band_end(self._h))

    # exact bands, device-side exchange (include/slpr.h; host logic in parallel.py)
    def band_mailbox(self):
        p = C.c_void_p(); n = C.c_size_t()
        _check(lib().slpr_band_mailbox(self._h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def alloc_device(self, nbytes):
        p = C.c_void_p()
        _check(lib().slpr_alloc_device(self._h, C.c_size_t(nbytes), C.byref(p)))
        return int(p.value)

    def ipc_export(self, dev_ptr):
        h = (C.c_ubyte * 64)()
        _check(lib().slpr_ipc_export(self._h, C.c_void_p(dev_ptr), h))
        return bytes(h)

    def ipc_import(self, handle):
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        _check(lib().slpr_ipc_import(self._h, h, C.byref(p)))
        return int(p.value)

    def set_band_peers(self, n_bands, band, root, mailboxes):
        arr = (C.c_void_p * max(1, len(mailboxes)))(*[C.c_void_p(m) for m in mailboxes])
        _check(lib().slpr_set_band_peers(self._h, int(n_bands), int(band), int(root), arr))

    def render_band(self, frame_seq):
        _check(lib().slpr_render_band(self._h, C.c_uint32(frame_seq & 0xFFFFFFFF)))

    def band_push(self, frame_seq, root_band, dst_frame_ptr, dst_stride):
        _check(lib().slpr_band_push(self._h, C.c_uint32(frame_seq & 0xFFFFFFFF), int(root_band), C.c_void_p(dst_frame_ptr), C.c_size_t(dst_stride)))

    def band_wait_gather(self, frame_seq):
        _check(lib().slpr_band_wait_gather(self._h, C.c_uint32(frame_seq & 0xFFFFFFFF)))

    def diff_u32(self, dev_a, dev_b, n_words):
        n = C.c_uint64()
        _check(lib().slpr_debug_diff_u32(self._h, C.c_void_p(dev_a), C.c_void_p(dev_b), C.c_size_t(n_words), C.byref(n)))
        return int(n.value)

    def host_alloc(self, shape, dtype=np.uint8):
        """Pinned host array next to this context's GPU (slpr_host_alloc); returns (array, numa node or -1)."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p(); node = C.c_int(-1)
        _check(lib().slpr_host_alloc(self._h, C.c_size_t(nbytes), C.byref(p), C.byref(node)))
        buf = (C.c_ubyte * nbytes).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape), int(node.value)

    def framebuffer(self):
        p = C.c_void_p(); s = C.c_size_t()
        _check(lib().slpr_framebuffer(self._h, C.byref(p), C.byref(s)))
        return p.value, s.value

    def synchronize(self):
        _check(lib().slpr_synchronize(self._h))

    def counts(self):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().slpr_get_counts(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(n_fragments=a.value, n_out_frag=b.value, n_span=c.value)

    def sort_info(self):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _check(lib().slpr_sort_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(key_bits=a.value, passes=b.value, key_bytes=c.value)

    def sort_mode(self):
        m = C.c_int()
        _check(lib().slpr_sort_mode(self._h, C.byref(m)))
        return "radix" if m.value else "segmented"

    def fill_fused(self):
        m = C.c_int()
        _check(lib().slpr_fill_mode(self._h, C.byref(m)))
        return bool(m.value)

    def n_pieces(self):
        n = C.c_uint32()
        _check(lib().slpr_walk_info(self._h, C.byref(n)))
        return int(n.value)

    def long_walk_info(self):
        on = C.c_int(); n = C.c_uint32()
        _check(lib().slpr_long_walk_info(self._h, C.byref(on), C.byref(n)))
        return bool(on.value), int(n.value)

    def stage_ms(self):
        ms = (C.c_float * len(STAGES))()
        _check(lib().slpr_stage_ms(self._h, ms, len(STAGES)))
        return dict(zip(STAGES, [float(x) for x in ms]))

    def launch_count(self):
        return int(lib().slpr_launch_count(self._h))

    def tap(self, name):
        """Copy one reference-format intermediate buffer of the last frame (needs FLAG_TAPS for planes)."""
        cnt = self.counts()
        nf, no = cnt["n_fragments"], cnt["n_out_frag"] + cnt["n_span"]
        sc = self._scene
        shape, dt = {
            "transformed_pos": ((sc.n_points, 2), np.float32), "path_visible": ((sc.n_paths,), np.int32),
            "cut_cache": ((sc.n_curves, 5), np.float32), "curve_count": ((sc.n_curves,), np.int32),
            "curve_offset": ((sc.n_curves + 1,), np.int32), "intersection": ((nf, 2), np.int32),
            "key": ((nf + 1,), np.int32), "path": ((nf,), np.int32), "winding": ((nf,), np.int32),
            "sorted_key": ((nf,), np.int32), "sorted_index": ((nf,), np.int32),
            "winding_scan": ((nf + 1,), np.int32), "flags": ((2 * nf,), np.int32),
            "flag_scan": ((2 * nf + 1,), np.int32), "records": ((no, 4), np.int32),
            "segments": ((sc.n_paths + 1,), np.int32),
        }[name]
        out = np.empty(shape, dtype=dt)
        _check(lib().slpr_debug_copy(self._h, TAPS[name], _p(out), C.c_size_t(out.nbytes)))
        return out

    # stand-alone primitives on raw device pointers (torch tensors' data_ptr())
    def scan_i32(self, in_ptr, out_ptr, n):
        _check(lib().slpr_scan_i32(self._h, C.c_void_p(in_ptr), C.c_void_p(out_ptr), C.c_uint64(n)))

    def sort_pairs(self, keys_ptr, vals_ptr, keys_tmp_ptr, vals_tmp_ptr, n, key_bits):
        r = C.c_int()
        _check(lib().slpr_sort_pairs(self._h, C.c_void_p(keys_ptr), C.c_void_p(vals_ptr), C.c_void_p(keys_tmp_ptr),
                                     C.c_void_p(vals_tmp_ptr), C.c_uint64(n), C.c_uint32(key_bits), C.byref(r)))
        return r.value

    def close(self):
        if self._h:
            lib().slpr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
