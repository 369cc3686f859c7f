"""ctypes binding of libcontour3d_b200.so (include/contour3d_b200.h) and of the
drop-in entry point contour3d_execute (include/contour3d.h).

Python is only the host language of the tests and the benchmark; all compute
is in the CUDA library.  There is no fallback: if the library is missing or no
B200 is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libcontour3d_b200.so")
COMPAT_PATH = os.path.join(PKG, "libc3d_sdecomp_compat.so")

NCCL_ID_BYTES = 128

TRIANGLE_DTYPE = np.dtype(
    [
        ("cube", "<u8"),
        ("tri", "<u4"),
        ("has_vn", "<u4"),
        ("edge", "<u4", (3,)),
        ("pad", "<u4"),
        ("area", "<f8"),
        ("v", "<f8", (3, 3)),
        ("fn", "<f8", (3,)),
        ("vn", "<f8", (3, 3)),
    ]
)


class Decomp(C.Structure):
    _fields_ = [("nprocs", C.c_int * 3), ("myrank", C.c_int * 3), ("neighbours", (C.c_int * 2) * 3)]


class Scene(C.Structure):
    _fields_ = [
        ("camera", C.c_double * 3),
        ("light", C.c_double * 3),
        ("center", C.c_double * 3),
        ("local_x", C.c_double * 3),
        ("local_y", C.c_double * 3),
        ("width", C.c_size_t),
        ("height", C.c_size_t),
        ("bg", C.c_uint8 * 3),
    ]


class Positions(C.Structure):
    """c3d_positions_t: a coordinate converter tabulated over the extended local array"""
    _fields_ = [("table", C.POINTER(C.c_double) * 3), ("stride", (C.c_size_t * 3) * 3)]


class ContourDesc(C.Structure):
    _fields_ = [
        ("field", C.c_void_p),
        ("grids", C.POINTER(C.c_double) * 3),
        ("threshold", C.c_double),
        ("color", C.c_uint8 * 3),
        ("positions", C.POINTER(Positions)),
    ]


class LineDesc(C.Structure):
    _fields_ = [("edges", (C.c_double * 3) * 2), ("nitems", C.c_size_t), ("width", C.c_double), ("color", C.c_uint8 * 3),
                ("points", C.POINTER(C.c_double))]


def camera_from_angles(elevation, azimuth, roll=0.0, camera_distance=2.0, screen_distance=1.0, screen_width=16.0 / 9.0, screen_height=1.0) -> dict:
    """c3d_camera_from_angles -> dict(camera, center, local_x, local_y) (the CameraAdjuster construction); host only"""
    s = Scene()
    fn = lib().c3d_camera_from_angles
    fn.restype = C.c_int
    fn.argtypes = [C.c_double] * 7 + [C.POINTER(Scene)]
    if 0 != fn(elevation, azimuth, roll, camera_distance, screen_distance, screen_width, screen_height, C.byref(s)):
        raise ValueError("c3d_camera_from_angles: lengths must be positive")
    return {k: tuple(getattr(s, k)) for k in ("camera", "center", "local_x", "local_y")}


CONVERTER_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)


def positions_probe(convert, grids, offset, count, max_bytes=1 << 30):
    """c3d_positions_probe on a Python function (x, y, z) -> (X, Y, Z); host only, no device needed.
    Returns (identity, tables, strides): tables are numpy copies of the three component tables (None when identity)."""
    def tramp(pin, pout, _user):
        out = convert(pin[0], pin[1], pin[2])
        pout[0], pout[1], pout[2] = float(out[0]), float(out[1]), float(out[2])

    cb = CONVERTER_FN(tramp)
    gs = [np.ascontiguousarray(g, dtype=np.float64) for g in grids]
    garr = (C.POINTER(C.c_double) * 3)(*[g.ctypes.data_as(C.POINTER(C.c_double)) for g in gs])
    off = (C.c_size_t * 3)(*[int(v) for v in offset])
    cnt = (C.c_size_t * 3)(*[int(v) for v in count])
    pos = Positions()
    ident = C.c_int(1)
    fn = lib().c3d_positions_probe
    fn.restype = C.c_int
    fn.argtypes = [CONVERTER_FN, C.c_void_p, C.POINTER(C.POINTER(C.c_double) * 3), C.POINTER(C.c_size_t * 3), C.POINTER(C.c_size_t * 3),
                   C.c_size_t, C.POINTER(Positions), C.POINTER(C.c_int)]
    rc = fn(cb, None, C.byref(garr), C.byref(off), C.byref(cnt), int(max_bytes), C.byref(pos), C.byref(ident))
    if rc != 0:
        raise RuntimeError("c3d_positions_probe failed (see stderr)")
    if ident.value:
        return True, None, None
    strides = [[int(pos.stride[c][a]) for a in range(3)] for c in range(3)]
    tables = []
    for c in range(3):
        n = 1 + sum((int(count[a]) - 1) * strides[c][a] for a in range(3))
        tables.append(np.ctypeslib.as_array(pos.table[c], shape=(n,)).copy())
    free = lib().c3d_positions_free
    free.restype = None
    free.argtypes = [C.POINTER(Positions)]
    free(C.byref(pos))
    return False, tables, strides


class Timings(C.Structure):
    _STAGES = ("upload", "halo", "clear", "sweep", "expand", "raster", "lines", "resolve", "reduce", "shade", "pack", "download", "total")
    _fields_ = [(k, C.c_float) for k in _STAGES] + [
        ("sweep_each", C.c_float * 16),
        ("sweep_bytes", C.c_uint64 * 16),
        ("sweep_thresholds", C.c_uint32 * 16),
        ("active_cubes", C.c_uint64 * 16),
        ("culled_cubes", C.c_uint64),
        ("culled_triangles", C.c_uint64),
        ("shade_spills", C.c_uint64),
        ("nsweeps", C.c_uint32),
        ("ncontours", C.c_uint32),
        ("launches", C.c_uint32),
        ("exchange_mode", C.c_uint32),
        ("list_entries", C.c_uint64),
        ("detail", C.c_float * 16),
        ("tile_share", C.c_float * 8),
    ]

    def as_dict(self):
        n, m = int(self.ncontours), int(self.nsweeps)
        d = {k: float(getattr(self, k)) for k in self._STAGES}
        d["sweep_each"] = [float(self.sweep_each[i]) for i in range(m)]
        d["sweep_bytes"] = [int(self.sweep_bytes[i]) for i in range(m)]
        d["sweep_thresholds"] = [int(self.sweep_thresholds[i]) for i in range(m)]
        d["active_cubes"] = [int(self.active_cubes[i]) for i in range(n)]
        d["culled_cubes"] = int(self.culled_cubes)
        d["culled_triangles"] = int(self.culled_triangles)
        d["shade_spills"] = int(self.shade_spills)
        d["launches"] = int(self.launches)
        d["exchange_mode"] = int(self.exchange_mode)
        d["list_entries"] = int(self.list_entries)
        d["detail"] = [float(self.detail[i]) for i in range(11)]
        d["tile_share"] = [float(self.tile_share[i]) for i in range(8)]
        return d


# ---- the public drop-in structs (include/contour3d.h) ----
class Vector(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]


class Color(C.Structure):
    _fields_ = [("r", C.c_uint8), ("g", C.c_uint8), ("b", C.c_uint8)]


CONVERTER = C.CFUNCTYPE(Vector, Vector)


class ContourObj(C.Structure):
    _fields_ = [
        ("pencil", C.c_uint8),
        ("glsizes", C.c_size_t * 3),
        ("grids", C.POINTER(C.c_double) * 3),
        ("converter", C.c_void_p),
        ("threshold", C.c_double),
        ("color", Color),
        ("array", C.c_void_p),
    ]


class LineObj(C.Structure):
    _fields_ = [("edges", Vector * 2), ("nitems", C.c_size_t), ("converter", C.c_void_p), ("color", Color), ("width", C.c_double)]


assert C.sizeof(ContourObj) == 88 and C.sizeof(LineObj) == 80 and C.sizeof(Color) == 3

_lib = None
_compat = None

EXPORTS = [
    "c3d_nccl_unique_id", "c3d_context_create", "c3d_context_destroy", "c3d_context_device", "c3d_context_stream",
    "c3d_context_synchronize", "c3d_context_exchange_mode", "c3d_field_wrap", "c3d_peer_selftest", "c3d_selftest_work_plan", "c3d_selftest_tile_map", "c3d_positions_probe", "c3d_positions_free", "c3d_camera_from_angles", "c3d_frame_collectives_probe", "c3d_decomp_make", "c3d_decomp_local", "c3d_field_create", "c3d_field_destroy",
    "c3d_field_upload", "c3d_field_geometry", "c3d_field_interior", "c3d_field_mark_dirty", "c3d_field_exchange",
    "c3d_field_download_ext", "c3d_field_synthesize", "c3d_frame_render", "c3d_frame_timings", "c3d_frame_canvas",
    "c3d_case_codes", "c3d_sweep", "c3d_tessellate", "c3d_version", "contour3d_execute",
]


def compat():
    """SimpleDecomp + MPI stand-in, loaded globally so that libcontour3d_b200.so finds `sdecomp` and MPI_*."""
    global _compat
    if _compat is None:
        if not os.path.exists(COMPAT_PATH):
            raise RuntimeError(f"{COMPAT_PATH} is missing: run `python -m contour3d_b200.build`")
        _compat = C.CDLL(COMPAT_PATH, mode=C.RTLD_GLOBAL)
    return _compat


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m contour3d_b200.build` (there is no CPU fallback)")
        compat()
        L = C.CDLL(LIB_PATH)
        vp, sz, dbl, i32 = C.c_void_p, C.c_size_t, C.c_double, C.c_int
        L.c3d_version.restype = C.c_char_p
        L.c3d_nccl_unique_id.argtypes = [vp]
        L.c3d_context_create.argtypes = [C.POINTER(vp), i32, i32, i32, vp]
        L.c3d_context_destroy.argtypes = [vp]
        L.c3d_context_device.argtypes = [vp]
        L.c3d_context_stream.argtypes = [vp]
        L.c3d_context_stream.restype = C.c_size_t
        L.c3d_context_synchronize.argtypes = [vp]
        L.c3d_decomp_make.argtypes = [C.POINTER(i32 * 3), i32, C.POINTER(Decomp)]
        L.c3d_decomp_local.argtypes = [C.POINTER(Decomp), C.POINTER(sz * 3), C.POINTER(sz * 3), C.POINTER(sz * 3)]
        L.c3d_field_create.argtypes = [vp, C.POINTER(Decomp), C.POINTER(sz * 3), i32, C.POINTER(vp)]
        L.c3d_field_destroy.argtypes = [vp, vp]
        L.c3d_field_upload.argtypes = [vp, vp, vp]
        L.c3d_field_geometry.argtypes = [vp, C.POINTER(sz * 3), C.POINTER(sz * 3), C.POINTER(sz * 3), C.POINTER(sz * 3)]
        L.c3d_field_interior.argtypes = [vp, C.POINTER(vp), C.POINTER(sz * 2)]
        L.c3d_field_mark_dirty.argtypes = [vp]
        L.c3d_context_exchange_mode.argtypes = [vp]
        L.c3d_field_wrap.argtypes = [vp, C.POINTER(Decomp), C.POINTER(sz * 3), i32, vp, C.POINTER(vp)]
        L.c3d_field_exchange.argtypes = [vp, vp]
        L.c3d_field_download_ext.argtypes = [vp, vp, vp]
        L.c3d_field_synthesize.argtypes = [vp, vp, i32, C.c_uint32, i32, dbl]
        L.c3d_frame_render.argtypes = [vp, C.POINTER(Scene), sz, C.POINTER(ContourDesc), sz, C.POINTER(LineDesc), vp, C.c_char_p]
        L.c3d_frame_timings.argtypes = [vp, C.POINTER(Timings)]
        L.c3d_frame_canvas.argtypes = [vp, vp, vp]
        L.c3d_case_codes.argtypes = [vp, vp, dbl, vp, vp]
        L.c3d_sweep.argtypes = [vp, vp, dbl, i32, sz, vp, C.POINTER(C.c_uint64)]
        L.c3d_tessellate.argtypes = [vp, vp, C.POINTER(C.POINTER(C.c_double) * 3), dbl, sz, vp, C.POINTER(C.c_uint64)]
        L.contour3d_execute.argtypes = [
            vp, C.POINTER(Vector), C.POINTER(Vector), C.POINTER(Vector), C.POINTER(sz * 2), C.POINTER(Vector),
            C.POINTER(Vector * 2), C.POINTER(Color), sz, C.POINTER(ContourObj), sz, C.POINTER(LineObj), C.c_char_p,
        ]
        _lib = L
    return _lib


class C3DError(RuntimeError):
    pass


def _check(rc, what):
    if rc != 0:
        raise C3DError(f"{what} failed (see the [CONTOUR3D ERROR] line on stderr)")


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    _check(lib().c3d_nccl_unique_id(buf), "c3d_nccl_unique_id")
    return buf.raw


def make_decomp(dims=(1, 1, 1), rank=0) -> Decomp:
    d = Decomp()
    _check(lib().c3d_decomp_make(C.byref((C.c_int * 3)(*dims)), int(rank), C.byref(d)), "c3d_decomp_make")
    return d


def decomp_local(decomp: Decomp, glsizes):
    my, off = (C.c_size_t * 3)(), (C.c_size_t * 3)()
    _check(lib().c3d_decomp_local(C.byref(decomp), C.byref((C.c_size_t * 3)(*glsizes)), C.byref(my), C.byref(off)), "c3d_decomp_local")
    return tuple(my), tuple(off)


class Field:
    """A device-resident local array with halo room (c3d_field_t)."""

    def __init__(self, ctx: "Context", glsizes, dtype=np.float32, decomp: Decomp | None = None):
        self.ctx = ctx
        self.glsizes = tuple(int(g) for g in glsizes)  # (nx, ny, nz)
        self.dtype = np.dtype(dtype)
        assert self.dtype in (np.dtype(np.float32), np.dtype(np.float64))
        self.decomp = decomp if decomp is not None else ctx.decomp
        h = C.c_void_p()
        _check(lib().c3d_field_create(ctx.handle, C.byref(self.decomp), C.byref((C.c_size_t * 3)(*self.glsizes)),
                                      self.dtype.itemsize, C.byref(h)), "c3d_field_create")
        self.handle = h
        ext, eoff, my, off = ((C.c_size_t * 3)() for _ in range(4))
        _check(lib().c3d_field_geometry(h, C.byref(ext), C.byref(eoff), C.byref(my), C.byref(off)), "c3d_field_geometry")
        self.ext, self.ext_off, self.mysizes, self.offsets = tuple(ext), tuple(eoff), tuple(my), tuple(off)

    def upload(self, local):
        """local: numpy array (mz, my, mx) of this field's dtype, or an int device pointer."""
        if isinstance(local, np.ndarray):
            a = np.ascontiguousarray(local, dtype=self.dtype)
            assert a.shape == (self.mysizes[2], self.mysizes[1], self.mysizes[0]), (a.shape, self.mysizes)
            self._keep = a
            ptr = a.ctypes.data
        else:
            ptr = int(local)
        _check(lib().c3d_field_upload(self.ctx.handle, self.handle, ptr), "c3d_field_upload")
        return self

    def upload_global(self, global_array: np.ndarray):
        """Convenience: slice this rank's pencil out of a global (nz, ny, nx) array."""
        o, m = self.offsets, self.mysizes
        return self.upload(global_array[o[2]:o[2] + m[2], o[1]:o[1] + m[1], o[0]:o[0] + m[0]])

    def synthesize(self, kind: int, seed: int = 1234, octaves: int = 5, gain: float = 0.5):
        _check(lib().c3d_field_synthesize(self.ctx.handle, self.handle, kind, seed, octaves, gain), "c3d_field_synthesize")
        return self

    def exchange(self):
        _check(lib().c3d_field_exchange(self.ctx.handle, self.handle), "c3d_field_exchange")
        return self

    def mark_dirty(self):
        """the interior changed (or is to be treated so): the next frame exchanges the halos again"""
        _check(lib().c3d_field_mark_dirty(self.handle), "c3d_field_mark_dirty")
        return self

    def download_ext(self) -> np.ndarray:
        out = np.empty((self.ext[2], self.ext[1], self.ext[0]), dtype=self.dtype)
        _check(lib().c3d_field_download_ext(self.ctx.handle, self.handle, out.ctypes.data), "c3d_field_download_ext")
        return out

    def destroy(self):
        if self.handle:
            lib().c3d_field_destroy(self.ctx.handle, self.handle)
            self.handle = None


class Context:
    """Device context (c3d_context_t): one per process / GPU."""

    def __init__(self, device=-1, rank=0, size=1, nccl_id: bytes | None = None, dims=None):
        h = C.c_void_p()
        idbuf = C.create_string_buffer(nccl_id, NCCL_ID_BYTES) if nccl_id is not None else None
        _check(lib().c3d_context_create(C.byref(h), int(device), int(rank), int(size), idbuf), "c3d_context_create")
        self.handle = h
        self.rank, self.size = rank, size
        self.dims = tuple(dims) if dims is not None else (1, 1, size)
        self.decomp = make_decomp(self.dims, rank)

    @property
    def stream(self) -> int:
        return int(lib().c3d_context_stream(self.handle))

    @property
    def device(self) -> int:
        return int(lib().c3d_context_device(self.handle))

    def synchronize(self):
        _check(lib().c3d_context_synchronize(self.handle), "c3d_context_synchronize")

    def field(self, glsizes, dtype=np.float32) -> Field:
        return Field(self, glsizes, dtype)

    # -- whole frame --
    def render(self, scene: dict, contours, lines=(), want_rgb=True, fname: str | None = None):
        """contours: list of dict(field=Field, grids=[gx, gy, gz] (global), threshold, color)
        lines: list of dict(edges, nitems, width, color)
        returns the y-flipped image (H, W, 3) uint8 on rank 0 (None elsewhere / when want_rgb is False)."""
        # a movie loop renders the same objects again and again: the C structures of the previous call are kept and reused
        # when the very same Python objects come back (identity, not equality: cheap to check, and nothing here is mutated)
        sig = (id(scene), tuple(sorted(scene.items())) if isinstance(scene, dict) else None,
               tuple((id(c["field"]), id(c["grids"]), float(c["threshold"]), tuple(c.get("color", (255, 255, 255))), id(c.get("positions"))) for c in contours),
               tuple((tuple(map(tuple, l["edges"])), int(l["nitems"]), float(l["width"]), tuple(l.get("color", (255, 255, 255))), id(l.get("points"))) for l in lines))
        cached = getattr(self, "_render_cache", None)
        if cached is not None and cached[0] == sig:
            s, carr, larr, keep = cached[1]
            return self._render_call(s, carr, larr, len(contours), len(lines), want_rgb, fname)
        s = Scene()
        for k in ("camera", "light", "center", "local_x", "local_y"):
            getattr(s, k)[:] = scene[k]
        s.width, s.height = int(scene["width"]), int(scene["height"])
        s.bg[:] = list(scene.get("bg", (0, 0, 0)))
        keep = [scene, contours, lines]
        carr = (ContourDesc * max(1, len(contours)))()
        for n, c in enumerate(contours):
            carr[n].field = c["field"].handle
            for d in range(3):
                g = np.ascontiguousarray(c["grids"][d], dtype=np.float64)
                assert len(g) == c["field"].glsizes[d]
                keep.append(g)
                carr[n].grids[d] = g.ctypes.data_as(C.POINTER(C.c_double))
            carr[n].threshold = float(c["threshold"])
            carr[n].color[:] = list(c.get("color", (255, 255, 255)))
            if c.get("positions") is not None:
                # (tables, strides) over the field's EXTENDED local array, e.g. from positions_probe
                tables, strides = c["positions"]
                pos = Positions()
                for comp in range(3):
                    t = np.ascontiguousarray(tables[comp], dtype=np.float64)
                    keep.append(t)
                    pos.table[comp] = t.ctypes.data_as(C.POINTER(C.c_double))
                    for a in range(3):
                        pos.stride[comp][a] = int(strides[comp][a])
                keep.append(pos)
                carr[n].positions = C.pointer(pos)
        larr = (LineDesc * max(1, len(lines)))()
        for n, l in enumerate(lines):
            larr[n].edges[0][:] = l["edges"][0]
            larr[n].edges[1][:] = l["edges"][1]
            larr[n].nitems = int(l["nitems"])
            larr[n].width = float(l["width"])
            larr[n].color[:] = list(l.get("color", (255, 255, 255)))
            if l.get("points") is not None:
                pts = np.ascontiguousarray(l["points"], dtype=np.float64).reshape(int(l["nitems"]), 3)
                keep.append(pts)
                larr[n].points = pts.ctypes.data_as(C.POINTER(C.c_double))
        self._render_cache = (sig, (s, carr, larr, keep))
        return self._render_call(s, carr, larr, len(contours), len(lines), want_rgb, fname)

    def _render_call(self, s, carr, larr, nc, nl, want_rgb, fname):
        out = None
        ptr = None
        if want_rgb and self.rank == 0:
            out = np.empty((s.height, s.width, 3), dtype=np.uint8)
            ptr = out.ctypes.data
        _check(lib().c3d_frame_render(self.handle, C.byref(s), nc, carr, nl, larr, ptr,
                                      fname.encode() if fname else None), "c3d_frame_render")
        return out

    def collectives_probe(self, width, height, iters=10):
        """(ms per depth all-reduce, ms per colour reduce) of a width x height image on this context's communicator; collective"""
        a, b = C.c_float(0), C.c_float(0)
        fn = lib().c3d_frame_collectives_probe
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _check(fn(self.handle, int(width), int(height), int(iters), C.byref(a), C.byref(b)), "c3d_frame_collectives_probe")
        return float(a.value), float(b.value)

    def timings(self) -> dict:
        t = Timings()
        _check(lib().c3d_frame_timings(self.handle, C.byref(t)), "c3d_frame_timings")
        return t.as_dict()

    def canvas(self, width, height):
        depth = np.empty((height, width), dtype=np.float64)
        key = np.empty((height, width), dtype=np.uint64)
        _check(lib().c3d_frame_canvas(self.handle, depth.ctypes.data, key.ctypes.data), "c3d_frame_canvas")
        return depth, key

    # -- stage level --
    def case_codes(self, field: Field, threshold: float):
        shape = (field.ext[2] - 1, field.ext[1] - 1, field.ext[0] - 1)
        codes = np.empty(shape, dtype=np.uint32)
        counts = np.empty(shape, dtype=np.uint8)
        _check(lib().c3d_case_codes(self.handle, field.handle, float(threshold), codes.ctypes.data, counts.ctypes.data), "c3d_case_codes")
        return codes, counts

    def sweep(self, field: Field, threshold: float, rendered_only=True) -> np.ndarray:
        n = C.c_uint64(0)
        _check(lib().c3d_sweep(self.handle, field.handle, float(threshold), int(rendered_only), 0, None, C.byref(n)), "c3d_sweep")
        out = np.empty(int(n.value), dtype=np.uint64)
        if n.value:
            _check(lib().c3d_sweep(self.handle, field.handle, float(threshold), int(rendered_only), out.size, out.ctypes.data, C.byref(n)), "c3d_sweep")
        return out

    def tessellate(self, field: Field, grids, threshold: float) -> np.ndarray:
        g = [np.ascontiguousarray(x, dtype=np.float64) for x in grids]
        garr = (C.POINTER(C.c_double) * 3)(*[x.ctypes.data_as(C.POINTER(C.c_double)) for x in g])
        n = C.c_uint64(0)
        _check(lib().c3d_tessellate(self.handle, field.handle, C.byref(garr), float(threshold), 0, None, C.byref(n)), "c3d_tessellate")
        out = np.zeros(int(n.value), dtype=TRIANGLE_DTYPE)
        if n.value:
            _check(lib().c3d_tessellate(self.handle, field.handle, C.byref(garr), float(threshold), out.size, out.ctypes.data, C.byref(n)), "c3d_tessellate")
        return out

    def destroy(self):
        if self.handle:
            lib().c3d_context_destroy(self.handle)
            self.handle = None


# ---------------------------------------------------------------------------
# the drop-in call, exactly as a C host would make it (host buffers, PPM file)
# ---------------------------------------------------------------------------

_sdecomp_info = {}


def sdecomp_construct(dims=(0, 0, 0)):
    """sdecomp.construct(MPI_COMM_WORLD, 3, dims, periods=false) through the stand-in's method table."""
    key = tuple(dims)
    if key in _sdecomp_info:
        return _sdecomp_info[key]
    cp = compat()
    cp.MPI_Init(None, None)

    class SDecompT(C.Structure):
        _fields_ = [(name, C.c_void_p) for name in (
            "construct", "destruct", "get_ndims", "get_comm_size", "get_comm_rank", "get_nprocs", "get_myrank",
            "get_neighbours", "get_pencil_mysize", "get_pencil_offset", "get_comm_cart", "t_construct", "t_execute", "t_destruct")]

    table = SDecompT.in_dll(cp, "sdecomp")
    construct = C.CFUNCTYPE(C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_size_t * 3), C.POINTER(C.c_bool * 3), C.POINTER(C.c_void_p))(table.construct)
    info = C.c_void_p()
    rc = construct(1, 3, C.byref((C.c_size_t * 3)(*dims)), C.byref((C.c_bool * 3)(False, False, False)), C.byref(info))
    if rc != 0:
        raise C3DError("sdecomp.construct failed")
    _sdecomp_info[key] = info
    return info


def execute(scene: dict, grids, arrays, contours, lines, fname: str, sdecomp_info=None):
    """contour3d_execute(...) with host buffers, as the reference's caller would.

    arrays: list of float64 (nz, ny, nx) numpy arrays (local pencil == global for one rank), or int device pointers
    contours: list of dict(field=index into arrays, threshold, color)
    """
    L = lib()
    info = sdecomp_info if sdecomp_info is not None else sdecomp_construct()
    g = [np.ascontiguousarray(x, dtype=np.float64) for x in grids]
    nx, ny, nz = len(g[0]), len(g[1]), len(g[2])
    cobjs = (ContourObj * max(1, len(contours)))()
    for n, c in enumerate(contours):
        a = arrays[c.get("field", 0)]
        cobjs[n].pencil = 0
        cobjs[n].glsizes[:] = [nx, ny, nz]
        for d in range(3):
            cobjs[n].grids[d] = g[d].ctypes.data_as(C.POINTER(C.c_double))
        cobjs[n].converter = None
        cobjs[n].threshold = float(c["threshold"])
        cobjs[n].color = Color(*c.get("color", (255, 255, 255)))
        if isinstance(a, np.ndarray):
            assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
            cobjs[n].array = a.ctypes.data
        else:
            cobjs[n].array = int(a)
    lobjs = (LineObj * max(1, len(lines)))()
    for n, l in enumerate(lines):
        lobjs[n].edges[0] = Vector(*l["edges"][0])
        lobjs[n].edges[1] = Vector(*l["edges"][1])
        lobjs[n].nitems = int(l["nitems"])
        lobjs[n].converter = None
        lobjs[n].color = Color(*l.get("color", (255, 255, 255)))
        lobjs[n].width = float(l["width"])
    cam, light, center = Vector(*scene["camera"]), Vector(*scene["light"]), Vector(*scene["center"])
    look = Vector(*scene.get("look_at", (0.0, 0.0, 0.0)))
    local = (Vector * 2)(Vector(*scene["local_x"]), Vector(*scene["local_y"]))
    bg = Color(*scene.get("bg", (0, 0, 0)))
    sizes = (C.c_size_t * 2)(int(scene["width"]), int(scene["height"]))
    rc = L.contour3d_execute(info, C.byref(cam), C.byref(look), C.byref(light), C.byref(sizes), C.byref(center), C.byref(local),
                             C.byref(bg), len(contours), cobjs, len(lines), lobjs, fname.encode())
    return rc


def read_ppm(path: str) -> np.ndarray:
    """P6 -> (H, W, 3) uint8 in file order."""
    with open(path, "rb") as fp:
        data = fp.read()
    parts = data.split(b"\n", 3)
    assert parts[0] == b"P6" and parts[2] == b"255", parts[:3]
    w, h = (int(v) for v in parts[1].split())
    return np.frombuffer(parts[3], dtype=np.uint8, count=3 * w * h).reshape(h, w, 3)
