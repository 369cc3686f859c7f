"""ctypes binding of the CPU oracle (oracle/c3d_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(contour3d_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libc3d_oracle.so")

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
assert TRIANGLE_DTYPE.itemsize == 208


class Vec(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]


CONVERTER = C.CFUNCTYPE(Vec, Vec)


class Contour(C.Structure):
    _fields_ = [
        ("dims", C.c_size_t * 3),
        ("grid", C.POINTER(C.c_double) * 3),
        ("field", C.POINTER(C.c_double)),
        ("threshold", C.c_double),
        ("color", C.c_uint8 * 3),
        ("converter", C.c_void_p),
    ]


class Line(C.Structure):
    _fields_ = [
        ("edges", (C.c_double * 3) * 2),
        ("nitems", C.c_size_t),
        ("width", C.c_double),
        ("color", C.c_uint8 * 3),
        ("converter", C.c_void_p),
    ]


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


class Stats(C.Structure):
    _fields_ = [
        ("cubes", C.c_uint64),
        ("active_cubes", C.c_uint64),
        ("triangles", C.c_uint64),
        ("rendered", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement (and, when /root/reference is present, the
    reference itself into oracle/_ref) with oracle/Makefile."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
        os.path.join(HERE, "c3d_oracle.c")
    ):
        subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.oracle_case_codes.restype = C.c_size_t
        L.oracle_case_codes.argtypes = [C.POINTER(Contour), C.c_void_p, C.c_void_p]
        L.oracle_tessellate.restype = C.c_size_t
        L.oracle_tessellate.argtypes = [C.POINTER(Contour), C.c_size_t, C.c_void_p, C.POINTER(Stats)]
        L.oracle_canvas_clear.argtypes = [C.POINTER(Scene), C.c_void_p, C.c_void_p]
        L.oracle_render_contour.argtypes = [C.POINTER(Scene), C.POINTER(Contour), C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        L.oracle_render_line.argtypes = [C.POINTER(Scene), C.POINTER(Line), C.c_void_p, C.c_void_p]
        L.oracle_canvas_merge.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_write_ppm.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_void_p]
        L.oracle_frame.argtypes = [
            C.POINTER(Scene), C.c_size_t, C.POINTER(Contour), C.c_size_t, C.POINTER(Line), C.c_int,
            C.c_void_p, C.c_void_p, C.POINTER(Stats),
        ]
        L.oracle_converter_of.restype = C.c_void_p
        L.oracle_converter_of.argtypes = [C.c_int]
        L.oracle_convert_points.restype = None
        L.oracle_convert_points.argtypes = [C.c_int, C.c_size_t, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def converter_ptr(kind: int):
    """C function pointer of a built-in converter (0 identity -> None, 1 cylindrical, 2 stretched, 3 twisted: the case files' numbering)"""
    return lib().oracle_converter_of(int(kind)) if kind else None


def convert_points(kind: int, points) -> np.ndarray:
    """the built-in converter applied to (n, 3) points, in C arithmetic (bit-identical to what the oracle / the case driver compute)"""
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    out = np.empty_like(pts)
    lib().oracle_convert_points(int(kind), len(pts), pts.ctypes.data, out.ctypes.data)
    return out


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class _Keep:
    """ctypes struct + the numpy arrays it points into."""

    def __init__(self, struct, refs):
        self.struct = struct
        self.refs = refs


def make_contour(field: np.ndarray, grids, threshold: float, color=(255, 255, 255), converter: int = 0) -> _Keep:
    """field: (nz, ny, nx) array (x fastest), widened to float64 exactly."""
    f = np.ascontiguousarray(field, dtype=np.float64)
    nz, ny, nx = f.shape
    g = [np.ascontiguousarray(x, dtype=np.float64) for x in grids]
    assert (len(g[0]), len(g[1]), len(g[2])) == (nx, ny, nz)
    c = Contour()
    c.dims[:] = [nx, ny, nz]
    for d in range(3):
        c.grid[d] = _dptr(g[d])
    c.field = _dptr(f)
    c.threshold = float(threshold)
    c.color[:] = list(color)
    c.converter = converter_ptr(converter)
    return _Keep(c, (f, g))


def make_scene(scene: dict) -> Scene:
    s = Scene()
    s.camera[:] = scene["camera"]
    s.light[:] = scene["light"]
    s.center[:] = scene["center"]
    s.local_x[:] = scene["local_x"]
    s.local_y[:] = scene["local_y"]
    s.width = int(scene["width"])
    s.height = int(scene["height"])
    s.bg[:] = list(scene.get("bg", (0, 0, 0)))
    return s


def make_line(line: dict) -> Line:
    l = Line()
    l.edges[0][:] = line["edges"][0]
    l.edges[1][:] = line["edges"][1]
    l.nitems = int(line["nitems"])
    l.width = float(line["width"])
    l.color[:] = list(line.get("color", (255, 255, 255)))
    l.converter = converter_ptr(line.get("converter", 0))
    return l


def case_codes(field, grids, threshold):
    """(codes uint32 [cz,cy,cx] with six 4-bit tetra masks, counts uint8)."""
    k = make_contour(field, grids, threshold)
    nz, ny, nx = k.refs[0].shape
    shape = (nz - 1, ny - 1, nx - 1)
    codes = np.zeros(shape, dtype=np.uint32)
    counts = np.zeros(shape, dtype=np.uint8)
    n = lib().oracle_case_codes(C.byref(k.struct), codes.ctypes.data, counts.ctypes.data)
    assert n == codes.size
    return codes, counts


def tessellate(field, grids, threshold):
    """All triangles in (k, j, i, tri) order as a structured array + stats."""
    k = make_contour(field, grids, threshold)
    stats = Stats()
    total = lib().oracle_tessellate(C.byref(k.struct), 0, None, C.byref(stats))
    out = np.zeros(total, dtype=TRIANGLE_DTYPE)
    if total:
        got = lib().oracle_tessellate(C.byref(k.struct), total, out.ctypes.data, C.byref(stats))
        assert got == total
    return out, stats.as_dict()


def frame(scene: dict, contours, lines=(), nthreads: int = 1):
    """Render a whole frame.

    contours: list of dicts {field, grids, threshold, color[, converter: built-in kind 0..3]}
    lines:    list of dicts {edges, nitems, width, color[, converter]}
    returns (rgb uint8 [H, W, 3] in canvas order (not flipped), depth float64 [H, W], stats)
    """
    s = make_scene(scene)
    keeps = [make_contour(c["field"], c["grids"], c["threshold"], c.get("color", (255, 255, 255)), c.get("converter", 0)) for c in contours]
    carr = (Contour * max(1, len(keeps)))()
    for n, k in enumerate(keeps):
        carr[n] = k.struct
    larr = (Line * max(1, len(lines)))()
    for n, l in enumerate(lines):
        larr[n] = make_line(l)
    H, W = int(scene["height"]), int(scene["width"])
    depth = np.zeros((H, W), dtype=np.float64)
    rgb = np.zeros((H, W, 3), dtype=np.uint8)
    stats = Stats()
    rc = lib().oracle_frame(
        C.byref(s), len(keeps), carr, len(lines), larr, int(nthreads), depth.ctypes.data, rgb.ctypes.data, C.byref(stats)
    )
    if rc != 0:
        raise RuntimeError("oracle_frame failed")
    return rgb, depth, stats.as_dict()


def write_ppm(fname: str, rgb: np.ndarray):
    H, W, _ = rgb.shape
    rgb = np.ascontiguousarray(rgb)
    rc = lib().oracle_write_ppm(fname.encode(), W, H, rgb.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle_write_ppm failed")
