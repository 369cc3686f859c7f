"""Case files: one self-contained binary description of a frame (scene, grids,
contour / line objects, global fields).  The same file is consumed by

* the unmodified reference through oracle/ref_driver.c (CPU baseline),
* the product's C demo program (contour3d_b200/csrc/c3d_case_main.c),
* the Python tests.

Layout (little endian, no implicit padding):
  header   char magic[8]="C3DCASE1"; u64 n[3], width, height, ncontours, nlines,
           elem_bytes (4|8), nfields; f64 camera[3], look_at[3], light[3],
           center[3], local_x[3], local_y[3]; u8 bg[3], converter (0 identity, 1 cylindrical,
           2 stretched, 3 twisted: functions built into c3d_case_main.c), pad[4]
  grids    f64 x[n0], y[n1], z[n2]
  contours ncontours x { u64 field; f64 threshold; u8 color[3], pad[5] }
  lines    nlines x { f64 edges[6]; u64 nitems; f64 width; u8 color[3], pad[5] }
  fields   nfields x n2*n1*n0 elements (f32 or f64), x fastest
"""
from __future__ import annotations

import numpy as np

MAGIC = b"C3DCASE1"

HEADER = np.dtype(
    [
        ("magic", "S8"),
        ("n", "<u8", (3,)),
        ("width", "<u8"),
        ("height", "<u8"),
        ("ncontours", "<u8"),
        ("nlines", "<u8"),
        ("elem_bytes", "<u8"),
        ("nfields", "<u8"),
        ("camera", "<f8", (3,)),
        ("look_at", "<f8", (3,)),
        ("light", "<f8", (3,)),
        ("center", "<f8", (3,)),
        ("local_x", "<f8", (3,)),
        ("local_y", "<f8", (3,)),
        ("bg", "u1", (3,)),
        ("pad", "u1", (5,)),
    ]
)
CONTOUR = np.dtype([("field", "<u8"), ("threshold", "<f8"), ("color", "u1", (3,)), ("pad", "u1", (5,))])
LINE = np.dtype([("edges", "<f8", (6,)), ("nitems", "<u8"), ("width", "<f8"), ("color", "u1", (3,)), ("pad", "u1", (5,))])


def write_case(path, scene: dict, grids, fields, contours, lines=()):
    """scene: dict(camera, look_at?, light, center, local_x, local_y, width, height, bg?)
    fields: list of (nz, ny, nx) arrays, all float32 or all float64
    contours: list of dict(field=index, threshold, color)
    lines: list of dict(edges=((x,y,z),(x,y,z)), nitems, width, color)
    """
    fields = [np.ascontiguousarray(f) for f in fields]
    dt = fields[0].dtype
    assert dt in (np.float32, np.float64) and all(f.dtype == dt for f in fields)
    nz, ny, nx = fields[0].shape
    h = np.zeros((), dtype=HEADER)
    h["magic"] = MAGIC
    h["n"] = (nx, ny, nz)
    h["width"], h["height"] = scene["width"], scene["height"]
    h["ncontours"], h["nlines"] = len(contours), len(lines)
    h["elem_bytes"] = dt.itemsize
    h["nfields"] = len(fields)
    for key in ("camera", "light", "center", "local_x", "local_y"):
        h[key] = scene[key]
    h["look_at"] = scene.get("look_at", (0.0, 0.0, 0.0))
    h["bg"] = scene.get("bg", (0, 0, 0))
    h["pad"][0] = int(scene.get("converter", 0))    # 0 identity, 1 cylindrical, 2 stretched, 3 twisted (c3d_case_main.c)
    cs = np.zeros(len(contours), dtype=CONTOUR)
    for n, c in enumerate(contours):
        cs[n]["field"] = c.get("field", 0)
        cs[n]["threshold"] = c["threshold"]
        cs[n]["color"] = c.get("color", (255, 255, 255))
    ls = np.zeros(len(lines), dtype=LINE)
    for n, l in enumerate(lines):
        ls[n]["edges"] = np.asarray(l["edges"], dtype=np.float64).reshape(6)
        ls[n]["nitems"] = l["nitems"]
        ls[n]["width"] = l["width"]
        ls[n]["color"] = l.get("color", (255, 255, 255))
    with open(path, "wb") as fp:
        fp.write(h.tobytes())
        for d in range(3):
            g = np.ascontiguousarray(grids[d], dtype=np.float64)
            assert len(g) == (nx, ny, nz)[d]
            fp.write(g.tobytes())
        fp.write(cs.tobytes())
        fp.write(ls.tobytes())
        for f in fields:
            f.tofile(fp)


def read_case(path):
    """Inverse of write_case -> (scene, grids, fields, contours, lines)."""
    with open(path, "rb") as fp:
        h = np.frombuffer(fp.read(HEADER.itemsize), dtype=HEADER)[0]
        if bytes(h["magic"]) != MAGIC:
            raise ValueError(f"{path}: not a case file")
        nx, ny, nz = (int(v) for v in h["n"])
        grids = [np.frombuffer(fp.read(8 * n), dtype="<f8").copy() for n in (nx, ny, nz)]
        cs = np.frombuffer(fp.read(CONTOUR.itemsize * int(h["ncontours"])), dtype=CONTOUR)
        ls = np.frombuffer(fp.read(LINE.itemsize * int(h["nlines"])), dtype=LINE)
        dt = np.float32 if int(h["elem_bytes"]) == 4 else np.float64
        fields = [np.fromfile(fp, dtype=dt, count=nx * ny * nz).reshape(nz, ny, nx) for _ in range(int(h["nfields"]))]
    scene = {k: tuple(float(v) for v in h[k]) for k in ("camera", "look_at", "light", "center", "local_x", "local_y")}
    scene.update(width=int(h["width"]), height=int(h["height"]), bg=tuple(int(v) for v in h["bg"]), converter=int(h["pad"][0]))
    contours = [dict(field=int(c["field"]), threshold=float(c["threshold"]), color=tuple(int(v) for v in c["color"])) for c in cs]
    lines = [
        dict(edges=(tuple(l["edges"][:3]), tuple(l["edges"][3:])), nitems=int(l["nitems"]), width=float(l["width"]),
             color=tuple(int(v) for v in l["color"]))
        for l in ls
    ]
    return scene, grids, fields, contours, lines
