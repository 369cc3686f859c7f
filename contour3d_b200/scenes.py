"""Scenes and synthetic fields of the BASELINE.json configurations (host side,
numpy; the benchmark-size fields are generated on the device by
c3d_field_synthesize).

The camera is the reference demo's (src/main.c:127-163 of NaokiHori/Contour3D):
position (0.866, -1.5, 1.0), screen centre = position / 2, screen vectors of
length 4/3 and 1 (4:3).  Other aspect ratios scale the horizontal vector.
"""
from __future__ import annotations

import numpy as np

DEMO_CAMERA = (+8.660254037844386e-1, -1.500000000000000e0, +1.000000000000000e0)
DEMO_CENTER = (+4.330127018922194e-1, -7.500000000000003e-1, +5.000000000000003e-1)
DEMO_LOCAL_X = (+1.154700538379251e0, +6.666666666666665e-1, 0.0)
DEMO_LOCAL_Y = (-2.500000000000000e-1, +4.330127018922195e-1, +8.660254037844386e-1)
DEMO_LIGHT = (0.0, 2.0, -1.0)


def scene(width=1440, height=1080, bg=(0, 0, 0)) -> dict:
    """Demo camera with the horizontal screen vector stretched to width:height."""
    scale = (width / height) / (4.0 / 3.0)
    return dict(
        camera=DEMO_CAMERA, look_at=(0.0, 0.0, 0.0), light=DEMO_LIGHT, center=DEMO_CENTER,
        local_x=tuple(v * scale for v in DEMO_LOCAL_X), local_y=DEMO_LOCAL_Y,
        width=int(width), height=int(height), bg=tuple(bg),
    )


def uniform_grids(nx, ny, nz):
    """cell-centred uniform coordinates on [-0.5, 0.5] (src/main.c:45-53)"""
    return [-0.5 + (np.arange(n, dtype=np.float64) + 0.5) / n for n in (nx, ny, nz)]


def box_lines(nitems=128, width=4.0, color=(255, 255, 255)):
    """the demo's eight domain edges: four along y, four along z (src/main.c:223-306)"""
    out = []
    for x, z in ((-0.5, -0.5), (0.5, -0.5), (-0.5, 0.5), (0.5, 0.5)):
        out.append(dict(edges=((x, -0.5, z), (x, 0.5, z)), nitems=nitems, width=width, color=color))
    for x, y in ((-0.5, -0.5), (0.5, -0.5), (-0.5, 0.5), (0.5, 0.5)):
        out.append(dict(edges=((x, y, -0.5), (x, y, 0.5)), nitems=nitems, width=width, color=color))
    return out


def gyroid(n, dtype=np.float32, periods=4.0):
    """config 2: sin X cos Y + sin Y cos Z + sin Z cos X, X = 2 pi periods x; (nz, ny, nx)"""
    gx, gy, gz = uniform_grids(n, n, n)
    s = 2.0 * np.pi * periods
    X, Y, Z = s * gx[None, None, :], s * gy[None, :, None], s * gz[:, None, None]
    f = np.sin(X) * np.cos(Y) + np.sin(Y) * np.cos(Z) + np.sin(Z) * np.cos(X)
    return f.astype(dtype)


def _lowbias32(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def _lattice(ix, iy, iz, seed):
    with np.errstate(over="ignore"):
        h = _lowbias32(iz.astype(np.uint32) * np.uint32(0xC2B2AE3D) ^ np.uint32(seed))
        h = _lowbias32(iy.astype(np.uint32) * np.uint32(0x85EBCA77) ^ h)
        h = _lowbias32(ix.astype(np.uint32) * np.uint32(0x9E3779B1) ^ h)
    return (h >> np.uint32(8)).astype(np.float32) / np.float32(16777216.0)


def noise(n, seed=1234, octaves=5, gain=0.5, dtype=np.float32):
    """config 3 style multi-octave value noise in [-1, 1]; (nz, ny, nx).
    Same construction as the device generator (k_synth kind 1) but NOT bit-identical to it
    (float evaluation order differs); tests that need identical inputs upload this array."""
    g = uniform_grids(n, n, n)
    x = (g[0][None, None, :] + 0.5).astype(np.float32)
    y = (g[1][None, :, None] + 0.5).astype(np.float32)
    z = (g[2][:, None, None] + 0.5).astype(np.float32)
    total = np.zeros((n, n, n), dtype=np.float32)
    amp, norm, freq = 1.0, 0.0, 4.0
    for o in range(octaves):
        px, py, pz = x * freq, y * freq, z * freq
        fx, fy, fz = np.floor(px), np.floor(py), np.floor(pz)
        tx, ty, tz = px - fx, py - fy, pz - fz
        tx, ty, tz = tx * tx * (3 - 2 * tx), ty * ty * (3 - 2 * ty), tz * tz * (3 - 2 * tz)
        ix, iy, iz = fx.astype(np.int64), fy.astype(np.int64), fz.astype(np.int64)
        acc = np.zeros((n, n, n), dtype=np.float32)
        for c in range(8):
            dx, dy, dz = c & 1, (c >> 1) & 1, c >> 2
            w = (tx if dx else 1 - tx) * (ty if dy else 1 - ty) * (tz if dz else 1 - tz)
            acc = acc + w * _lattice(np.broadcast_to(ix + dx, (n, n, n)), np.broadcast_to(iy + dy, (n, n, n)),
                                     np.broadcast_to(iz + dz, (n, n, n)), seed + o)
        total += np.float32(amp) * (2 * acc - 1)
        norm += amp
        amp *= gain
        freq *= 2.0
    return (total / np.float32(norm)).astype(dtype)


def sines(n, seed=7, dtype=np.float64):
    """a smooth random field like the demo's (sum of 64 sine products, normalised to [-1, 1])
    on a Chebyshev x grid; returns (field, grids)"""
    rng = np.random.default_rng(seed)
    gx = -0.5 + 0.5 * (1.0 + np.cos(np.pi - np.pi * (np.arange(n) + 0.5) / n))
    gy = -0.5 + (np.arange(n) + 0.5) / n
    gz = gy.copy()
    f = np.zeros((n, n, n))
    for kz in range(1, 5):
        for ky in range(1, 5):
            for kx in range(1, 5):
                py, pz = rng.uniform(0, 2 * np.pi, 2)
                f += (np.sin(2 * np.pi * kx * gx)[None, None, :] * np.sin(2 * np.pi * ky * gy + py)[None, :, None]
                      * np.sin(2 * np.pi * kz * gz + pz)[:, None, None])
    f = (f - f.min()) / (f.max() - f.min()) * 2.0 - 1.0
    return f.astype(dtype), [gx, gy, gz]


def synth_host(gl, kind=1, seed=1234, octaves=5, gain=0.5, window=None, offset=(0, 0, 0), nthreads=None):
    """The device generator's values (c3d_field_synthesize kinds 1 and 2) computed on the HOST, bit for bit, by
    libc3d_synth_host.so (plain C, no CUDA): the window `window` = (nx, ny, nz) at global `offset` of the gl = (nx, ny, nz)
    field, returned as a float32 array (nz, ny, nx).  Used where the CUDA library must not be involved (the CPU
    reference arm of bench.py)."""
    import ctypes as C
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libc3d_synth_host.so")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -m contour3d_b200.build`")
    lib = C.CDLL(path)
    lib.c3d_synth_host.restype = C.c_int
    lib.c3d_synth_host.argtypes = [C.c_void_p, C.POINTER(C.c_longlong * 3), C.POINTER(C.c_longlong * 3), C.POINTER(C.c_longlong * 3),
                                   C.c_int, C.c_uint32, C.c_int, C.c_double, C.c_int]
    window = tuple(int(v) for v in (window if window is not None else gl))
    out = np.empty((window[2], window[1], window[0]), dtype=np.float32)
    if nthreads is None:
        try:
            nthreads = len(os.sched_getaffinity(0))
        except AttributeError:
            nthreads = os.cpu_count() or 1
    rc = lib.c3d_synth_host(out.ctypes.data, C.byref((C.c_longlong * 3)(*window)), C.byref((C.c_longlong * 3)(*[int(v) for v in offset])),
                            C.byref((C.c_longlong * 3)(*[int(v) for v in gl])), int(kind), int(seed), int(octaves), float(gain), int(nthreads))
    if rc != 0:
        raise ValueError("c3d_synth_host: bad argument")
    return out
