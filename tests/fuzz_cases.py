"""Seeded random frames for the oracle-vs-reference fuzz (tests/test_oracle_vs_reference.py).

Everything the path accepts is drawn at random: ragged shapes, fp32 / fp64 arrays, non-uniform coordinates, one to three
iso-values (a third of them EXACTLY a value of the array: vertices on grid points, degenerate triangles, NaN normals),
cameras anywhere between inside the volume and far away, screens from 5 pixels to a few hundred with any aspect ratio,
non-square pixels, up to three lines of random width, and the reference's decomposition (1-3 ranks, y or z split).
"""
import numpy as np


def random_case(seed):
    rng = np.random.default_rng(seed)
    nx, ny, nz = (int(v) for v in rng.integers(6, 19, 3))
    dt = np.float32 if rng.random() < 0.5 else np.float64
    grids = []
    for n in (nx, ny, nz):
        x = np.cumsum(rng.uniform(0.3, 1.7, n))
        x = (x - x[0]) / (x[-1] - x[0]) - 0.5
        grids.append(x * rng.uniform(0.6, 1.0))
    kind = rng.integers(0, 3)
    Z, Y, X = np.meshgrid(grids[2], grids[1], grids[0], indexing="ij")
    if kind == 0:
        f = rng.normal(size=(nz, ny, nx)) * 0.3
    elif kind == 1:
        f = np.sin(5 * X + rng.uniform(0, 6)) * np.cos(4 * Y) + np.sin(6 * Z + 1) * rng.uniform(0.2, 1)
    else:
        f = np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.3 + 0.05 * rng.normal(size=X.shape)
    f = f.astype(dt)
    contours = []
    for _ in range(int(rng.integers(1, 4))):
        if rng.random() < 0.3:
            thr = float(f.flat[rng.integers(0, f.size)])
        else:
            thr = float(dt(rng.uniform(np.quantile(f, 0.2), np.quantile(f, 0.8))))
        contours.append(dict(field=0, threshold=thr, color=tuple(int(v) for v in rng.integers(0, 256, 3))))
    d = rng.normal(size=3)
    d /= np.linalg.norm(d)
    cam = d * rng.uniform(0.25, 2.5)
    look = rng.uniform(-0.15, 0.15, 3)
    fwd = look - cam
    dist = np.linalg.norm(fwd)
    fwd /= dist
    up = np.array([0, 0, 1.0]) if abs(fwd[2]) < 0.9 else np.array([0, 1.0, 0])
    lx = np.cross(fwd, up)
    lx /= np.linalg.norm(lx)
    ly = np.cross(lx, fwd)
    center = cam + fwd * dist * rng.uniform(0.3, 0.7)
    W, H = int(rng.integers(5, 260)), int(rng.integers(5, 200))
    sx = rng.uniform(0.5, 2.0)
    sy = sx * H / W * rng.uniform(0.8, 1.25)
    scene = dict(camera=tuple(cam), look_at=tuple(look), light=tuple(rng.normal(size=3)), center=tuple(center),
                 local_x=tuple(lx * sx), local_y=tuple(ly * sy), width=W, height=H, bg=tuple(int(v) for v in rng.integers(0, 256, 3)))
    lines = []
    for _ in range(int(rng.integers(0, 4))):
        lines.append(dict(edges=(tuple(rng.uniform(-0.6, 0.6, 3)), tuple(rng.uniform(-0.6, 0.6, 3))), nitems=int(rng.integers(2, 40)),
                          width=float(rng.uniform(0.5, 5.0)), color=tuple(int(v) for v in rng.integers(0, 256, 3))))
    ranks = int(rng.integers(1, 4))
    dims = None
    if ranks > 1:
        dims = (1, ranks, 1) if rng.random() < 0.5 else (1, 1, ranks)
    return scene, grids, f, contours, lines, ranks, dims
