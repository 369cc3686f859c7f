"""Sweep-only probe on a GPU box: time of the classification sweep vs number of fused iso-values and element type.

    python tools/sweep_probe.py [size]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from contour3d_b200 import api, scenes

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = api.Context()
grids = scenes.uniform_grids(size, size, size)
scene = scenes.scene(640, 360)      # small image: the raster is not what is measured here
iso = (-0.3, 0.05, 0.4, 0.6)
for dtype, n in ((np.float32, size), (np.float64, min(size, 768))):
    f = ctx.field((n, n, n), dtype=dtype).synthesize(kind=1, seed=1234, octaves=5, gain=0.5)
    g = scenes.uniform_grids(n, n, n)
    for nthr in (1, 2, 3, 4):
        cs = [dict(field=f, grids=g, threshold=float(np.float32(t)), color=(255, 255, 255)) for t in iso[:nthr]]
        best = None
        for _ in range(4):
            ctx.render(scene, cs, [], want_rgb=False)
            t = ctx.timings()
            ms = t["sweep_each"][0]
            best = ms if best is None else min(best, ms)
        gb = t["sweep_bytes"][0] / 1e9
        print(f"{np.dtype(dtype).name} {n}^3 iso-values/launch={t['sweep_thresholds'][0]} sweeps={len(t['sweep_each'])} "
              f"best {best:.3f} ms  {gb / best * 1e3:.0f} GB/s  expand {t['expand']:.3f} ms active {t['active_cubes']}")
    f.destroy()
ctx.destroy()
