"""Do the raster's visibility shortcuts change anything?  Renders large frames (the bench field, several sizes and
cameras) twice -- every shortcut on, and the named ones off through their environment switches -- and compares the
canvases (depth + winner key of every pixel) and the images bit for bit.  The oracle cannot referee frames of this
size in reasonable time; the shortcut-free pipeline is itself checked against it at small sizes by the parity tests.

    python tools/gpu_shortcut_check.py [CONTOUR3D_NO_CULL] ...
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from contour3d_b200 import api, scenes

NOISE = dict(kind=1, seed=1234, octaves=5, gain=0.5)
ISO = (-0.3, 0.05, 0.4)
COL = ((255, 96, 64), (96, 255, 96), (64, 128, 255))


def cameras(width, height):
    yield "demo", scenes.scene(width, height)
    for el, az, roll, dc in ((75.0, 200.0, 0.0, 2.2), (20.0, 40.0, 15.0, 1.6), (89.0, 0.0, 0.0, 3.0), (50.0, 135.0, -30.0, 1.2)):
        s = scenes.scene(width, height)
        s.update(api.camera_from_angles(el, az, roll, camera_distance=dc, screen_distance=dc / 2, screen_width=width / height, screen_height=1.0))
        yield f"el{el:g}_az{az:g}_roll{roll:g}_d{dc:g}", s


def frames(switches, size, width, height, kind=1):
    for k in switches:
        os.environ.pop(k, None)
    out = {}
    for label, env in (("on", {}), ("off", {k: "1" for k in switches})):
        os.environ.update(env)
        ctx = api.Context(device=0)
        field = ctx.field((size, size, size), dtype=np.float32).synthesize(**(NOISE if kind == 1 else dict(kind=0, seed=0, octaves=1, gain=0.5)))
        grids = scenes.uniform_grids(size, size, size)
        cs = [dict(field=field, grids=grids, threshold=float(np.float32(t)), color=c) for t, c in zip(ISO, COL)]
        for name, scene in cameras(width, height):
            img = ctx.render(scene, cs, scenes.box_lines(nitems=64, width=3.0))
            depth, key = ctx.canvas(width, height)
            t = ctx.timings()
            out.setdefault(name, {})[label] = (img, depth, key, t["raster"], t["culled_triangles"])
        field.destroy()
        ctx.destroy()
        for k in env:
            os.environ.pop(k, None)
    return out


def main():
    switches = sys.argv[1:] or ["CONTOUR3D_NO_CULL"]
    bad = 0
    for size, (w, h), kind in ((1024, (3840, 2160), 1), (512, (1920, 1080), 1), (384, (2560, 1440), 0), (256, (640, 360), 1)):
        res = frames(switches, size, w, h, kind)
        for name, r in res.items():
            (ia, da, ka, ta, ca), (ib, db, kb, tb, cb) = r["on"], r["off"]
            same = np.array_equal(ia, ib) and np.array_equal(da, db, equal_nan=True) and np.array_equal(ka, kb)
            covered = float((ka != 0).mean())
            print(f"{size}^3 {w}x{h} kind {kind} {name:28s} identical={same} covered={covered:.2f} raster on/off {ta:.3f}/{tb:.3f} ms, triangles dropped {ca}/{cb}")
            if not same:
                bad += 1
                print("   pixels differing:", int((ka != kb).sum()), "depth:", int((da != db).sum()))
    print("FAILED" if bad else "ok: the shortcuts changed no pixel, depth or winner")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
