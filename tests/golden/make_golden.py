"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref).

    python tests/golden/make_golden.py

Run in the development container (needs oracle/_ref, i.e. /root/reference at
build time).  Each fixture stores its INPUT arrays next to the reference's
outputs, so consumers need neither the reference nor numpy's RNG stream.
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from contour3d_b200 import casefile, scenes  # noqa: E402
from oracle import reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def ref_image(scene, grids, fields, contours, lines, ranks=1, dims=None):
    with tempfile.TemporaryDirectory() as tmp:
        case, ppm = os.path.join(tmp, "case.bin"), os.path.join(tmp, "out.ppm")
        casefile.write_case(case, scene, grids, fields, contours, lines)
        reference.run_case(case, ppm, ranks=ranks, dims=dims)
        raw = open(ppm, "rb").read()
    from contour3d_b200.api import read_ppm  # pure python helper
    with tempfile.NamedTemporaryFile(suffix=".ppm") as fp:
        fp.write(raw)
        fp.flush()
        return read_ppm(fp.name).copy(), hashlib.md5(raw).hexdigest()


def main():
    reference.build()
    os.makedirs(GOLDEN, exist_ok=True)
    rng = np.random.default_rng(20261017)

    # 1. stage level: ragged 11 x 10 x 9 smooth field on a stretched grid, fp64
    nx, ny, nz = 11, 10, 9
    gx = np.sort(rng.uniform(-0.5, 0.5, nx))
    gy = -0.5 + (np.arange(ny) + 0.5) / ny
    gz = np.sort(rng.uniform(-0.4, 0.6, nz))
    x, y, z = gx[None, None, :], gy[None, :, None], gz[:, None, None]
    field = np.sin(5.1 * x + 0.3) * np.cos(4.3 * y) + 0.7 * np.sin(6.2 * z + 1.1) * np.cos(3.3 * x * y)
    thr = 0.137
    tris = reference.tessellate(field, [gx, gy, gz], thr)
    np.savez_compressed(os.path.join(GOLDEN, "tess_ragged.npz"), field=field, gx=gx, gy=gy, gz=gz, threshold=thr, triangles=tris)
    print("tess_ragged", len(tris))

    # 2. stage level: integer-valued field, threshold hit exactly (zero-area triangles -> NaN normals)
    n = 8
    fint = rng.integers(-2, 3, size=(n, n, n)).astype(np.float64)
    g = scenes.uniform_grids(n, n, n)
    tris = reference.tessellate(fint, g, 1.0)
    np.savez_compressed(os.path.join(GOLDEN, "tess_exact_hits.npz"), field=fint, gx=g[0], gy=g[1], gz=g[2], threshold=1.0, triangles=tris)
    print("tess_exact_hits", len(tris), "nan areas", int((tris["area"] == 0).sum()))

    # 3. whole frame: 24^3 fp64, two contours + box lines, 160 x 120, reference run on 1 and 4 ranks
    f24, g24 = scenes.sines(24, seed=3)
    scene = scenes.scene(160, 120, bg=(10, 20, 30))
    contours = [dict(field=0, threshold=-0.25, color=(0, 255, 255)), dict(field=0, threshold=0.25, color=(255, 255, 0))]
    lines = scenes.box_lines(nitems=24, width=2.0)
    img1, md5_1 = ref_image(scene, g24, [f24], contours, lines, ranks=1)
    img4, md5_4 = ref_image(scene, g24, [f24], contours, lines, ranks=4)
    assert md5_1 == md5_4, "reference is not decomposition invariant?"
    np.savez_compressed(os.path.join(GOLDEN, "frame_sines24.npz"), field=f24, gx=g24[0], gy=g24[1], gz=g24[2],
                        scene=json.dumps(scene), contours=json.dumps(contours), lines=json.dumps(lines), image=img1)
    print("frame_sines24", md5_1)

    # 4. whole frame: fp32 gyroid 32^3 (config 2 in miniature), 1 isovalue (double)0.1f, 16:9
    gy32 = scenes.gyroid(32, dtype=np.float32, periods=2.0)
    g32 = scenes.uniform_grids(32, 32, 32)
    scene2 = scenes.scene(192, 108)
    contours2 = [dict(field=0, threshold=float(np.float32(0.1)), color=(200, 120, 40))]
    img, md5 = ref_image(scene2, g32, [gy32], contours2, [], ranks=2)
    np.savez_compressed(os.path.join(GOLDEN, "frame_gyroid32.npz"), field=gy32, gx=g32[0], gy=g32[1], gz=g32[2],
                        scene=json.dumps(scene2), contours=json.dumps(contours2), lines=json.dumps([]), image=img)
    print("frame_gyroid32", md5)

    # 5. the reference's own demo (config 1): md5 of ./a.out's output.ppm at 1 and 2 ranks, and of the
    #    same scene rendered from the case file written by oracle/demo_case.c
    with tempfile.TemporaryDirectory() as tmp:
        md5s = {}
        for ranks in (1, 2):
            ppm = reference.run_demo(tmp, ranks=ranks)
            md5s[f"a.out ranks={ranks}"] = hashlib.md5(open(ppm, "rb").read()).hexdigest()
        case = os.path.join(tmp, "demo.bin")
        reference.write_demo_case(case)
        out = os.path.join(tmp, "case.ppm")
        reference.run_case(case, out, ranks=8)
        md5s["ref_driver(demo_case) ranks=8"] = hashlib.md5(open(out, "rb").read()).hexdigest()
        small = os.path.join(tmp, "demo32.bin")
        reference.write_demo_case(small, 32, 320, 240)
        reference.run_case(small, out, ranks=2)
        md5s["ref_driver(demo_case 32 320 240)"] = hashlib.md5(open(out, "rb").read()).hexdigest()
    assert len({md5s["a.out ranks=1"], md5s["a.out ranks=2"], md5s["ref_driver(demo_case) ranks=8"]}) == 1
    json.dump(dict(note="gcc 13.3 -std=c99 -O3, glibc rand()/sin(); survey md5 37d9181b1c4d4f6d8ebb2cf642897f32", md5=md5s),
              open(os.path.join(GOLDEN, "demo_md5.json"), "w"), indent=1)
    print(md5s)


if __name__ == "__main__":
    main()
