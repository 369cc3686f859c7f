"""GPU: the CUDA path, called through the C-ABI, against the oracle and the golden vectors.

Bars (BASELINE.json): triangle counts and per-cube case codes bit-exact; vertex
positions / normals within 1e-5 relative (we assert bit-equality, which is
stronger); image differing in <= 0.1 % of pixels by <= 1 LSB (we assert 0)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_records_equal, image_diff, load_golden
from contour3d_b200 import api, casefile, scenes
from oracle import oracle

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5          # positions and normals, relative (BASELINE.json north_star)
MAX_PIXEL_FRACTION = 1e-3
MAX_LSB = 1


def _grids_for(shape, rng=None):
    nz, ny, nx = shape
    if rng is None:
        return scenes.uniform_grids(nx, ny, nz)
    return [np.sort(rng.uniform(-0.5, 0.5, n)) for n in (nx, ny, nz)]


def _check_stage(ctx, field, grids, thr):
    nz, ny, nx = field.shape
    f = ctx.field((nx, ny, nz), dtype=field.dtype).upload(field)
    try:
        codes, counts = ctx.case_codes(f, thr)
        ocodes, ocounts = oracle.case_codes(field, grids, thr)
        assert np.array_equal(codes, ocodes), "per-cube case codes differ"
        assert np.array_equal(counts, ocounts), "per-cube triangle counts differ"
        # streaming sweep: all active cubes, then the rendered subset
        for rendered_only in (False, True):
            entries = np.sort(ctx.sweep(f, thr, rendered_only=rendered_only))
            cz, cy, cx = counts.shape
            act = counts > 0
            if rendered_only:
                m = np.zeros_like(act)
                m[1:cz - 1, 1:cy - 1, 1:cx - 1] = True
                act &= m
            want_cubes = np.flatnonzero(act.reshape(-1)).astype(np.uint64)
            assert np.array_equal(entries >> np.uint64(8), want_cubes), "active-cube list differs"
        got = ctx.tessellate(f, grids, thr)
        want, _ = oracle.tessellate(field, grids, thr)
        assert_records_equal(got, want, "tessellate")
        # the tolerance the north star states, written out (bit-equality above implies it)
        if len(got):
            scale = np.abs(want["v"]).max()
            assert np.abs(got["v"] - want["v"]).max() <= REL_TOL * scale
    finally:
        f.destroy()


@pytest.mark.parametrize("shape,thr,dtype,seed", [
    ((16, 16, 16), 0.05, np.float64, 0),
    ((9, 10, 11), -0.2, np.float64, 1),       # ragged, scalar-load path (nx % 4 != 0)
    ((7, 6, 132), 0.1, np.float32, 2),        # two x segments, aligned
    ((5, 9, 130), 0.0, np.float32, 3),        # two x segments, unaligned rows
    ((4, 4, 4), 0.1, np.float32, 4),          # smallest array with a rendered cube
    ((2, 2, 2), 0.0, np.float64, 5),          # single cube, nothing rendered
    ((40, 33, 20), 0.3, np.float32, 6),       # several y tiles and z chunks
])
def test_stage_parity_random_fields(ctx, shape, thr, dtype, seed):
    rng = np.random.default_rng(seed)
    field = (rng.normal(size=shape) * 0.4).astype(dtype)
    _check_stage(ctx, field, _grids_for(shape, rng), float(dtype(thr)))


def test_stage_parity_smooth_field(ctx):
    f, g = scenes.sines(36, seed=5)
    _check_stage(ctx, f, g, 0.1)


def test_stage_parity_empty_surface(ctx):
    field = np.zeros((8, 8, 8), dtype=np.float32)
    _check_stage(ctx, field, _grids_for(field.shape), 0.5)     # all below
    _check_stage(ctx, field, _grids_for(field.shape), -0.5)    # all above


@pytest.mark.parametrize("name", ["tess_ragged.npz", "tess_exact_hits.npz"])
def test_stage_parity_golden_from_reference(ctx, name):
    g = load_golden(name)
    grids = [g["gx"], g["gy"], g["gz"]]
    nz, ny, nx = g["field"].shape
    f = ctx.field((nx, ny, nz), dtype=np.float64).upload(g["field"])
    try:
        assert_records_equal(ctx.tessellate(f, grids, g["threshold"]), g["triangles"], name)
    finally:
        f.destroy()


def _render_both(ctx, scene, fields, grids, contours, lines, dtype=None):
    nz, ny, nx = fields[0].shape
    dev = [ctx.field((nx, ny, nz), dtype=f.dtype).upload(f) for f in fields]
    try:
        cs = [dict(field=dev[c.get("field", 0)], grids=grids, threshold=c["threshold"], color=c["color"]) for c in contours]
        img = ctx.render(scene, cs, lines)
        depth, key = ctx.canvas(scene["width"], scene["height"])
    finally:
        for d in dev:
            d.destroy()
    ocs = [dict(field=fields[c.get("field", 0)], grids=grids, threshold=c["threshold"], color=c["color"]) for c in contours]
    rgb, odepth, stats = oracle.frame(scene, ocs, lines, nthreads=4)
    return img, rgb[::-1], depth, odepth, key


def _assert_image(img, want):
    frac, worst = image_diff(img, want)
    assert frac <= MAX_PIXEL_FRACTION and worst <= MAX_LSB, (frac, worst)
    assert frac == 0.0, f"{frac} of the pixels differ (bit-exactness expected)"


@pytest.mark.parametrize("name", ["frame_sines24.npz", "frame_gyroid32.npz"])
def test_frame_golden_from_reference(ctx, name):
    g = load_golden(name)
    grids = [g["gx"], g["gy"], g["gz"]]
    img, want, depth, odepth, key = _render_both(ctx, g["scene"], [g["field"]], grids, g["contours"], g["lines"])
    _assert_image(img, g["image"])
    _assert_image(img, want)
    assert np.array_equal(depth, odepth), "z-buffer differs from the oracle"


def test_frame_three_isovalues_noise_fp32(ctx):
    """config 3 in miniature: multi-octave noise, 3 isovalues, 16:9"""
    f = scenes.noise(48, seed=1234)
    g = scenes.uniform_grids(48, 48, 48)
    scene = scenes.scene(384, 216)
    contours = [dict(threshold=float(np.float32(t)), color=c) for t, c in
                ((-0.3, (255, 80, 80)), (0.05, (80, 255, 80)), (0.4, (80, 80, 255)))]
    img, want, depth, odepth, key = _render_both(ctx, scene, [f], g, contours, scenes.box_lines(nitems=32, width=3.0))
    _assert_image(img, want)
    assert np.array_equal(depth, odepth)
    assert (key != 0).sum() > 0.05 * key.size


def test_frame_two_fields(ctx):
    """config 4 in miniature: two different fields sharing camera and light"""
    a = scenes.noise(32, seed=1)
    b = scenes.gyroid(32, periods=1.5)
    g = scenes.uniform_grids(32, 32, 32)
    scene = scenes.scene(256, 144)
    contours = [dict(field=0, threshold=0.1, color=(250, 250, 50)), dict(field=1, threshold=float(np.float32(0.2)), color=(50, 250, 250)),
                dict(field=0, threshold=-0.1, color=(250, 50, 250))]
    img, want, depth, odepth, _ = _render_both(ctx, scene, [a, b], g, contours, [])
    _assert_image(img, want)
    assert np.array_equal(depth, odepth)


def test_frame_lines_only_and_background(ctx):
    scene = scenes.scene(160, 120, bg=(7, 8, 9))
    g = scenes.uniform_grids(8, 8, 8)
    f = np.zeros((8, 8, 8), dtype=np.float32)
    img, want, *_ = _render_both(ctx, scene, [f], g, [dict(threshold=0.5, color=(1, 2, 3))], scenes.box_lines(nitems=9, width=5.0))
    _assert_image(img, want)
    assert (img == np.array([7, 8, 9], dtype=np.uint8)).all(axis=2).any()


def test_frame_without_contour_objects(ctx):
    """num_contours = 0 is legal (src/contour3d/main.c:105-117 simply loops zero times): lines only, then nothing at all"""
    scene = scenes.scene(96, 64, bg=(10, 20, 30))
    lines = scenes.box_lines(nitems=7, width=3.0, color=(200, 100, 50))
    img = ctx.render(scene, [], lines)
    rgb, _, _ = oracle.frame(scene, [], lines, nthreads=1)
    _assert_image(img, rgb[::-1])
    assert (img == np.array([200, 100, 50], dtype=np.uint8)).all(axis=2).any()
    img = ctx.render(scene, [], [])
    assert (img == np.array([10, 20, 30], dtype=np.uint8)).all()


def test_frame_partly_off_screen_and_behind_camera(ctx):
    """triangles whose vertices fail projection are dropped silently (render.c:62-72), bbox clamps to the screen"""
    f, g = scenes.sines(24, seed=9)
    scene = scenes.scene(200, 150)
    scene["camera"] = (0.2, -0.45, 0.3)            # inside the volume: part of the surface is behind the screen
    scene["center"] = (0.1, -0.225, 0.15)
    img, want, depth, odepth, _ = _render_both(ctx, scene, [f], g, [dict(threshold=0.0, color=(255, 255, 255))], scenes.box_lines(nitems=20, width=2.0))
    _assert_image(img, want)
    assert np.array_equal(depth, odepth)


def test_frame_exact_threshold_hits_nan_normals(ctx):
    """zero-area triangles give NaN normals that fall back to the 0.15 ambient floor (render.c:151-155)"""
    g = load_golden("tess_exact_hits.npz")
    grids = [g["gx"], g["gy"], g["gz"]]
    scene = scenes.scene(200, 150)
    img, want, depth, odepth, _ = _render_both(ctx, scene, [g["field"]], grids, [dict(threshold=1.0, color=(200, 200, 200))], [])
    _assert_image(img, want)
    assert np.array_equal(depth, odepth)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_frame_collapsed_triangles_behind_occluders(ctx, dtype):
    """Grid values exactly on the iso-value collapse triangles to points, which the reference draws as +-1 pixel
    blobs of infinite depth (clamped area, render.c:103-106 -> weights +-0 -> 1 / +-0): they show through nearer
    surfaces.  The raster's visibility shortcuts (cube / triangle occlusion tests, front-to-back order) must not
    lose them: such cubes are flagged by k_expand and always take the exact path."""
    n = 40
    f = scenes.noise(n, seed=21).astype(dtype)
    thr_in, thr_out = dtype(0.05), dtype(-0.25)
    rng = np.random.default_rng(5)
    for _ in range(120):                                  # exact hits deep inside, below the outer surface
        k, j, i = rng.integers(6, n - 6, size=3)
        f[k, j, i] = thr_in
    g = scenes.uniform_grids(n, n, n)
    scene = scenes.scene(480, 270)
    contours = [dict(threshold=float(thr_out), color=(250, 160, 60)), dict(threshold=float(thr_in), color=(60, 160, 250))]
    img, want, depth, odepth, _ = _render_both(ctx, scene, [f], g, contours, scenes.box_lines(nitems=10, width=2.0))
    _assert_image(img, want)
    assert np.array_equal(depth, odepth)
    assert np.isinf(odepth).any(), "the scene was meant to contain collapsed triangles (infinite depth)"


@pytest.mark.parametrize("cuts", ["4,12,32", "1,2,3,5,8,13,21,34,55", ""])
def test_frame_hiz_refresh_does_not_change_the_image(cuts, monkeypatch):
    """The raster's list drawn in parts with a coarse z-buffer rebuilt in between (the cross-rank occlusion path of
    N > 1, forced on here on one GPU): many nested surfaces so that most cubes are hidden; image and z-buffer stay exact."""
    monkeypatch.setenv("CONTOUR3D_HIZ_CUTS", cuts)
    own = api.Context(device=0)
    try:
        n = 56
        f = scenes.noise(n, seed=77)
        g = scenes.uniform_grids(n, n, n)
        scene = scenes.scene(640, 360)
        contours = [dict(threshold=float(np.float32(t)), color=c) for t, c in
                    ((-0.35, (255, 80, 80)), (-0.1, (200, 200, 80)), (0.1, (80, 255, 80)), (0.3, (80, 80, 255)))]
        img, want, depth, odepth, _ = _render_both(own, scene, [f], g, contours, scenes.box_lines(nitems=24, width=2.0))
        _assert_image(img, want)
        assert np.array_equal(depth, odepth)
        t = own.timings()
        if cuts:
            assert t["launches"] >= 2 * len(cuts.split(",")), t["launches"]
            assert t["culled_cubes"] > 0
    finally:
        own.destroy()


def test_frame_shading_without_room_for_normals(monkeypatch):
    """Vertices that find no room in the normal array are smoothed per pixel (vertex_normal, the serial path of
    k_shade_blend): forced here by a tiny array; the image must not change."""
    monkeypatch.setenv("CONTOUR3D_SHADE_NORMALS_CAP", "100")
    own = api.Context(device=0)
    try:
        f = scenes.noise(40, seed=4)
        g = scenes.uniform_grids(40, 40, 40)
        scene = scenes.scene(320, 180)
        contours = [dict(threshold=float(np.float32(t)), color=c) for t, c in ((-0.2, (255, 120, 80)), (0.15, (80, 120, 255)))]
        img, want, depth, odepth, _ = _render_both(own, scene, [f], g, contours, [])
        _assert_image(img, want)
        assert np.array_equal(depth, odepth)
    finally:
        own.destroy()


def test_frame_work_buffers_grow_and_the_frame_is_redone(monkeypatch):
    """The number of cut cubes is unknown before the sweep: record and list buffers that overflow are grown
    to the counted size and the frame runs again.  Forced here by a tiny first allocation."""
    monkeypatch.setenv("CONTOUR3D_LIST_CAP", "300")
    own = api.Context(device=0)
    try:
        f = scenes.noise(40, seed=8)
        g = scenes.uniform_grids(40, 40, 40)
        scene = scenes.scene(320, 180)
        contours = [dict(threshold=float(np.float32(t)), color=c) for t, c in ((-0.2, (255, 120, 80)), (0.15, (80, 120, 255)))]
        for _ in range(2):                                   # the second frame finds the grown buffers
            img, want, depth, odepth, _ = _render_both(own, scene, [f], g, contours, scenes.box_lines(nitems=8, width=2.0))
            _assert_image(img, want)
            assert np.array_equal(depth, odepth)
        assert sum(own.timings()["active_cubes"]) > 300
    finally:
        own.destroy()


def test_dropin_contour3d_execute_host_buffers(ctx, tmp_path):
    """the public entry point, host double arrays in, PPM file out; config 1 in miniature"""
    g = load_golden("frame_sines24.npz")
    grids = [g["gx"], g["gy"], g["gz"]]
    out = str(tmp_path / "out.ppm")
    rc = api.execute(g["scene"], grids, [np.ascontiguousarray(g["field"], dtype=np.float64)], g["contours"], g["lines"], out)
    assert rc == 0
    raw = open(out, "rb").read()
    h, w, _ = g["image"].shape
    assert raw.startswith(f"P6\n{w} {h}\n255\n".encode())
    _assert_image(api.read_ppm(out), g["image"])


@pytest.mark.parametrize("kind,name", [(1, "cylindrical"), (2, "stretched"), (3, "twisted")])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_dropin_coordinate_converter_matches_the_reference(ctx, tmp_path, kind, name, dtype):
    """Non-identity `converter` callbacks (include/contour3d.h:31-33): the SAME C program on the public API
    (c3d_case_main.c, converters built in) linked against the unmodified reference (CPU) and against the drop-in
    (callback tabulated on the host, kernels on the GPU) must write the same PPM, byte for byte."""
    import subprocess

    from oracle import reference

    if not reference.available():
        pytest.skip("oracle/_ref not built")
    n = 36
    f = scenes.noise(n, seed=11 + kind).astype(dtype)
    g = scenes.uniform_grids(n, n, n)
    scene = scenes.scene(400, 225)
    scene["converter"] = kind
    contours = [dict(field=0, threshold=float(np.float32(-0.15)), color=(250, 200, 60)), dict(field=0, threshold=float(np.float32(0.2)), color=(60, 200, 250))]
    case = str(tmp_path / f"{name}.bin")
    casefile.write_case(case, scene, g, [f], contours, scenes.box_lines(nitems=24, width=3.0))
    want_ppm, got_ppm = str(tmp_path / "ref.ppm"), str(tmp_path / "ours.ppm")
    reference.run_case(case, want_ppm, ranks=2, dims=(1, 2, 1))
    exe = os.path.join(os.path.dirname(api.__file__), "c3d_case")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_PORT", "MASTER_ADDR")}
    env["MINIMPI_NP"] = "1"
    proc = subprocess.run([exe, case, got_ppm], env=env, capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout[-1000:] + proc.stderr[-2000:]
    want, got = api.read_ppm(want_ppm), api.read_ppm(got_ppm)
    assert (want != np.array(scene.get("bg", (0, 0, 0)), dtype=np.uint8)).any(), "the converted surface is off screen: the case tests nothing"
    _assert_image(got, want)
    # and it is not the identity image
    scene["converter"] = 0
    casefile.write_case(case, scene, g, [f], contours, scenes.box_lines(nitems=24, width=3.0))
    reference.run_case(case, want_ppm, ranks=1)
    assert (api.read_ppm(want_ppm) != want).any()


@pytest.mark.parametrize("kind", [1, 3])
def test_frame_converter_tables_through_the_c_abi(ctx, kind):
    """c3d_contour_t.positions / c3d_line_t.points handed in directly (what a non-C host would do): the tables come from
    c3d_positions_probe on the built-in cylindrical / twisted converter, the image must equal the oracle's, which calls
    the converter itself (pinned to the reference by tests/test_oracle_vs_reference.py)."""
    n = 28
    f = scenes.noise(n, seed=40 + kind)
    g = scenes.uniform_grids(n, n, n)
    scene = scenes.scene(320, 180)
    ident, tables, strides = api.positions_probe(lambda x, y, z: tuple(oracle.convert_points(kind, [[x, y, z]])[0]), g, (0, 0, 0), (n, n, n))
    assert not ident
    lines = scenes.box_lines(nitems=10, width=2.0)
    glines = []
    for l in lines:
        e0, e1 = (np.array(v, dtype=np.float64) for v in l["edges"])
        delta = (1.0 / (l["nitems"] - 1)) * (e1 - e0)                    # line.c:138-140
        pts = np.stack([e0 + float(m) * delta for m in range(l["nitems"])])   # line.c:157-158
        glines.append(dict(l, points=oracle.convert_points(kind, pts)))
    thr = [float(np.float32(-0.15)), float(np.float32(0.2))]
    cols = [(250, 200, 60), (60, 200, 250)]
    dev = ctx.field((n, n, n), dtype=f.dtype).upload(f)
    try:
        cs = [dict(field=dev, grids=g, threshold=t, color=c, positions=(tables, strides)) for t, c in zip(thr, cols)]
        img = ctx.render(scene, cs, glines)
        depth, _ = ctx.canvas(scene["width"], scene["height"])
    finally:
        dev.destroy()
    ocs = [dict(field=f, grids=g, threshold=t, color=c, converter=kind) for t, c in zip(thr, cols)]
    rgb, odepth, _ = oracle.frame(scene, ocs, [dict(l, converter=kind) for l in lines], nthreads=4)
    assert (rgb != 0).any()
    _assert_image(img, rgb[::-1])
    assert np.array_equal(depth, odepth)


def test_dropin_rejects_what_it_cannot_honour(ctx, tmp_path, capfd):
    g = load_golden("frame_gyroid32.npz")
    grids = [g["gx"], g["gy"], g["gz"]]
    arr = np.ascontiguousarray(g["field"], dtype=np.float64)
    bad_line = [dict(edges=((0, 0, 0), (1, 1, 1)), nitems=1, width=1.0)]
    rc = api.execute(g["scene"], grids, [arr], g["contours"], bad_line, str(tmp_path / "x.ppm"))
    assert rc == 1                                   # nitems < 2 is an error in the reference too (line.c:144-147)
    assert "[CONTOUR3D ERROR]" in capfd.readouterr().err
    rc = api.execute(g["scene"], grids, [arr], g["contours"], [], str(tmp_path / "no" / "such" / "dir" / "x.ppm"))
    assert rc == 1                                   # fopen failure (output.c:126-130)


def test_demo_config1_md5(ctx, tmp_path):
    """BASELINE config 1: the reference's built-in demo; the PPM must have the reference's md5"""
    demo_case = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "_ref", "demo_case")
    if not os.path.exists(demo_case):
        pytest.skip("oracle/_ref/demo_case not built")
    import subprocess

    case = str(tmp_path / "demo.bin")
    subprocess.run([demo_case, case], check=True)
    scene, grids, fields, contours, lines = casefile.read_case(case)
    out = str(tmp_path / "demo.ppm")
    rc = api.execute(scene, grids, fields, contours, lines, out)
    assert rc == 0
    golden = json.load(open(os.path.join(GOLDEN, "demo_md5.json")))["md5"]
    assert hashlib.md5(open(out, "rb").read()).hexdigest() == golden["a.out ranks=1"]


def test_full_size_properties_512(ctx):
    """config 2 at full size (gyroid 512^3 fp32, generated on the device): size-independent checks"""
    n = 512
    f = ctx.field((n, n, n), dtype=np.float32).synthesize(0)
    try:
        thr = float(np.float32(0.1))
        entries = ctx.sweep(f, thr, rendered_only=True)
        # every entry is unique, inside the rendered range, and its corner bits are a cut pattern
        cubes = entries >> np.uint64(8)
        code = (entries & np.uint64(255)).astype(np.int64)
        assert len(np.unique(cubes)) == len(cubes)
        assert ((code != 0) & (code != 255)).all()
        c = n - 1
        i, j, k = cubes % c, (cubes // c) % c, cubes // (c * c)
        for v in (i, j, k):
            assert v.min() >= 1 and v.max() <= n - 3
        # idempotence: a second sweep gives the same set
        assert np.array_equal(np.sort(entries), np.sort(ctx.sweep(f, thr, rendered_only=True)))
        # cross-check a slab against the oracle on the downloaded bits
        host = f.download_ext()
        sub = host[100:108, 200:232, 64:200]
        g = scenes.uniform_grids(n, n, n)
        ocodes, ocounts = oracle.case_codes(sub, [g[0][64:200], g[1][200:232], g[2][100:108]], thr)
        sel = (k >= 100) & (k < 107) & (j >= 200) & (j < 231) & (i >= 64) & (i < 199)
        assert sel.sum() == int((ocounts > 0).sum())
        # frame renders and is deterministic
        scene = scenes.scene(1920, 1080)
        cs = [dict(field=f, grids=g, threshold=thr, color=(255, 200, 100))]
        a = ctx.render(scene, cs, [])
        b = ctx.render(scene, cs, [])
        assert np.array_equal(a, b)
        assert (a.reshape(-1, 3).max(axis=1) > 0).mean() > 0.2
    finally:
        f.destroy()


def test_full_size_visibility_shortcuts_change_nothing():
    """config 2's size (512^3, 1920x1080), five cameras: the frame with every visibility shortcut of the raster (front-to-back
    order, cube / triangle occlusion tests, candidate pixel boxes) against the frame without them (CONTOUR3D_NO_CULL):
    depth, winner key and colour of every pixel must be identical.  The oracle cannot referee frames of this size in
    seconds; the shortcut-free pipeline is what the small-size parity tests pin to it."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("gpu_shortcut_check", os.path.join(os.path.dirname(GOLDEN), "..", "tools", "gpu_shortcut_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res = mod.frames(["CONTOUR3D_NO_CULL"], 512, 1920, 1080)
    assert len(res) == 5
    for name, r in res.items():
        (ia, da, ka, _, dropped), (ib, db, kb, _, none) = r["on"], r["off"]
        assert dropped > 0 and none == 0, name
        assert np.array_equal(ka, kb) and np.array_equal(da, db, equal_nan=True) and np.array_equal(ia, ib), name
        assert (ka != 0).mean() > 0.1, name


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_frame_tiny_magnitude_field_takes_the_absolute_branch(ctx, dtype):
    """A field whose values lie within ~1e-8 of the iso-value everywhere near the surface: the reference's ABSOLUTE branch
    |v1 - v0| < 1e-8 puts every vertex on the tail corner (triangulate.c:62-71), all triangles collapse and are drawn as
    infinite-depth blobs.  The raster's visibility shortcuts must leave such cubes alone (patch_irregular's absolute test)."""
    n = 24
    f = (scenes.noise(n, seed=31).astype(np.float64) * 1e-9).astype(dtype)
    g = scenes.uniform_grids(n, n, n)
    scene = scenes.scene(320, 180)
    thr = float(dtype(1e-10))
    # a second, ordinary surface in front: its occlusion must not remove the blobs behind it
    big = scenes.noise(n, seed=32).astype(dtype)
    contours = [dict(field=1, threshold=float(dtype(-0.3)), color=(250, 160, 60)), dict(field=0, threshold=thr, color=(60, 160, 250))]
    img, want, depth, odepth, _ = _render_both(ctx, scene, [f, big], g, contours, [])
    _assert_image(img, want)
    assert np.array_equal(depth, odepth)
    assert np.isinf(odepth).any(), "the scene was meant to contain collapsed triangles"


def test_dropin_device_array_is_swept_in_place(ctx, tmp_path):
    """contour3d_execute with a DEVICE double array (a GPU-resident solver): one rank needs no halo room, so the array is
    borrowed (c3d_field_wrap) instead of copied; the image is the host-array image"""
    import torch

    g = load_golden("frame_sines24.npz")
    grids = [g["gx"], g["gy"], g["gz"]]
    host = np.ascontiguousarray(g["field"], dtype=np.float64)
    dev = torch.from_numpy(host).cuda()
    before = dev.clone()
    out_h, out_d = str(tmp_path / "host.ppm"), str(tmp_path / "dev.ppm")
    assert 0 == api.execute(g["scene"], grids, [host], g["contours"], g["lines"], out_h)
    assert 0 == api.execute(g["scene"], grids, [int(dev.data_ptr())], g["contours"], g["lines"], out_d)
    assert open(out_h, "rb").read() == open(out_d, "rb").read()
    _assert_image(api.read_ppm(out_d), g["image"])
    assert torch.equal(dev, before), "the caller's array must not be written"


@pytest.mark.parametrize("kind,seed,octaves,gain", [(1, 1234, 5, 0.5), (1, 7, 7, 2.0 ** (-5.0 / 6.0)), (2, 2, 5, 0.5)])
def test_host_generator_is_the_device_generator(ctx, kind, seed, octaves, gain):
    """bench.py's reference arm takes its field from libc3d_synth_host.so (plain C, no CUDA); the GPU arm from
    c3d_field_synthesize.  Same bits, or the two arms would not render the same job."""
    n = 40
    f = ctx.field((n, n, n), dtype=np.float32).synthesize(kind=kind, seed=seed, octaves=octaves, gain=gain)
    try:
        dev = f.download_ext()
    finally:
        f.destroy()
    host = scenes.synth_host((n, n, n), kind=kind, seed=seed, octaves=octaves, gain=gain)
    assert np.array_equal(dev, host), f"{int((dev != host).sum())} of {dev.size} values differ (max {np.abs(dev - host).max()})"
    # a window of a larger field is that field's window
    win = scenes.synth_host((n, n, n), kind=kind, seed=seed, octaves=octaves, gain=gain, window=(8, 9, 10), offset=(5, 6, 7))
    assert np.array_equal(win, host[7:17, 6:15, 5:13])


_FUZZ = int(os.environ.get("C3D_GPU_FUZZ", "0") or 0)


@pytest.mark.skipif(_FUZZ <= 0, reason="opt-in: C3D_GPU_FUZZ=<number of seeds> (written after this round's last GPU visit; the same cases pin the oracle to the reference on the CPU)")
def test_random_frames_match_the_oracle(ctx):
    """the seeded random frames of tests/fuzz_cases.py (ragged fp32 / fp64 arrays, non-uniform coordinates, exact hits, cameras
    from inside the volume to far away, tiny and skewed screens, random lines) through the CUDA path: image and z-buffer
    equal to the oracle's, every pixel"""
    from fuzz_cases import random_case

    bad = []
    for seed in range(1000, 1000 + _FUZZ):
        scene, grids, f, contours, lines, _, _ = random_case(seed)
        img, want, depth, odepth, _ = _render_both(ctx, scene, [f], grids, contours, lines)
        if not (np.array_equal(img, want) and np.array_equal(depth, odepth, equal_nan=True)):
            bad.append((seed, int((img != want).any(axis=2).sum())))
    assert not bad, f"(seed, differing pixels): {bad}"
