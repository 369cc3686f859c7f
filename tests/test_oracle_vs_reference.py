"""CPU: pin the oracle restatement to the UNMODIFIED reference (oracle/_ref).

Skipped when oracle/_ref is absent (it is built from /root/reference by
oracle/Makefile in the development container and travels to the GPU box as
prebuilt binaries; nothing here reads /root/reference at run time)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_records_equal, image_diff
from contour3d_b200 import casefile, scenes
from contour3d_b200.api import read_ppm
from oracle import oracle, reference

pytestmark = pytest.mark.skipif(not reference.available(), reason="oracle/_ref not built")


@pytest.mark.parametrize("shape,thr,seed", [((9, 8, 7), 0.05, 0), ((16, 16, 16), -0.3, 1), ((5, 21, 6), 0.2, 2), ((4, 4, 4), 0.0, 3)])
def test_stage_level_triangles(shape, thr, seed):
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    grids = [np.sort(rng.uniform(-0.5, 0.5, n)) for n in (nx, ny, nz)]
    field = rng.normal(size=shape) * 0.3
    assert_records_equal(oracle.tessellate(field, grids, thr)[0], reference.tessellate(field, grids, thr), str(shape))


def test_stage_level_smooth_field():
    f, g = scenes.sines(20)
    assert_records_equal(oracle.tessellate(f, g, 0.1)[0], reference.tessellate(f, g, 0.1), "sines20")


def test_demo_md5_pins_the_toolchain(tmp_path):
    """config 1: the reference's own a.out, the case-file path and the oracle all give the same image"""
    golden = json.load(open(os.path.join(GOLDEN, "demo_md5.json")))["md5"]
    case = str(tmp_path / "demo32.bin")
    reference.write_demo_case(case, 32, 320, 240)
    out = str(tmp_path / "ref.ppm")
    reference.run_case(case, out, ranks=2)
    md5_ref = hashlib.md5(open(out, "rb").read()).hexdigest()
    assert md5_ref == golden["ref_driver(demo_case 32 320 240)"]
    scene, grids, fields, contours, lines = casefile.read_case(case)
    cs = [dict(field=fields[c["field"]], grids=grids, threshold=c["threshold"], color=c["color"]) for c in contours]
    rgb, _, _ = oracle.frame(scene, cs, lines, nthreads=2)
    mine = str(tmp_path / "oracle.ppm")
    oracle.write_ppm(mine, rgb)
    assert hashlib.md5(open(mine, "rb").read()).hexdigest() == md5_ref


@pytest.mark.parametrize("ranks,dims", [(1, None), (2, (1, 2, 1)), (4, (1, 1, 4)), (6, (1, 3, 2))])
def test_reference_is_decomposition_invariant_and_matches_oracle(tmp_path, ranks, dims):
    f, g = scenes.sines(28, seed=11)
    scene = scenes.scene(200, 150)
    contours = [dict(field=0, threshold=0.2, color=(255, 128, 0)), dict(field=0, threshold=-0.2, color=(0, 128, 255))]
    lines = scenes.box_lines(nitems=16, width=3.0)
    case, out = str(tmp_path / "c.bin"), str(tmp_path / "o.ppm")
    casefile.write_case(case, scene, g, [f], contours, lines)
    reference.run_case(case, out, ranks=ranks, dims=dims)
    cs = [dict(field=f, grids=g, threshold=c["threshold"], color=c["color"]) for c in contours]
    rgb, _, _ = oracle.frame(scene, cs, lines, nthreads=1)
    frac, worst = image_diff(read_ppm(out), rgb[::-1])
    assert frac == 0.0, (frac, worst)


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_coordinate_converters_oracle_matches_reference(tmp_path, kind):
    """non-identity `converter` callbacks (triangulate.c:165-170, line.c:39-40): the oracle calling the built-in
    cylindrical / stretched / twisted functions == the unmodified reference calling the same functions (ref_driver)"""
    from contour3d_b200 import api, casefile, scenes

    n = 20
    f = scenes.noise(n, seed=30 + kind)
    g = scenes.uniform_grids(n, n, n)
    scene = scenes.scene(200, 150)
    scene["converter"] = kind
    contours = [dict(field=0, threshold=float(np.float32(-0.1)), color=(250, 200, 60)), dict(field=0, threshold=float(np.float32(0.2)), color=(60, 200, 250))]
    lines = scenes.box_lines(nitems=12, width=2.0)
    case, ppm = str(tmp_path / "c.bin"), str(tmp_path / "ref.ppm")
    casefile.write_case(case, scene, g, [f], contours, lines)
    reference.run_case(case, ppm, ranks=2, dims=(1, 1, 2))
    want = api.read_ppm(ppm)
    ocs = [dict(field=f, grids=g, threshold=c["threshold"], color=c["color"], converter=kind) for c in contours]
    ols = [dict(l, converter=kind) for l in lines]
    rgb, _, _ = oracle.frame(scene, ocs, ols, nthreads=3)
    assert (want != 0).any()
    assert np.array_equal(rgb[::-1], want)


@pytest.mark.parametrize("ranks,dims,flags", [(4, (0, 0), 1), (8, (0, 0), 3), (4, (4, 1), 2)])
def test_reference_with_default_grids_periodic_and_x2_pencils(tmp_path, ranks, dims, flags):
    """what tests/test_gpu_multirank.py::test_dropin_against_the_genuine_simpledecomp compares with: the unmodified reference
    itself gives the one-rank image on SimpleDecomp's default process grids, periodic constructs and X2 pencils"""
    import subprocess

    from contour3d_b200 import api, casefile, scenes

    n = 24
    f = scenes.noise(n, seed=3)
    g = scenes.uniform_grids(n, n, n)
    scene = scenes.scene(160, 90)
    contours = [dict(field=0, threshold=float(np.float32(0.05)), color=(250, 200, 60))]
    case = str(tmp_path / "case.bin")
    casefile.write_case(case, scene, g, [f], contours, scenes.box_lines(nitems=8, width=2.0))
    one, many = str(tmp_path / "one.ppm"), str(tmp_path / "many.ppm")
    reference.run_case(case, one, ranks=1)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_PORT", "MASTER_ADDR")}
    env["MINIMPI_NP"] = str(ranks)
    proc = subprocess.run([reference.REF_DRIVER, case, many, str(dims[0]), str(dims[1]), "1", "0", str(flags)], env=env, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert np.array_equal(api.read_ppm(one), api.read_ppm(many))


@pytest.mark.parametrize("seed", range(1000, 1120))
def test_random_frames_oracle_matches_reference(tmp_path, seed):
    """fuzz (tests/fuzz_cases.py): ragged shapes, fp32 / fp64, non-uniform coordinates, iso-values that hit array values exactly,
    cameras from inside the volume to far away, tiny and skewed screens, random lines, 1-3 ranks of the reference: the oracle's
    image == the unmodified reference's, every pixel.  (400 further seeds were run once when this was written: no difference.)"""
    from fuzz_cases import random_case

    scene, grids, f, contours, lines, ranks, dims = random_case(seed)
    case, out = str(tmp_path / "c.bin"), str(tmp_path / "o.ppm")
    casefile.write_case(case, scene, grids, [f], contours, lines)
    reference.run_case(case, out, ranks=ranks, dims=dims, timeout=300)
    cs = [dict(field=f, grids=grids, threshold=c["threshold"], color=c["color"]) for c in contours]
    rgb, _, _ = oracle.frame(scene, cs, lines, nthreads=1 + seed % 3)
    assert np.array_equal(read_ppm(out), rgb[::-1])


@pytest.mark.parametrize("seed", range(2000, 2060))
def test_random_frames_with_converters_oracle_matches_reference(tmp_path, seed):
    """the same fuzz with a non-identity `converter` (cylindrical / stretched / twisted, by seed) on contours and lines"""
    from fuzz_cases import random_case

    scene, grids, f, contours, lines, ranks, dims = random_case(seed)
    kind = 1 + seed % 3
    scene["converter"] = kind
    case, out = str(tmp_path / "c.bin"), str(tmp_path / "o.ppm")
    casefile.write_case(case, scene, grids, [f], contours, lines)
    reference.run_case(case, out, ranks=ranks, dims=dims, timeout=300)
    cs = [dict(field=f, grids=grids, threshold=c["threshold"], color=c["color"], converter=kind) for c in contours]
    ls = [dict(l, converter=kind) for l in lines]
    rgb, _, _ = oracle.frame(scene, cs, ls, nthreads=1 + seed % 3)
    assert np.array_equal(read_ppm(out), rgb[::-1])
