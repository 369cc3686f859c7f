"""CPU: the oracle restatement against golden vectors produced by the unmodified reference."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_records_equal, image_diff, load_golden
from oracle import oracle


@pytest.mark.parametrize("name", ["tess_ragged.npz", "tess_exact_hits.npz"])
def test_triangles_match_reference(name):
    g = load_golden(name)
    got, stats = oracle.tessellate(g["field"], [g["gx"], g["gy"], g["gz"]], g["threshold"])
    assert_records_equal(got, g["triangles"], name)
    assert stats["triangles"] == len(g["triangles"])


def test_exact_hits_reproduce_nan_normals():
    # a grid value equal to the threshold at an isolated corner collapses a triangle: area 0, normal NaN
    g = load_golden("tess_exact_hits.npz")
    tri = g["triangles"]
    zero = tri["area"] == 0
    assert zero.any()
    assert np.isnan(tri["fn"][zero]).all()


def test_case_codes_consistent_with_triangles():
    g = load_golden("tess_ragged.npz")
    codes, counts = oracle.case_codes(g["field"], [g["gx"], g["gy"], g["gz"]], g["threshold"])
    per_cube = np.bincount(g["triangles"]["cube"].astype(np.int64), minlength=counts.size)
    assert np.array_equal(per_cube, counts.reshape(-1))
    # triangle count per tetra mask: {0,1,1,2,1,2,2,1} after folding masks >= 8
    lut = np.array([0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0])
    total = sum(lut[(codes >> (4 * n)) & 15] for n in range(6))
    assert np.array_equal(total, counts)


@pytest.mark.parametrize("name,threads", [("frame_sines24.npz", 1), ("frame_sines24.npz", 3), ("frame_gyroid32.npz", 1), ("frame_gyroid32.npz", 2)])
def test_frames_match_reference(name, threads):
    g = load_golden(name)
    grids = [g["gx"], g["gy"], g["gz"]]
    cs = [dict(field=g["field"], grids=grids, threshold=c["threshold"], color=c["color"]) for c in g["contours"]]
    rgb, depth, stats = oracle.frame(g["scene"], cs, g["lines"], nthreads=threads)
    frac, worst = image_diff(rgb[::-1], g["image"])
    assert frac == 0.0 and worst == 0, (frac, worst)


def test_write_ppm_layout(tmp_path):
    g = load_golden("frame_gyroid32.npz")
    out = tmp_path / "o.ppm"
    oracle.write_ppm(str(out), np.ascontiguousarray(g["image"][::-1]))
    raw = out.read_bytes()
    h, w, _ = g["image"].shape
    head = f"P6\n{w} {h}\n255\n".encode()
    assert raw.startswith(head) and raw[len(head):] == g["image"].tobytes()
