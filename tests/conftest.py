import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """the single-rank parity tests first, the multi-rank ones (one process per rank, minutes) last: with -x a failure there
    must not hide what the single-rank path has already shown"""
    items.sort(key=lambda item: 1 if "test_gpu_multirank" in item.nodeid else 0)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    out = {k: z[k] for k in z.files}
    for k in ("scene", "contours", "lines"):
        if k in out:
            out[k] = json.loads(str(out[k]))
    if "threshold" in out:
        out["threshold"] = float(out["threshold"])
    return out


def assert_records_equal(a, b, what=""):
    """bit-exact comparison of two triangle record arrays (NaN == NaN); vn only where has_vn"""
    assert len(a) == len(b), f"{what}: {len(a)} vs {len(b)} triangles"
    for name in ("cube", "tri", "has_vn", "edge"):
        assert np.array_equal(a[name], b[name]), f"{what}: {name} differs"
    for name in ("area", "v", "fn"):
        assert np.array_equal(a[name], b[name], equal_nan=True), f"{what}: {name} differs"
    m = a["has_vn"] == 1
    assert np.array_equal(a["vn"][m], b["vn"][m], equal_nan=True), f"{what}: vertex normals differ"


def image_diff(a, b):
    """(fraction of pixels that differ, max per-channel difference)"""
    assert a.shape == b.shape, (a.shape, b.shape)
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    return float((d.max(axis=2) > 0).mean()), int(d.max())


@pytest.fixture(scope="session")
def ctx():
    from contour3d_b200 import api

    c = api.Context()
    yield c
    c.destroy()
