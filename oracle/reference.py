"""Access to the UNMODIFIED reference built into oracle/_ref by oracle/Makefile.

TEST / BASELINE INFRASTRUCTURE ONLY (tests/, tests/golden/make_golden.py, bench.py's
reference arm).  Nothing here reads /root/reference at run time: only the
prebuilt binaries under oracle/_ref are used, so it also works on the GPU box.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

from . import oracle as _oracle

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_LIB = os.path.join(REF_DIR, "libc3dref.so")
REF_DRIVER = os.path.join(REF_DIR, "ref_driver")
REF_DEMO = os.path.join(REF_DIR, "ref_demo")
DEMO_CASE = os.path.join(REF_DIR, "demo_case")


def build():
    """(Re)build oracle/_ref when the reference tree is present; otherwise keep what is there."""
    subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def available() -> bool:
    return all(os.path.exists(p) for p in (REF_LIB, REF_DRIVER, DEMO_CASE))


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_LIB)
        L.ref_tessellate.restype = C.c_size_t
        L.ref_tessellate.argtypes = [C.POINTER(C.c_size_t * 3), C.POINTER(C.POINTER(C.c_double) * 3), C.POINTER(C.c_double),
                                     C.c_double, C.c_size_t, C.c_void_p]
        _lib = L
    return _lib


def tessellate(field, grids, threshold) -> np.ndarray:
    """The reference's triangulate_slice + compute_vertex_normals over a whole array -> records."""
    f = np.ascontiguousarray(field, dtype=np.float64)
    nz, ny, nx = f.shape
    g = [np.ascontiguousarray(x, dtype=np.float64) for x in grids]
    dims = (C.c_size_t * 3)(nx, ny, nz)
    garr = (C.POINTER(C.c_double) * 3)(*[x.ctypes.data_as(C.POINTER(C.c_double)) for x in g])
    fptr = f.ctypes.data_as(C.POINTER(C.c_double))
    total = lib().ref_tessellate(C.byref(dims), C.byref(garr), fptr, float(threshold), 0, None)
    out = np.zeros(total, dtype=_oracle.TRIANGLE_DTYPE)
    if total:
        lib().ref_tessellate(C.byref(dims), C.byref(garr), fptr, float(threshold), total, out.ctypes.data)
    return out


def run_case(case_path: str, out_ppm: str, ranks: int = 1, dims=None, repeat: int = 1, warmup: int = 0, timeout=None) -> dict:
    """contour3d_execute of the reference on a case file, `ranks` processes (fork mode of compat/minimpi)."""
    cmd = [REF_DRIVER, case_path, out_ppm]
    if dims is not None:
        cmd += [str(dims[1]), str(dims[2])]
        if repeat != 1 or warmup:
            cmd += [str(repeat), str(warmup)]
    elif repeat != 1 or warmup:
        raise ValueError("repeat needs explicit dims")
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_PORT", "MASTER_ADDR"):
        env.pop(k, None)
    env["MINIMPI_NP"] = str(ranks)
    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    if proc.returncode != 0:
        raise RuntimeError(f"ref_driver failed ({proc.returncode}): {proc.stderr[-2000:]}")
    return json.loads(proc.stdout.strip().splitlines()[-1])


def write_demo_case(path: str, n: int = 128, width: int = 1440, height: int = 1080):
    subprocess.run([DEMO_CASE, path, str(n), str(width), str(height)], check=True)


def run_demo(workdir: str, ranks: int = 1) -> str:
    """The reference's own ./a.out (src/main.c) -> path of output.ppm"""
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_PORT", "MASTER_ADDR"):
        env.pop(k, None)
    env["MINIMPI_NP"] = str(ranks)
    subprocess.run([REF_DEMO], cwd=workdir, env=env, check=True)
    return os.path.join(workdir, "output.ppm")
