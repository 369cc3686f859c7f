"""In-tree build of the native libraries (nvcc for sm_100a + gcc).

    python -m contour3d_b200.build [--force]

Outputs (git-ignored, but they travel to the GPU box with the snapshot):
  contour3d_b200/libcontour3d_b200.so     kernels + runtime + contour3d_execute
  contour3d_b200/libc3d_sdecomp_compat.so SimpleDecomp stand-in + minimpi (hosts without MPI)
  contour3d_b200/c3d_case                 case-file demo program on the public API
  contour3d_b200/libc3d_synth_host.so     host twin of the device field generator (bench.py's reference arm; no CUDA)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
MINIMPI = os.path.join(ROOT, "compat", "minimpi")
OBJ = os.path.join(PKG, "_build")

LIB = os.path.join(PKG, "libcontour3d_b200.so")
COMPAT = os.path.join(PKG, "libc3d_sdecomp_compat.so")
CASE = os.path.join(PKG, "c3d_case")
SYNTH = os.path.join(PKG, "libc3d_synth_host.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # fp64 geometry must round exactly like the reference's -std=c99 build: no FMA contraction
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA toolkit is required to build contour3d_b200 (no CPU fallback exists)")


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        raise RuntimeError(f"build step failed: {cmd[0]}")
    return proc


def have_real_mpi() -> bool:
    return shutil.which("mpicc") is not None


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, h) for h in ("c3d_internal.h", "c3d_tables.h", "c3d_device.cuh", "c3d_peer.h", "c3d_synth.h")]
    headers += [os.path.join(INCLUDE, h) for h in ("contour3d_b200.h", "contour3d.h", "sdecomp.h")]
    headers += [os.path.join(MINIMPI, "mpi.h")]
    inc = ["-I", INCLUDE, "-I", CSRC]
    # on an MPI-less host the public header's <mpi.h> resolves to the stand-in
    cc = "mpicc" if have_real_mpi() else "gcc"
    mpi_inc = [] if have_real_mpi() else ["-I", MINIMPI]
    objs = []
    for name in ("c3d_kernels.cu", "c3d_sweep.cu", "c3d_shade.cu", "c3d_peer.cu", "c3d_runtime.cu"):
        src = os.path.join(CSRC, name)
        obj = os.path.join(OBJ, name + ".o")
        if force or _newer(obj, [src] + headers):
            extra = ["-Xptxas", "-v"] if verbose else []
            p = _run([nvcc] + NVCC_FLAGS + extra + inc + ["-c", src, "-o", obj])
            if verbose:
                sys.stderr.write(p.stderr)
        objs.append(obj)
    for name in ("contour3d_execute.c", "c3d_positions.c", "c3d_camera.c"):
        src = os.path.join(CSRC, name)
        obj = os.path.join(OBJ, name[:-2] + ".o")
        if force or _newer(obj, [src] + headers):
            # -ffp-contract=off: the line points of a converted poly-line follow the reference's -std=c99 arithmetic
            _run([cc, "-std=c99", "-O2", "-ffp-contract=off", "-Wall", "-Wextra", "-fPIC"] + inc + mpi_inc + ["-c", src, "-o", obj])
        objs.append(obj)
    if force or _newer(LIB, objs):
        # `sdecomp` and the MPI_* symbols stay undefined: the embedding program
        # (or libc3d_sdecomp_compat.so) provides them
        _run([nvcc, "-shared", "-o", LIB] + objs + ["-ldl", "-Xlinker", "--allow-shlib-undefined"])
    # SimpleDecomp stand-in (the getters contour3d_execute consumes + sdecomp.construct for the demo programs).  Without an MPI
    # installation it carries the MPI stand-in too; with one it is compiled by mpicc against the real <mpi.h> (it only needs
    # MPI_Cart_* / MPI_Dims_create), so that c3d_case links either way.  A solver that brings the genuine SimpleDecomp links that instead.
    if have_real_mpi():
        srcs = [os.path.join(CSRC, "c3d_sdecomp_compat.c")]
        if force or _newer(COMPAT, srcs + headers):
            _run(["mpicc", "-std=gnu11", "-O2", "-Wall", "-Wextra", "-fPIC", "-shared", "-D_GNU_SOURCE"] + inc + ["-o", COMPAT] + srcs)
    else:
        srcs = [os.path.join(CSRC, "c3d_sdecomp_compat.c"), os.path.join(MINIMPI, "minimpi.c")]
        if force or _newer(COMPAT, srcs + headers):
            _run(["gcc", "-std=gnu11", "-O2", "-Wall", "-Wextra", "-fPIC", "-shared", "-D_GNU_SOURCE"] + inc + mpi_inc
                 + ["-o", COMPAT] + srcs + ["-lpthread", "-lrt"])
    synth_src = os.path.join(CSRC, "c3d_synth_host.c")
    if force or _newer(SYNTH, [synth_src, os.path.join(CSRC, "c3d_synth.h")]):
        # same single-precision arithmetic as the kernel: no contraction into fused multiply-adds
        _run(["gcc", "-std=c99", "-O3", "-ffp-contract=off", "-Wall", "-Wextra", "-fPIC", "-shared", "-I", CSRC, "-o", SYNTH, synth_src, "-lpthread", "-lm"])
    src = os.path.join(CSRC, "c3d_case_main.c")
    if force or _newer(CASE, [src, LIB, COMPAT]):
        link = ["-L", PKG, "-lcontour3d_b200", "-lc3d_sdecomp_compat", "-Wl,-rpath,$ORIGIN"]
        _run([cc, "-std=c99", "-O3", "-D_POSIX_C_SOURCE=200809L"] + inc + mpi_inc + ["-o", CASE, src] + link + ["-lm"])  # the reference driver's flags (oracle/Makefile): the built-in converters must round alike
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
