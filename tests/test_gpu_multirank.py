"""GPU, N > 1: halo exchange, screen-partitioned frames over peer memory and whole-image reduction over NCCL.

One process per rank.  With fewer GPUs than ranks several ranks share a GPU: CUDA IPC maps device memory between
processes on one device just as between devices, so the peer-memory path -- the default data path of multi-rank
frames -- is exercised on a single-GPU box too.  Only the NCCL path (coordinate converters, CONTOUR3D_P2P=0) needs
one GPU per rank and is skipped otherwise."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run_check(world, dims, env_extra):
    port = 29500 + (os.getpid() + world * 7 + len(dims) + 13 * len(env_extra)) % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py"), dims]
    env = dict(os.environ)
    env.update(env_extra)
    env.setdefault("CONTOUR3D_PEER_TIMEOUT", "120")
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "ok" in proc.stdout
    return proc.stdout


@pytest.mark.parametrize("world,dims", [(2, "1x1x2"), (2, "1x2x1"), (4, "1x2x2"), (4, "1x1x4"), (8, "1x4x2"), (8, "1x1x8"), (3, "1x1x3")])
def test_multi_rank_image_equals_single_rank_oracle(world, dims):
    """screen-partitioned frames over peer memory (exchange mode 2); ranks share GPUs when the box has fewer"""
    if _gpus() < 1:
        pytest.skip("needs a GPU")
    extra = {"C3D_EXPECT_MODE": "2"}
    if dims == "1x1x2":
        # also the grow-and-redo path: the ranks overflow their tiny work buffers by different amounts and must agree to run again
        extra["CONTOUR3D_LIST_CAP"] = "2000"
    if dims == "1x2x2":
        extra["CONTOUR3D_TILE_SHIFT"] = "4"       # 16-pixel tiles: many cubes straddle tiles of different ranks
    _run_check(world, dims, extra)


@pytest.mark.parametrize("world,dims", [(2, "1x1x2"), (2, "1x2x1"), (4, "1x2x2"), (8, "1x4x2")])
def test_multi_gpu_nccl_path_image_equals_single_rank_oracle(world, dims):
    """CONTOUR3D_P2P=0: every rank draws its own cubes, halos and whole images travel through NCCL (exchange mode 1)"""
    if _gpus() < world:
        pytest.skip(f"the NCCL path needs one GPU per rank ({world})")
    extra = {"CONTOUR3D_P2P": "0", "C3D_EXPECT_MODE": "1"}
    if dims == "1x1x2":
        extra["CONTOUR3D_LIST_CAP"] = "2000"
    _run_check(world, dims, extra)


@pytest.mark.parametrize("kind,dims", [(0, (1, 1, 2)), (1, (1, 2, 1)), (3, (1, 1, 2))])
def test_dropin_program_two_ranks_matches_the_reference(tmp_path, kind, dims):
    """The drop-in exactly as a solver would use it: a C program on the public API (c3d_case_main.c), two MPI ranks (the
    fork-mode stand-in, compat/minimpi), one GPU each, NCCL id broadcast with MPI_Bcast inside contour3d_execute, SimpleDecomp
    getters for the pencils -- against the same program linked with the unmodified reference on the CPU.  kind != 0:
    non-identity coordinate converter, tabulated per rank over its own extended block."""
    import numpy as np

    from contour3d_b200 import api, casefile, scenes
    from oracle import reference

    if _gpus() < 1 or (kind != 0 and _gpus() < 2):
        pytest.skip("coordinate converters take the NCCL path: one GPU per rank")
    if not reference.available():
        pytest.skip("oracle/_ref not built")
    n = 40
    f = scenes.noise(n, seed=5 + kind)
    g = scenes.uniform_grids(n, n, n)
    scene = scenes.scene(400, 225)
    scene["converter"] = kind
    contours = [dict(field=0, threshold=float(np.float32(-0.1)), color=(250, 200, 60)), dict(field=0, threshold=float(np.float32(0.25)), color=(60, 200, 250))]
    case = str(tmp_path / "case.bin")
    casefile.write_case(case, scene, g, [f], contours, scenes.box_lines(nitems=24, width=3.0))
    want_ppm, got_ppm = str(tmp_path / "ref.ppm"), str(tmp_path / "ours.ppm")
    reference.run_case(case, want_ppm, ranks=1)
    exe = os.path.join(os.path.dirname(api.__file__), "c3d_case")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_PORT", "MASTER_ADDR")}
    env["MINIMPI_NP"] = "2"
    proc = subprocess.run([exe, case, got_ppm, str(dims[1]), str(dims[2])], env=env, capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout[-1000:] + proc.stderr[-2000:]
    want, got = api.read_ppm(want_ppm), api.read_ppm(got_ppm)
    assert (want != 0).any()
    assert np.array_equal(got, want), f"{int((got != want).any(axis=2).sum())} pixels differ from the reference image"


@pytest.mark.parametrize("ranks,dims,flags", [
    (2, (0, 0), 1),      # MPI_Dims_create -> {1, 2, 1}; periodic in all directions, as the reference's demo constructs it (src/main.c:119-125)
    (4, (0, 0), 1),      # {1, 2, 2}
    (8, (0, 0), 1),      # {1, 4, 2}
    (8, (0, 0), 3),      # the same grid, arrays handed over as X2 pencils: y split by 2, z by 4 (sdecomp/get.c:44-57)
    (4, (4, 1), 2),      # X2 pencils of a {1, 4, 1} grid: z slabs
    (6, (3, 2), 0),
])
def test_dropin_against_the_genuine_simpledecomp(tmp_path, ranks, dims, flags):
    """SURVEY 8(f) rank 1, as far as it goes without an MPI installation: the case-file program compiled against the
    REFERENCE's headers and linked with the GENUINE SimpleDecomp sources (sdecomp.construct -> MPI_Dims_create,
    MPI_Cart_create(reorder = 1), pencil getters through table_3d), contour3d_execute taken from this library.
    Ranks are processes of the MPI stand-in, one per GPU (sharing GPUs when the box has fewer).  The PPM must equal,
    byte for byte, what the same program linked against the unmodified reference writes on one CPU rank."""
    import numpy as np

    from contour3d_b200 import api, casefile, scenes
    from oracle import reference

    exe = os.path.join(ROOT, "oracle", "_ref", "c3d_case_sd")
    if _gpus() < 1:
        pytest.skip("needs a GPU")
    if not reference.available() or not os.path.exists(exe):
        pytest.skip("oracle/_ref not built")
    n = 48
    f = scenes.noise(n, seed=17 + ranks)
    g = scenes.uniform_grids(n, n, n)
    scene = scenes.scene(400, 225)
    contours = [dict(field=0, threshold=float(np.float32(-0.1)), color=(250, 200, 60)), dict(field=0, threshold=float(np.float32(0.25)), color=(60, 200, 250))]
    case = str(tmp_path / "case.bin")
    casefile.write_case(case, scene, g, [f], contours, scenes.box_lines(nitems=24, width=3.0))
    want_ppm, got_ppm = str(tmp_path / "ref.ppm"), str(tmp_path / "ours.ppm")
    reference.run_case(case, want_ppm, ranks=1)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_PORT", "MASTER_ADDR")}
    env["MINIMPI_NP"] = str(ranks)
    env.setdefault("CONTOUR3D_PEER_TIMEOUT", "120")
    proc = subprocess.run([exe, case, got_ppm, str(dims[0]), str(dims[1]), "2", "0", str(flags)], env=env, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-1000:] + proc.stderr[-2000:]
    assert '"status": 0' in proc.stdout
    want, got = api.read_ppm(want_ppm), api.read_ppm(got_ppm)
    assert (want != 0).any()
    assert np.array_equal(got, want), f"{int((got != want).any(axis=2).sum())} pixels differ from the reference image"
