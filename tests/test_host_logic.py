"""CPU: host-side logic and the C-ABI surface (no compute calls)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from contour3d_b200 import api, casefile, scenes


def test_library_exports_every_declared_symbol():
    declared = set()
    for header in ("contour3d_b200.h", "contour3d.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b(c3d_[a-z0-9_]+|contour3d_execute)\s*\(", text))
    declared = {d for d in declared if not d.endswith("_t")}
    assert declared, "no declarations found"
    L = api.lib()
    missing = [name for name in sorted(declared) if not hasattr(L, name)]
    assert not missing, missing
    assert set(api.EXPORTS) == declared, set(api.EXPORTS) ^ declared


def test_version_string():
    assert b"sm_100a" in api.lib().c3d_version()


def test_public_struct_layout_matches_reference_abi():
    # sizes probed from the reference on x86-64 SysV
    assert C.sizeof(api.ContourObj) == 88
    assert C.sizeof(api.LineObj) == 80
    assert C.sizeof(api.Color) == 3
    assert C.sizeof(api.Vector) == 24


def test_no_gpu_means_loud_failure_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.C3DError):
        api.Context()


@pytest.mark.parametrize("n,p", [(10, 3), (128, 8), (1024, 8), (7, 7), (100, 1)])
def test_block_decomposition_matches_simpledecomp_rule(n, p):
    # mysize = (glsize + rank) / nprocs, offset = prefix sum (sdecomp/kernel.c:89,111-116)
    total = 0
    for r in range(p):
        d = api.make_decomp((1, 1, p), r)
        my, off = api.decomp_local(d, (n, n, n))
        assert my[2] == (n + r) // p and off[2] == total
        assert my[0] == n and off[0] == 0
        total += my[2]
    assert total == n


def test_cartesian_neighbours_row_major_z_fastest():
    d = api.make_decomp((1, 3, 2), 3)  # coords (y, z) = (1, 1)
    assert list(d.myrank) == [0, 1, 1]
    assert [list(x) for x in d.neighbours] == [[-1, -1], [1, 5], [2, -1]]
    with pytest.raises(api.C3DError):
        api.make_decomp((2, 1, 1), 0)  # x is never split


def test_sdecomp_standin_getters():
    info = api.sdecomp_construct((0, 0, 0))
    assert info.value


def test_casefile_roundtrip(tmp_path):
    f = scenes.gyroid(8)
    g = scenes.uniform_grids(8, 8, 8)
    sc = scenes.scene(64, 48, bg=(1, 2, 3))
    cs = [dict(field=0, threshold=0.25, color=(9, 8, 7))]
    ls = scenes.box_lines(nitems=5, width=1.5)
    p = str(tmp_path / "c.bin")
    casefile.write_case(p, sc, g, [f], cs, ls)
    sc2, g2, f2, cs2, ls2 = casefile.read_case(p)
    assert np.array_equal(f2[0], f) and all(np.array_equal(a, b) for a, b in zip(g, g2))
    assert sc2["width"] == 64 and sc2["bg"] == (1, 2, 3) and cs2 == cs
    assert ls2[3]["nitems"] == 5 and ls2[3]["width"] == 1.5


def test_case_tables_agree_with_oracle_counts():
    """the product's 256-entry case table (c3d_tables.h) against the oracle's per-tetra classification"""
    from oracle import oracle

    # enumerate all 256 corner patterns as 2x2x2 fields
    lut = np.array([0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0])
    for code in range(256):
        f = np.array([(code >> n) & 1 for n in range(8)], dtype=np.float64).reshape(2, 2, 2)
        codes, counts = oracle.case_codes(f, [np.arange(2.0)] * 3, 0.5)
        total = sum(lut[(int(codes[0, 0, 0]) >> (4 * n)) & 15] for n in range(6))
        assert total == int(counts[0, 0, 0])
    src = open(os.path.join(ROOT, "contour3d_b200", "csrc", "c3d_tables.h")).read()
    assert "c3d_tables_build" in src


def test_minimpi_fork_mode_bcast_and_reduce(tmp_path):
    """compat/minimpi: 3 forked ranks agree on a broadcast and a max-allreduce"""
    src = tmp_path / "t.c"
    src.write_text(
        '#include <stdio.h>\n#include <mpi.h>\n'
        "int main(void){MPI_Init(NULL,NULL);int r,n;MPI_Comm_rank(MPI_COMM_WORLD,&r);MPI_Comm_size(MPI_COMM_WORLD,&n);"
        "int v=r==0?42:0;MPI_Bcast(&v,1,MPI_INT,0,MPI_COMM_WORLD);double m=r;MPI_Allreduce(MPI_IN_PLACE,&m,1,MPI_DOUBLE,MPI_MAX,MPI_COMM_WORLD);"
        'if(v!=42||m!=n-1)return 1;if(r==0)printf("ok %d\\n",n);MPI_Finalize();return 0;}\n'
    )
    exe = tmp_path / "t"
    mm = os.path.join(ROOT, "compat", "minimpi")
    subprocess.run(["gcc", "-std=gnu11", "-I", mm, str(src), os.path.join(mm, "minimpi.c"), "-o", str(exe), "-lpthread", "-lrt"], check=True)
    env = dict(os.environ, MINIMPI_NP="3")
    for k in ("RANK", "WORLD_SIZE"):
        env.pop(k, None)
    out = subprocess.run([str(exe)], env=env, capture_output=True, text=True, check=True).stdout
    assert out.strip() == "ok 3"


def test_converter_probe_finds_the_dependence_structure():
    """c3d_positions_probe (host only): which output component depends on which direction, tables bit-equal to the function"""
    import math

    g = [np.linspace(-0.5, 0.5, 14), np.linspace(-0.5, 0.5, 11), np.linspace(-0.5, 0.5, 9)]
    off, cnt = (2, 1, 0), (10, 9, 9)

    def check(fn, want_strides):
        ident, tabs, st = api.positions_probe(fn, g, off, cnt)
        assert not ident
        assert st == want_strides, st
        rng = np.random.default_rng(0)
        for _ in range(200):
            i, j, k = (int(rng.integers(0, cnt[d])) for d in range(3))
            want = fn(g[0][off[0] + i], g[1][off[1] + j], g[2][off[2] + k])
            got = tuple(tabs[c][i * st[c][0] + j * st[c][1] + k * st[c][2]] for c in range(3))
            assert got == tuple(float(v) for v in want)

    assert api.positions_probe(lambda x, y, z: (x, y, z), g, off, cnt)[0]
    check(lambda x, y, z: (1.5 * x, y - 0.2 * y * y, z + 1.0), [[1, 0, 0], [0, 1, 0], [0, 0, 1]])                     # separable
    check(lambda r, t, z: ((0.75 + r) * math.cos(t), (0.75 + r) * math.sin(t), z), [[1, 10, 0], [1, 10, 0], [0, 0, 1]])  # cylindrical
    check(lambda x, y, z: (x * math.cos(z) - y * math.sin(z), x * math.sin(z) + y * math.cos(z), z + 0.1 * x * y),
          [[1, 10, 90], [1, 10, 90], [1, 10, 90]])                                                                    # twisted
    check(lambda x, y, z: (0.25, y, z + x), [[0, 0, 0], [0, 1, 0], [1, 0, 10]])                                       # constant component


def test_converter_probe_respects_the_memory_bound(capfd):
    g = [np.linspace(0, 1, 20)] * 3
    with pytest.raises(RuntimeError):
        api.positions_probe(lambda x, y, z: (x + y * z, y, z), g, (0, 0, 0), (20, 20, 20), max_bytes=1000)
    assert "CONTOUR3D_CONVERTER_MAX_BYTES" in capfd.readouterr().err


def test_camera_helper_reproduces_the_demo_camera():
    """c3d_camera_from_angles follows CameraAdjuster/src/config.ts:140-192; elevation 60, azimuth 30, d = 2 is the camera
    of the reference's demo (src/main.c: position (0.8660254, -1.5, 1.0), screen centre = position / 2)"""
    c = api.camera_from_angles(60.0, 30.0, 0.0, camera_distance=2.0, screen_distance=1.0, screen_width=4.0 / 3.0, screen_height=1.0)
    assert np.allclose(c["camera"], (0.8660254037844386, -1.5, 1.0), atol=1e-12)
    assert np.allclose(c["center"], np.array(c["camera"]) / 2, atol=1e-15)
    lx, ly, cam = (np.array(c[k]) for k in ("local_x", "local_y", "camera"))
    assert abs(lx @ ly) < 1e-12 and abs(np.linalg.norm(lx) - 4.0 / 3.0) < 1e-12 and abs(np.linalg.norm(ly) - 1.0) < 1e-12
    normal = np.cross(lx, ly) / np.linalg.norm(np.cross(lx, ly))
    assert np.allclose(normal, cam / np.linalg.norm(cam), atol=1e-12)          # the screen faces the camera
    assert abs(lx[2]) < 1e-12                                                  # no roll: the horizontal vector stays horizontal
    rolled = api.camera_from_angles(60.0, 30.0, 90.0, 2.0, 1.0, 4.0 / 3.0, 1.0)
    assert np.allclose(np.array(rolled["local_x"]) / (4.0 / 3.0), ly, atol=1e-12)
    with pytest.raises(ValueError):
        api.camera_from_angles(10.0, 10.0, camera_distance=0.0)


def test_converter_probe_rejects_a_function_that_is_not_one(capfd):
    """a callback that returns different values for the same argument cannot be tabulated (nor rendered reproducibly)"""
    import itertools

    counter = itertools.count()
    g = [np.linspace(0, 1, 6)] * 3
    with pytest.raises(RuntimeError):
        api.positions_probe(lambda x, y, z: (x + 1e-3 * next(counter), y, z), g, (0, 0, 0), (6, 6, 6))
    assert "does not return the same value" in capfd.readouterr().err


def test_converter_probe_keeps_nan_and_signed_zero_bit_patterns():
    g = [np.array([-1.0, 0.0, 1.0]), np.array([0.0, 1.0]), np.array([2.0, 3.0])]
    fn = lambda x, y, z: (float("nan") if x == 0.0 else -0.0 * x, y, z)      # noqa: E731
    ident, tabs, st = api.positions_probe(fn, g, (0, 0, 0), (3, 2, 2))
    assert not ident and st[0] == [1, 0, 0]
    assert np.isnan(tabs[0][1]) and np.signbit(tabs[0][2]) and not np.signbit(tabs[0][0])


def test_casefile_carries_the_converter_kind(tmp_path):
    f = scenes.gyroid(6)
    g = scenes.uniform_grids(6, 6, 6)
    sc = scenes.scene(32, 24)
    sc["converter"] = 3
    path = str(tmp_path / "c.bin")
    casefile.write_case(path, sc, g, [f], [dict(field=0, threshold=0.1, color=(1, 2, 3))], [])
    scene, *_ = casefile.read_case(path)
    assert scene["converter"] == 3


@pytest.mark.parametrize("ranks,dims", [(1, (0, 0, 0)), (2, (0, 0, 0)), (4, (0, 0, 0)), (6, (1, 3, 2)), (8, (0, 0, 0)), (8, (1, 1, 8)), (5, (1, 5, 1))])
def test_sdecomp_standin_agrees_with_simpledecomp(tmp_path, ranks, dims):
    """The SimpleDecomp stand-in of MPI-less hosts (c3d_sdecomp_compat.c) against the genuine SimpleDecomp (compiled into
    oracle/_ref/libc3dref.so): process-grid sizes, ranks, neighbours, pencil sizes and offsets for all six pencils and
    three directions, on every rank, under the fork-mode MPI stand-in.  Periodic in x and z, not in y."""
    import subprocess

    from oracle import reference

    if not reference.available():
        pytest.skip("oracle/_ref not built")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tests", "csrc", "sdecomp_probe.c")
    pkg = os.path.join(root, "contour3d_b200")
    builds = {
        "standin": ["-I", os.path.join(root, "include"), "-L", pkg, "-lc3d_sdecomp_compat", f"-Wl,-rpath,{pkg}"],
        # the genuine library is declared by the stand-in header too (link-compatible by construction: same method-table layout)
        "genuine": ["-I", os.path.join(root, "include"), "-L", reference.REF_DIR, "-l:libc3dref.so", f"-Wl,-rpath,{reference.REF_DIR}"],
    }
    out = {}
    for name, flags in builds.items():
        exe = str(tmp_path / f"probe_{name}")
        subprocess.run(["gcc", "-std=gnu11", "-O1", "-I", os.path.join(root, "compat", "minimpi"), "-o", exe, src] + flags, check=True)
        env = dict(os.environ, MINIMPI_NP=str(ranks))
        prefix = str(tmp_path / f"out_{name}")
        subprocess.run([exe] + [str(v) for v in dims] + ["37", "50", "41", prefix], check=True, env=env, timeout=60)
        out[name] = [open(f"{prefix}.{r}").read() for r in range(ranks)]
    assert out["standin"] == out["genuine"]
    assert "nprocs" in out["genuine"][0]


def test_host_generator_matches_the_numpy_construction():
    """libc3d_synth_host.so (the reference arm's field source) against the numpy statement of the same noise (equal where
    the coordinates are exact in both, i.e. power-of-two sizes; bit-equality with the DEVICE generator is a GPU test)"""
    a = scenes.synth_host((32, 32, 32), kind=1, seed=99, octaves=4, gain=0.5)
    b = scenes.noise(32, seed=99, octaves=4, gain=0.5)
    assert a.dtype == np.float32 and a.shape == (32, 32, 32)
    assert np.array_equal(a, b)
    c = scenes.synth_host((24, 24, 24), kind=1, seed=99, octaves=4, gain=0.5)
    assert np.abs(c - scenes.noise(24, seed=99, octaves=4, gain=0.5)).max() < 1e-6
    v = scenes.synth_host((24, 24, 24), kind=2, seed=2)
    assert np.isfinite(v).all() and v.min() >= -1.0 and v.std() > 0.05


def test_bench_workloads_follow_baseline_configs():
    import argparse
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    want = {2: (512, (1920, 1080), 1, 1), 3: (1024, (3840, 2160), 1, 3), 4: (1024, (3840, 2160), 2, 4), 5: (2048, (7680, 4320), 1, 4)}
    for c, (size, image, nfields, ncontours) in want.items():
        w = bench.workload(argparse.Namespace(config=c, size=0, image=None, isovalues=None))
        assert (w["size"], w["image"], len(w["fields"]), len(w["contours"])) == (size, image, nfields, ncontours)
        assert all(v == float(np.float32(v)) for _, v in w["contours"])


def test_converter_probe_catches_a_function_that_is_the_identity_almost_everywhere():
    """the identity except inside a box away from the domain's corners and centre lines: whichever of the probe's stages
    notices (sample lines, the scattered check of the identity verdict, the verification of the tables), the tables it
    returns must reproduce the function inside and outside the box"""
    n = 40
    g = scenes.uniform_grids(n, n, n)
    ylo, yhi, zlo, zhi = g[1][3], g[1][12], g[2][3], g[2][12]

    def conv(x, y, z):
        inside = ylo <= y <= yhi and zlo <= z <= zhi
        return (x + (0.25 if inside else 0.0), y, z)

    identity, tables, strides = api.positions_probe(conv, g, (0, 0, 0), (n, n, n))
    assert not identity

    def lookup(c, i, j, k):
        return tables[c][i * strides[c][0] + j * strides[c][1] + k * strides[c][2]]

    rng = np.random.default_rng(1)
    for i, j, k in rng.integers(0, n, size=(400, 3)):
        want = conv(g[0][i], g[1][j], g[2][k])
        assert all(lookup(c, i, j, k) == want[c] for c in range(3)), (i, j, k)


def _work_plans(counts, remote_cost):
    """c3d_selftest_work_plan for every rank: [(chunks as rows (owner, d, first, count, dest), nlocal, total)]"""
    fn = api.lib().c3d_selftest_work_plan
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
    n = counts.shape[0]
    flat = np.ascontiguousarray(counts, dtype=np.uint32)
    out = []
    for me in range(n):
        chunks = np.zeros((8 * 10, 5), dtype=np.uint32)
        nch, nlocal, total = C.c_uint(0), C.c_uint(0), C.c_uint(0)
        assert 0 == fn(flat.ctypes.data, n, me, float(remote_cost), chunks.ctypes.data, C.byref(nch), C.byref(nlocal), C.byref(total))
        out.append((chunks[:nch.value].astype(np.int64), int(nlocal.value), int(total.value)))
    return out


@pytest.mark.parametrize("n,remote_cost,seed", [(2, 8.0, 0), (4, 8.0, 1), (8, 8.0, 2), (8, 3.0, 3), (8, 1.0, 4), (8, 0.0, 5), (3, 5.0, 6), (8, 8.0, 7)])
def test_shading_work_plan_hands_out_every_pixel_exactly_once(n, remote_cost, seed):
    """k_work_plan (c3d_peer.cu), run on the host: the ranks of a screen-partitioned frame decide, each for itself and from
    the same n x n counts, which pixels they shade.  Whatever the counts: every queued pixel is taken by exactly one rank,
    a rank's pieces tile its work list without gaps, its own array's pixels come first, nobody ends above the busiest owner,
    and with helping switched off everybody takes exactly its own."""
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, 50000, size=(n, n))           # counts[d][o]: pixels of rank d's tiles won by cubes of rank o's array
    counts[:, n - 1] *= 6 if seed != 7 else 0               # the slab facing the camera wins most of the image (or nothing at all)
    if seed == 7:
        counts[rng.integers(0, n)] = 0
    plans = _work_plans(counts, remote_cost)
    taken = {(d, o): np.zeros(counts[d, o], dtype=np.int32) for d in range(n) for o in range(n)}
    owner_total = counts.sum(axis=0)
    for me, (chunks, nlocal, total) in enumerate(plans):
        at = 0
        for owner, d, first, count, dest in chunks:
            assert dest == at and count > 0, "pieces must tile the work list in order"
            assert first + count <= counts[d, owner]
            taken[(d, owner)][first:first + count] += 1
            if at < nlocal:
                assert owner == me, "a rank's own pixels come first"
            else:
                assert owner != me
            at += count
        assert at == total
        assert total <= max(int(owner_total.max()), 1)
        if remote_cost < 1.0:
            assert total == nlocal == owner_total[me]
    for key, hits in taken.items():
        assert (hits == 1).all(), f"pixels of segment {key} taken {np.unique(hits)} times"
    if remote_cost >= 1.0 and owner_total.max() > 0:
        # the estimated time is levelled: an owner above the level keeps the level, a helper's foreign pixels count remote_cost times
        est = [nl + remote_cost * (tot - nl) for _, nl, tot in plans]
        assert max(est) <= owner_total.max() + 1
        busiest = int(np.argmax(owner_total))
        others = [e for r, e in enumerate(est) if r != busiest]
        if owner_total[busiest] > 2 * np.mean(owner_total) and n > 1:
            assert plans[busiest][2] < owner_total[busiest], "the busiest owner hands part of its pixels over"
            assert max(others) <= est[busiest] + remote_cost + 1


@pytest.mark.parametrize("n,tiles,shares", [
    (8, (60, 34), [0.125] * 8),
    (8, (120, 68), [0.135, 0.134, 0.135, 0.135, 0.131, 0.130, 0.127, 0.073]),     # the shares measured at 8 GPUs
    (2, (5, 3), [0.7, 0.3]),
    (3, (7, 11), [0.2, 0.5, 0.3]),
])
def test_tile_table_follows_the_shares_and_scatters_every_ranks_tiles(n, tiles, shares):
    """tile_map_build (c3d_runtime.cu): the screen tiles of a partitioned frame are dealt out in proportion to the ranks'
    shares -- every tile to exactly one rank, counts within one tile of share x tiles, and no rank's tiles bunched together
    (every band of rows holds about its share of them)"""
    fn = api.lib().c3d_selftest_tile_map
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    tx, ty = tiles
    share = np.asarray(shares, dtype=np.float64)
    share = share / share.sum()
    out = np.full(tx * ty, 255, dtype=np.uint8)
    assert 0 == fn(share.ctypes.data, n, tx, ty, out.ctypes.data)
    assert out.max() < n
    got = np.bincount(out, minlength=n)
    assert np.abs(got - share * tx * ty).max() <= 1.0 + 1e-9, (got, share * tx * ty)
    if tx * ty >= 1000:
        grid = out.reshape(ty, tx)
        bands = np.array_split(np.arange(ty), 4)
        for r in range(n):
            per_band = np.array([(grid[b] == r).sum() / (len(b) * tx) for b in bands])
            assert np.abs(per_band - share[r]).max() < 0.25 * share[r] + 0.01, (r, per_band, share[r])
