"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 path.

The compute of the product is CUDA-only, so what can be exercised here is the
plumbing the multi-GPU path relies on: id broadcast through torch.distributed,
the block decomposition / halo windows every rank derives, and the two-stage
image reduction algebra (max over order-preserving depth keys, then max over
the RGBA words of the ranks that own the nearest depth) -- run with the oracle
as the per-rank renderer and compared with the single-rank image."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ord(depth):
    b = depth.view(np.uint64)
    neg = (b >> np.uint64(63)).astype(bool)
    return np.where(neg, ~b, b | np.uint64(1 << 63))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from contour3d_b200 import api, scenes
    from oracle import oracle

    # 1. bootstrap plumbing: rank 0's 128-byte id reaches everybody
    ident = torch.zeros(api.NCCL_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        ident = torch.arange(api.NCCL_ID_BYTES, dtype=torch.uint8)
    dist.broadcast(ident, src=0)
    assert ident[77].item() == 77

    # 2. decomposition: z slabs {1, 1, world}; windows tile the domain with a 2-point halo per interior side
    n = 20
    d = api.make_decomp((1, 1, world), rank)
    my, off = api.decomp_local(d, (n, n, n))
    lo = 2 if rank > 0 else 0
    hi = 2 if rank < world - 1 else 0
    sizes = torch.tensor([my[2]])
    dist.all_reduce(sizes)
    assert sizes.item() == n
    f, g = scenes.sines(n, seed=4)
    window = f[off[2] - lo: off[2] + my[2] + hi]
    gz = g[2][off[2] - lo: off[2] + my[2] + hi]

    # 3. per-rank render + two-stage reduction, as c3d_frame_render does it on the GPUs
    scene = scenes.scene(120, 90)
    lines = scenes.box_lines(nitems=8, width=2.0)
    cs = [dict(field=window, grids=[g[0], g[1], gz], threshold=0.15, color=(200, 100, 50))]
    rgb, depth, _ = oracle.frame(scene, cs, lines)
    key = torch.from_numpy(_ord(depth).astype(np.int64) ^ np.int64(-2 ** 63))  # order-preserving signed view for gloo
    best = key.clone()
    dist.all_reduce(best, op=dist.ReduceOp.MAX)
    covered = depth > -1e300
    mine = (key == best).numpy() & covered
    rgba = np.zeros(depth.shape, dtype=np.int64)
    packed = (255 << 24) | (rgb[..., 2].astype(np.int64) << 16) | (rgb[..., 1].astype(np.int64) << 8) | rgb[..., 0]
    rgba[mine] = packed[mine]
    t = torch.from_numpy(rgba)
    dist.reduce(t, dst=0, op=dist.ReduceOp.MAX)
    if rank == 0:
        out = t.numpy()
        img = np.stack([out & 255, (out >> 8) & 255, (out >> 16) & 255], axis=-1).astype(np.uint8)
        img[(out >> 24) == 0] = np.array(scene["bg"], dtype=np.uint8)
        ocs = [dict(field=f, grids=g, threshold=0.15, color=(200, 100, 50))]
        want, _, _ = oracle.frame(scene, ocs, lines)
        assert np.array_equal(img, want), "2-rank composite differs from the single-rank image"
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_two_gloo(tmp_path):
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def _board_worker(rank, world, ident, out):
    sys.path.insert(0, ROOT)
    import ctypes as C

    from contour3d_b200 import api

    fn = api.lib().c3d_peer_selftest
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_ulonglong, C.POINTER(C.c_ulonglong)]
    total = C.c_ulonglong(0)
    rc = fn(rank, world, ident, 7, C.byref(total))
    out.put((rank, rc, int(total.value)))


@pytest.mark.parametrize("world", [2, 5])
def test_ranks_bootstrap_board_allgathers_between_processes(world):
    """the shared-memory board over which the ranks of one node exchange the CUDA IPC handles of each other's device memory
    (c3d_peer.cu): independent processes that share only the 128-byte id of the run meet on it, every one contributes a
    value and all receive the same sum -- host logic of the N > 1 path, no device involved"""
    ctx = mp.get_context("spawn")
    ident = os.urandom(128)
    out = ctx.Queue()
    procs = [ctx.Process(target=_board_worker, args=(r, world, ident, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    want = 7 * world * (world + 1) // 2
    assert got == [(r, 0, want) for r in range(world)], got
