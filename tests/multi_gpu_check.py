"""N-rank parity check, one process per rank (launch with torch.distributed.run); ranks take GPU rank % device_count,
so on a box with fewer GPUs than ranks several ranks share one (the library then exchanges through peer memory only):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multi_gpu_check.py [dims]

Every rank uploads ITS pencil of a host field, the library exchanges halos over NCCL,
renders, reduces; rank 0 compares the image with the oracle's single-rank image of the
whole field (the reference is decomposition-invariant) bit for bit, and the
extended arrays with the slices the reference's extension would hold
(contour/extend.c:216-278).  dims: "1x1xN" (z slabs, default) or e.g. "1x2x2" (pencils).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from contour3d_b200 import api, scenes


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    dims = tuple(int(v) for v in sys.argv[1].split("x")) if len(sys.argv) > 1 else (1, 1, world)
    assert dims[0] * dims[1] * dims[2] == world
    local = local % max(1, torch.cuda.device_count())
    torch.cuda.set_device(local)
    # the test's own plumbing (id broadcast, verdict) goes over gloo: it must also work when ranks share a GPU
    dist.init_process_group("gloo")
    t = torch.zeros(api.NCCL_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        t = torch.frombuffer(bytearray(api.nccl_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(t, src=0)
    ctx = api.Context(device=local, rank=rank, size=world, nccl_id=bytes(t.numpy().tobytes()), dims=dims)
    mode = api.lib().c3d_context_exchange_mode(ctx.handle)
    want_mode = os.environ.get("C3D_EXPECT_MODE")
    assert want_mode is None or int(want_mode) == mode, f"exchange mode {mode}, expected {want_mode}"
    failures = []
    cases = [(np.float32, 44, (0.1, -0.2)), (np.float64, 38, (0.05,))]
    if torch.cuda.device_count() >= world:
        cases.append((np.float64, 37, (0.05,)))       # rows that are not 16-byte multiples: direct sweep + the NCCL path (one GPU per rank)
    for dtype, n, thr in cases:
        full = scenes.noise(n, seed=7).astype(dtype) if dtype == np.float32 else scenes.sines(n, seed=3)[0].astype(dtype)
        grids = scenes.uniform_grids(n, n, n)
        f = ctx.field((n, n, n), dtype=dtype).upload_global(full).exchange()
        modes = set()
        # 1. halo contents: the extended array is the window of the global field
        ext = f.download_ext()
        o, e = f.ext_off, f.ext
        want = full[o[2]:o[2] + e[2], o[1]:o[1] + e[1], o[0]:o[0] + e[0]]
        if not np.array_equal(ext, want):
            failures.append(f"rank {rank}: halo of the {np.dtype(dtype).name} field differs")
        # 2. the frame
        scene = scenes.scene(320, 180)
        lines = scenes.box_lines(nitems=12, width=2.0)
        cols = ((255, 200, 40), (40, 200, 255))
        cs = [dict(field=f, grids=grids, threshold=float(dtype(v)), color=c) for v, c in zip(thr, cols)]
        img = ctx.render(scene, cs, lines)
        modes.add(ctx.timings()["exchange_mode"])
        # again with stale halos: the exchange then runs inside the frame (contour/extend.c:211-213 does it for every object of every frame)
        f.upload_global(full)
        img2 = ctx.render(scene, cs, lines)
        if rank == 0 and not np.array_equal(img, img2):
            failures.append(f"{np.dtype(dtype).name}: the frame with the halo exchange inside differs")
        if rank == 0:
            from oracle import oracle

            ocs = [dict(field=full, grids=grids, threshold=float(dtype(v)), color=c) for v, c in zip(thr, cols)]
            rgb, _, _ = oracle.frame(scene, ocs, lines, nthreads=4)
            differ = int((img != rgb[::-1]).any(axis=2).sum())
            if differ:
                failures.append(f"{np.dtype(dtype).name}: {differ} pixels differ from the single-rank oracle image")
        f.destroy()
    flag = torch.tensor([len(failures)])
    dist.all_reduce(flag)
    for msg in failures:
        print("FAIL", msg, flush=True)
    if rank == 0:
        print(f"multi_gpu_check dims={dims} exchange_mode={mode} gpus={torch.cuda.device_count()}: {'ok' if flag.item() == 0 else 'FAILED'}", flush=True)
    ctx.destroy()
    dist.destroy_process_group()
    return 0 if flag.item() == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
