#!/usr/bin/env python
"""Benchmark of the per-frame pipeline (BASELINE.json: Gvoxel/s and frames/s at
1024^3, 3 isovalues, 3840x2160, slab-sharded over N B200).

    python bench.py --gpus N --steps K --warmup W                 (ours)
    python bench.py --impl reference --gpus N --steps K --warmup W (the reference's CPU path)

One "step" = one frame: all contour objects + lines + cross-rank reduction +
RGB packing, on synthetic data (device-generated multi-octave value noise).

  value   whole-job Gvoxel/s = global grid points x contour objects / frame time,
          field resident in HBM (fp32, 4 B/voxel), image left on the device
  e2e     the same frame through the reference-facing call contour3d_execute()
          with HOST double arrays (8 B/voxel, the public API's type) from pinned
          memory and the PPM written to /dev/shm, copies inside the timed region
  roofline  the classification sweep (k_sweep): algorithmic bytes = 4 B x points
          of the extended local array per launch, time from CUDA events recorded
          on the library's stream around each launch, averaged over the timed steps
  cpu_baseline  the unmodified reference (oracle/_ref/ref_driver, ranks = host
          cores, y-split) on a bounded crop of the same field
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ISOVALUES = (-0.3, 0.05, 0.4)
COLORS = ((255, 96, 64), (96, 255, 96), (64, 128, 255))
NOISE = dict(kind=1, seed=1234, octaves=5, gain=0.5)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.lines, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def pump():
                for line in self.proc.stdout:
                    self.lines.append((time.time(), line))

            self.thread = threading.Thread(target=pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        inside = [l for (ts, l) in self.lines if self.t0 is not None and self.t1 is not None and self.t0 - 0.02 <= ts <= self.t1 + 0.05]
        window = "timed region" if inside else "warm-up + timed region"
        for line in (inside or [l for (_, l) in self.lines]):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm), "window": window}


def sample_case(field_crop: np.ndarray, size: int, width: int, height: int, path: str):
    """case file of the bounded CPU sample: a centred crop of the benchmark field, same isovalues / camera / image"""
    from contour3d_b200 import casefile, scenes

    nz, ny, nx = field_crop.shape
    # keep the crop at its place in the global domain: coordinates of the crop's points
    g = scenes.uniform_grids(size, size, size)
    o = [(size - n) // 2 for n in (nx, ny, nz)]
    grids = [g[0][o[0]:o[0] + nx], g[1][o[1]:o[1] + ny], g[2][o[2]:o[2] + nz]]
    contours = [dict(field=0, threshold=float(np.float32(t)), color=c) for t, c in zip(ISOVALUES, COLORS)]
    casefile.write_case(path, scenes.scene(width, height), grids, [field_crop], contours, scenes.box_lines(nitems=128, width=4.0))


def run_reference_on(case: str, steps: int, warmup: int, cores: int):
    from oracle import reference

    ranks = max(1, cores)
    out = os.path.join(os.path.dirname(case), "ref.ppm")
    res = reference.run_case(case, out, ranks=ranks, dims=(1, ranks, 1), repeat=steps, warmup=warmup)
    return res


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_size(cores):
    # ~1 Mvoxel.contour/s per core: aim for 10-30 s of CPU work for one frame of 3 contours
    target = 12e6 * max(1, cores) / 3.0
    n = int(round(target ** (1.0 / 3.0) / 16.0)) * 16
    return int(min(384, max(96, n)))


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import reference

    if not reference.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built (no /root/reference at build time)"}))
        return 0
    import torch
    from contour3d_b200 import api

    cores = host_cores()
    n = cpu_sample_size(cores)
    size, (width, height) = args.size, args.image
    # the sample is a crop of the SAME device-generated field the GPU arm renders
    ctx = api.Context(device=0)
    crop = device_crop(ctx, size, n)
    ctx.destroy()
    with tempfile.TemporaryDirectory(dir="/dev/shm") as tmp:
        case = os.path.join(tmp, "sample.bin")
        sample_case(crop, size, width, height, case)
        ranks = min(cores, n // 4)
        res = run_reference_on(case, args.steps, args.warmup, ranks)
    work = n ** 3 * len(ISOVALUES)
    value = work / res["seconds_mean"] / 1e9
    line = {
        "impl": "reference", "metric": "Gvoxel/s", "value": value, "unit": "Gvoxel/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["seconds_mean"] * 1e3, "frames_per_s": 1.0 / res["seconds_mean"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config 3 of BASELINE.json: multi-octave noise {size}^3, 3 isovalues, {width}x{height}",
                   "sample": f"centred {n}^3 crop of the field, full image", "timing": "wall clock of contour3d_execute, max over ranks"},
        "cpu_baseline": {"value": value, "unit": "Gvoxel/s", "cores": ranks, "kind": "reference",
                         "sample": f"unmodified reference, {ranks} ranks (y split), centred {n}^3 crop x 3 isovalues, {width}x{height}, mean of {args.steps} frames"},
        "e2e": {"value": value, "unit": "Gvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def device_crop(ctx, size, n, field=None):
    """centred n^3 crop of the size^3 benchmark field: generated on the device (the ONE generator), copied back"""
    own = field is None
    if own:
        field = ctx.field((size, size, size), dtype=np.float32).synthesize(**NOISE)
    full = field.download_ext()
    if own:
        field.destroy()
    o = (size - n) // 2
    return np.ascontiguousarray(full[o:o + n, o:o + n, o:o + n])


def ours(args):
    import torch
    import torch.distributed as dist
    from contour3d_b200 import api, scenes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    ident = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        t = torch.zeros(api.NCCL_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t = torch.frombuffer(bytearray(api.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(t, src=0)
        ident = bytes(t.cpu().numpy().tobytes())
    size, (width, height) = args.size, args.image
    ctx = api.Context(device=local, rank=rank, size=world, nccl_id=ident, dims=(1, 1, world))
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    field = ctx.field((size, size, size), dtype=np.float32).synthesize(**NOISE)
    grids = scenes.uniform_grids(size, size, size)
    scene = scenes.scene(width, height)
    lines = scenes.box_lines(nitems=128, width=4.0)
    contours = [dict(field=field, grids=grids, threshold=float(np.float32(t)), color=c) for t, c in zip(ISOVALUES, COLORS)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    # ---- device-resident arm ----
    stage_sum, launches = {}, 0
    sweep_ms, sweep_bytes = [], []
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:          # nvidia-smi needs ~0.2 s to start: it runs through the warm-up too
        for _ in range(args.warmup):
            ctx.render(scene, contours, lines, want_rgb=False)
        barrier()
        clocks.mark_start()
        start.record(stream)
        for _ in range(args.steps):
            ctx.render(scene, contours, lines, want_rgb=False)
            t = ctx.timings()
            launches += t["launches"]
            sweep_ms += t["sweep_each"]
            sweep_bytes += t["sweep_bytes"]
            for k in ("halo", "clear", "sweep", "expand", "raster", "lines", "resolve", "reduce", "shade", "pack", "download", "total"):
                stage_sum[k] = stage_sum.get(k, 0.0) + t[k]
        end.record(stream)
        barrier()
        clocks.mark_end()
        time.sleep(0.05)
    elapsed_ms = start.elapsed_time(end)
    tt = torch.tensor([elapsed_ms], device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    elapsed_ms = float(tt.item())
    active = t["active_cubes"]
    # where each rank spent the frame (the front slab of this camera owns most of the visible surface)
    per_rank = None
    if world > 1:
        mine = {k: round(v / args.steps, 4) for k, v in stage_sum.items()}
        mine["active_cubes"] = int(sum(active[:len(ISOVALUES)]))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        per_rank = gathered
    voxels = size ** 3 * len(ISOVALUES)
    value = voxels * args.steps / (elapsed_ms * 1e-3) / 1e9

    # ---- the exchange steps on their own (N > 1): NVLink / NVSwitch throughput of the halo and of the image reduction ----
    collectives = None
    if world > 1:
        iters = 10
        ar_ms, rd_ms = ctx.collectives_probe(width, height, iters)
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        field.exchange()
        h0.record(stream)
        for _ in range(iters):
            field.exchange()
        h1.record(stream)
        barrier()
        halo_ms = h0.elapsed_time(h1) / iters
        faces = (1 if rank > 0 else 0) + (1 if rank < world - 1 else 0)           # z slabs: interior faces of this rank
        halo_bytes = 2 * faces * 2 * field.ext[0] * field.ext[1] * 4             # 2 planes out + 2 planes in per face, fp32
        npix = width * height
        mine = {"depth_allreduce_ms": ar_ms, "rgba_reduce_ms": rd_ms, "halo_ms": halo_ms, "halo_bytes": halo_bytes}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        ar = max(g["depth_allreduce_ms"] for g in gathered)
        rd = max(g["rgba_reduce_ms"] for g in gathered)
        hm = max(g["halo_ms"] for g in gathered)
        hb = max(g["halo_bytes"] for g in gathered)
        collectives = {
            "depth_allreduce": {"bytes": 8 * npix, "ms": ar, "algbw_GBps": 8 * npix / (ar * 1e-3) / 1e9,
                                "busbw_GBps": 2.0 * (world - 1) / world * 8 * npix / (ar * 1e-3) / 1e9},
            "rgba_reduce": {"bytes": 4 * npix, "ms": rd, "algbw_GBps": 4 * npix / (rd * 1e-3) / 1e9},
            "halo_exchange": {"bytes_per_interior_rank": hb, "ms": hm, "GBps_per_rank": hb / (hm * 1e-3) / 1e9},
            "note": "ncclAllReduce(max, uint64) / ncclReduce(max, uint32) / ncclSend+Recv of 2 planes per face, timed alone with CUDA events, max over ranks",
        }

    # ---- e2e arm: contour3d_execute with host double arrays, PPM to /dev/shm ----
    e2e = None
    if not args.no_e2e:
        e2e = e2e_arm(args, ctx, field, grids, scene, lines, rank, world, barrier)

    # ---- roofline of the sweep ----
    peak, peak_src = peaks()
    sweep_avg_ms = float(np.mean(sweep_ms))
    achieved = float(np.mean(sweep_bytes)) / (sweep_avg_ms * 1e-3) / 1e9
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(tfile):
        traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
    roofline = {"kernel": f"k_sweep_tiles<float,{t['sweep_thresholds'][0]},...> (c3d_sweep.cu; one pass classifies {t['sweep_thresholds'][0]} iso-values)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "bytes_per_launch": float(np.mean(sweep_bytes)), "ms_per_launch": sweep_avg_ms,
                "share_of_step": stage_sum["sweep"] / stage_sum["total"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, ctx, field, size, width, height)

    if rank == 0:
        line = {
            "metric": "Gvoxel/s", "value": value, "unit": "Gvoxel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "frames_per_s": args.steps / (elapsed_ms * 1e-3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 field / f64 geometry", "data": "synthetic",
            "config": {"workload": f"config 3 of BASELINE.json: multi-octave noise {size}^3 fp32 (device generated), 3 isovalues "
                                   f"{ISOVALUES}, {width}x{height}, 8 box lines, z slabs dims={{1,1,{world}}}",
                       "l2": "inputs larger than L2 (4.3 GB field vs 126 MB)", "timing": "CUDA events on the library stream, max over ranks"},
            "clocks": clocks.summary(),
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "stages_ms": {k: v / args.steps for k, v in stage_sum.items()},
            "stages_ms_per_rank": per_rank,
            "collectives": collectives,
            "active_cubes": active, "culled_cubes": t["culled_cubes"], "culled_triangles": t["culled_triangles"], "shade_spills": t["shade_spills"],
        }
        print(json.dumps(line))
    barrier()
    field.destroy()
    ctx.destroy()
    if world > 1:
        dist.destroy_process_group()
    return 0


def e2e_arm(args, ctx, field, grids, scene, lines, rank, world, barrier):
    """contour3d_execute (the public drop-in) with HOST double arrays from pinned memory"""
    import torch
    from contour3d_b200 import api

    size = args.size
    # host copy of my pencil, widened exactly to double (what a solver would hold), in pinned memory
    ext = field.download_ext()
    lo = [field.offsets[d] - field.ext_off[d] for d in range(3)]
    m = field.mysizes
    local = ext[lo[2]:lo[2] + m[2], lo[1]:lo[1] + m[1], lo[0]:lo[0] + m[0]]
    pinned = torch.empty(local.shape, dtype=torch.float64, pin_memory=True)
    pinned.numpy()[...] = local
    del ext, local
    host = pinned.numpy()
    info = api.sdecomp_construct((1, 1, world))
    contours = [dict(field=0, threshold=float(np.float32(t)), color=c) for t, c in zip(ISOVALUES, COLORS)]
    out = f"/dev/shm/contour3d_b200_bench_{os.getpid()}.ppm"
    steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(1):
        rc = api.execute(scene, grids, [host], contours, lines, out, sdecomp_info=info)
        assert rc == 0, "contour3d_execute failed"
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        rc = api.execute(scene, grids, [host], contours, lines, out, sdecomp_info=info)
        assert rc == 0
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    if rank == 0 and os.path.exists(out):
        os.unlink(out)
    voxels = size ** 3 * len(ISOVALUES)
    return {"value": voxels * steps / dt / 1e9, "unit": "Gvoxel/s", "frames_per_s": steps / dt, "ms_per_step": dt / steps * 1e3, "steps": steps,
            "h2d_bytes_per_step": int(host.nbytes + 3 * 8 * size), "d2h_bytes_per_step": int(3 * scene["width"] * scene["height"]) if rank == 0 else 0,
            "path": "contour3d_execute(host double arrays, pinned) -> PPM in /dev/shm; wall clock around the call, H2D of the array and D2H of the image inside"}


def cpu_baseline(args, ctx, field, size, width, height):
    from oracle import oracle, reference

    cores = host_cores()
    n = cpu_sample_size(cores)
    crop = device_crop(ctx, size, n, field)
    work = n ** 3 * len(ISOVALUES)
    with tempfile.TemporaryDirectory(dir="/dev/shm") as tmp:
        if reference.available():
            case = os.path.join(tmp, "sample.bin")
            sample_case(crop, size, width, height, case)
            ranks = min(cores, n // 4)
            res = run_reference_on(case, 1, 0, ranks)
            secs = res["seconds_mean"]
            return {"value": work / secs / 1e9, "unit": "Gvoxel/s", "cores": ranks, "kind": "reference", "seconds": secs,
                    "sample": f"unmodified reference (oracle/_ref), {ranks} ranks y-split, centred {n}^3 crop x 3 isovalues, {width}x{height}, 1 frame"}
    # the restatement, threads emulating ranks
    from contour3d_b200 import scenes
    g = scenes.uniform_grids(size, size, size)
    o = (size - n) // 2
    grids = [g[d][o:o + n] for d in range(3)]
    cs = [dict(field=crop, grids=grids, threshold=float(np.float32(t)), color=c) for t, c in zip(ISOVALUES, COLORS)]
    t0 = time.perf_counter()
    oracle.frame(scenes.scene(width, height), cs, scenes.box_lines(nitems=128, width=4.0), nthreads=cores)
    secs = time.perf_counter() - t0
    return {"value": work / secs / 1e9, "unit": "Gvoxel/s", "cores": cores, "kind": "port", "seconds": secs,
            "sample": f"oracle restatement, {cores} threads y-split, centred {n}^3 crop x 3 isovalues, {width}x{height}, 1 frame"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--image", type=lambda s: tuple(int(v) for v in s.lower().split("x")), default=(3840, 2160))
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    return reference_arm(args) if args.impl == "reference" else ours(args)


if __name__ == "__main__":
    sys.exit(main())
