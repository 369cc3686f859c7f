#!/usr/bin/env python
"""Benchmark of the per-frame pipeline (BASELINE.json: Gvoxel/s and frames/s at
1024^3, 3 isovalues, 3840x2160, slab-sharded over N B200).

    python bench.py --gpus N --steps K --warmup W [--config 2|3|4|5]        (ours)
    python bench.py --impl reference --gpus N --steps K --warmup W          (the reference's CPU path)

One "step" = one frame: halo exchange (N > 1) + all contour objects + lines +
cross-rank exchange + RGB packing, on synthetic data.

  value   whole-job Gvoxel/s = global grid points x contour objects / frame time,
          field resident in HBM (fp32, 4 B/voxel), image left on the device.  At N > 1
          the halos are marked stale before every frame, so the exchange is inside.
  e2e     the same frame through the reference-facing call contour3d_execute()
          with HOST double arrays (8 B/voxel, the public API's type) from pinned
          memory and the PPM written to /dev/shm, copies inside the timed region
  e2e_device  the same call with a DEVICE double array (what a GPU-resident solver passes)
  roofline  the classification sweep (k_sweep_tiles): algorithmic bytes = 4 B x points
          of the extended local array per launch, time from CUDA events recorded
          on the library's stream around each launch, averaged over the timed steps
  cpu_baseline  the unmodified reference (oracle/_ref/ref_driver, ranks = host
          cores, y-split) on a bounded crop of the same field, and `crop_parity`:
          that very crop rendered by the GPU path, compared pixel by pixel with the reference's image
  image_md5  md5 of the packed RGB of one extra untimed frame: must not depend on N

Configurations (BASELINE.json `configs`; 3 is the one the metric is quoted on and the default):
  2  gyroid 512^3, 1 isovalue, 1920x1080        4  two fields 1024^3 x 2 isovalues each, 3840x2160
  3  noise 1024^3, 3 isovalues, 3840x2160       5  turbulent-like 2048^3, 4 isovalues, 7680x4320
"""
from __future__ import annotations

import argparse
import hashlib
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

COLORS = ((255, 96, 64), (96, 255, 96), (64, 128, 255), (255, 224, 64))


def workload(args):
    """fields: generator parameters (kind 0 gyroid: numpy on the host, uploaded; kinds 1 / 2: device generator with a
    bit-identical host twin); contours: (field index, isovalue)"""
    c = args.config
    if c == 2:
        w = dict(size=512, image=(1920, 1080), fields=[dict(kind=0)], contours=[(0, 0.1)],
                 name="config 2 of BASELINE.json: gyroid 512^3 fp32, 1 isovalue")
    elif c == 3:
        w = dict(size=1024, image=(3840, 2160), fields=[dict(kind=1, seed=1234, octaves=5, gain=0.5)], contours=[(0, -0.3), (0, 0.05), (0, 0.4)],
                 name="config 3 of BASELINE.json: multi-octave noise 1024^3 fp32, 3 isovalues")
    elif c == 4:
        w = dict(size=1024, image=(3840, 2160), fields=[dict(kind=1, seed=1, octaves=5, gain=0.5), dict(kind=2, seed=2, octaves=5, gain=0.5)],
                 contours=[(0, -0.2), (0, 0.25), (1, 0.3), (1, 0.8)],
                 name="config 4 of BASELINE.json: temperature (noise) + vorticity magnitude (|curl| of a 3-component noise) 1024^3 fp32, 2 isovalues each")
    elif c == 5:
        w = dict(size=2048, image=(7680, 4320), fields=[dict(kind=1, seed=1234, octaves=7, gain=2.0 ** (-5.0 / 6.0))],
                 contours=[(0, -0.3), (0, -0.05), (0, 0.15), (0, 0.4)],
                 name="config 5 of BASELINE.json: turbulent-like noise (7 octaves, k^-5/6) 2048^3 fp32, 4 isovalues")
    else:
        raise SystemExit(f"unknown --config {c}")
    if args.size:
        w["size"] = args.size
    if args.image:
        w["image"] = args.image
    if args.isovalues:
        w["contours"] = [(0, v) for v in args.isovalues]
    w["contours"] = [(f, float(np.float32(v))) for f, v in w["contours"]]      # fp32-representable: fp32 and fp64 classification agree
    return w


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons during the timed region, through NVML in a thread of this process (a looping `nvidia-smi`
    takes the driver's locks for milliseconds per poll and delays the kernel launches of the rank it runs beside: measured,
    4 ms per frame on a 4-GPU box); `nvidia-smi` only where the NVML binding is missing."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc, self.thread = index, [], None, None      # samples: (time, sm_mhz, max_mhz, set of reasons)
        self.t0 = self.t1 = None
        self.stop = threading.Event()
        self.how = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def _nvml_loop(self, nv, handle):
        bits = ((nv.nvmlClocksThrottleReasonHwSlowdown, "hw_slowdown"), (nv.nvmlClocksThrottleReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_thermal_slowdown"), (nv.nvmlClocksThrottleReasonSwPowerCap, "sw_power_cap"))
        mx = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
        while not self.stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle))
                self.samples.append((time.time(), sm, mx, {name for bit, name in bits if mask & bit}))
            except Exception:      # noqa: BLE001 -- a failed poll is a missing sample, nothing more
                pass
            self.stop.wait(0.01)

    def _smi_pump(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                self.samples.append((time.time(), float(parts[0]), float(parts[1]),
                                     {name for name, flag in zip(self.NAMES, parts[3:7]) if flag.lower().startswith("active")}))
            except ValueError:
                continue

    def __enter__(self):
        if self.index is None:          # ranks other than 0 do not sample
            return self
        try:
            import pynvml as nv

            nv.nvmlInit()
            handle = nv.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[self.index]) if os.environ.get("CUDA_VISIBLE_DEVICES", "").replace(",", "").isdigit() else self.index)
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            self.how = "nvml"
            return self
        except Exception:      # noqa: BLE001
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._smi_pump, daemon=True)
            self.thread.start()
            self.how = "nvidia-smi"
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        self.stop.set()
        if self.proc:
            self.proc.terminate()
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self):
        inside = [x for x in self.samples if self.t0 is not None and self.t1 is not None and self.t0 - 0.02 <= x[0] <= self.t1 + 0.05]
        window = "timed region" if inside else "warm-up + timed region"
        use = inside or self.samples
        sm = [x[1] for x in use]
        reasons = set()
        for x in use:
            reasons |= x[3]
        mx = max([x[2] for x in use], default=0)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm), "window": window,
                "source": self.how}


def nvlink_bytes(index):
    """cumulative NVLink payload bytes (tx, rx) of GPU `index`, summed over its links: the driver's per-link throughput counters
    (`nvidia-smi nvlink -gt d`, KiB); None when the tool or the counters are not available"""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=20).stdout
    except (OSError, subprocess.SubprocessError):
        return None
    tx = rx = 0
    seen = False
    for line in out.splitlines():
        parts = line.replace(",", " ").split()
        if "Tx:" in parts and "KiB" in parts:
            tx += int(parts[parts.index("Tx:") + 1]) * 1024
            seen = True
        if "Rx:" in parts and "KiB" in parts:
            rx += int(parts[parts.index("Rx:") + 1]) * 1024
            seen = True
    return (tx, rx) if seen else None


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return float(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def cpu_sample_size(cores, ncontours, size):
    # ~3 Mvoxel.contour/s per core for the reference on these fields: aim for a few seconds of CPU work per frame
    target = 12e6 * max(1, cores) / max(1, ncontours)
    n = int(round(target ** (1.0 / 3.0) / 16.0)) * 16
    return int(min(size, 512, max(96, n)))


def host_window(w, field_index, n, offset):
    """the n^3 window at `offset` of field `field_index`, computed on the HOST (no CUDA): float32 (nz, ny, nx)"""
    from contour3d_b200 import scenes

    size, spec = w["size"], w["fields"][field_index]
    if spec["kind"] == 0:
        full = scenes.gyroid(size)
        o = offset
        return np.ascontiguousarray(full[o[2]:o[2] + n, o[1]:o[1] + n, o[0]:o[0] + n])
    return scenes.synth_host((size,) * 3, kind=spec["kind"], seed=spec["seed"], octaves=spec["octaves"], gain=spec["gain"],
                             window=(n, n, n), offset=offset)


def sample_case(w, crops, n, path):
    """case file of the bounded CPU sample: the centred n^3 crop of every field of the benchmark, same isovalues / camera / image"""
    from contour3d_b200 import casefile, scenes

    size, (width, height) = w["size"], w["image"]
    g = scenes.uniform_grids(size, size, size)
    o = (size - n) // 2
    grids = [g[d][o:o + n] for d in range(3)]      # the crop stays at its place in the global domain
    contours = [dict(field=f, threshold=v, color=COLORS[i % len(COLORS)]) for i, (f, v) in enumerate(w["contours"])]
    casefile.write_case(path, scenes.scene(width, height), grids, crops, contours, scenes.box_lines(nitems=128, width=4.0))
    return grids, contours


def run_reference_on(case, out, steps, warmup, ranks):
    from oracle import reference

    ranks = max(1, ranks)
    return reference.run_case(case, out, ranks=ranks, dims=(1, ranks, 1), repeat=steps, warmup=warmup)


def reference_arm(args):
    """The unmodified reference (oracle/_ref/ref_driver) on this box's host cores.  Loads nothing of the CUDA library:
    the field comes from the host twin of the device generator (libc3d_synth_host.so, plain C) or numpy."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import reference

    if not reference.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built (no /root/reference at build time)"}))
        return 0
    w = workload(args)
    size, (width, height) = w["size"], w["image"]
    cores = host_cores()
    nct = len(w["contours"])
    n = cpu_sample_size(cores, nct, size)
    o = (size - n) // 2
    full = None
    with tempfile.TemporaryDirectory(dir="/dev/shm") as tmp:
        crops = [host_window(w, f, n, (o, o, o)) for f in range(len(w["fields"]))]
        case = os.path.join(tmp, "sample.bin")
        sample_case(w, crops, n, case)
        del crops
        ranks = min(cores, n // 4)
        res = run_reference_on(case, os.path.join(tmp, "ref.ppm"), args.steps, args.warmup, ranks)
        os.unlink(case)
        # the SAME job once, when the box can hold it (the reference needs ~45 B per voxel at 1024^3: three double copies, slice rings, canvases)
        voxels_full = size ** 3 * len(w["fields"])
        if args.ref_full != "no" and n < size and (args.ref_full == "yes" or (mem_available_gb() > 60.0 * voxels_full / 1e9 + 40.0 and cores >= 16 and voxels_full <= 2.2e9)):
            t0 = time.perf_counter()
            fields = [host_window(w, f, size, (0, 0, 0)) for f in range(len(w["fields"]))]
            gen_s = time.perf_counter() - t0
            case = os.path.join(tmp, "full.bin")
            sample_case(w, fields, size, case)
            del fields
            franks = min(cores, size // 4)
            try:
                fres = run_reference_on(case, os.path.join(tmp, "full.ppm"), 1, 0, franks)
                full = {"value": size ** 3 * nct / fres["seconds_mean"] / 1e9, "unit": "Gvoxel/s", "seconds": fres["seconds_mean"], "ranks": franks,
                        "same_config": True, "field_generation_s": gen_s,
                        "note": "ONE frame of the whole job (same field, size, isovalues, image) through the unmodified reference"}
            except Exception as exc:      # noqa: BLE001 -- a box too small for the full job must not lose the bounded sample
                full = {"failed": str(exc)[-300:]}
    work = n ** 3 * nct
    value = work / res["seconds_mean"] / 1e9
    line = {
        "impl": "reference", "metric": "Gvoxel/s", "value": value, "unit": "Gvoxel/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": res["seconds_mean"] * 1e3, "frames_per_s": 1.0 / res["seconds_mean"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{w['name']}, {width}x{height}",
                   "sample": f"centred {n}^3 crop of the field(s), full image" if n < size else "the whole field(s)",
                   "timing": "wall clock of contour3d_execute, max over ranks",
                   "field": "host twin of the device generator (libc3d_synth_host.so) / numpy: the CUDA library is not loaded"},
        "cpu_baseline": {"value": value, "unit": "Gvoxel/s", "cores": ranks, "kind": "reference",
                         "sample": f"unmodified reference, {ranks} ranks (y split), centred {n}^3 crop x {nct} isovalues, {width}x{height}, mean of {args.steps} frames"},
        "full_job": full,
        "e2e": {"value": value, "unit": "Gvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def make_fields(ctx, w):
    """device-resident fp32 fields of the workload (this rank's block, halo room included)"""
    from contour3d_b200 import scenes

    size = w["size"]
    out = []
    for spec in w["fields"]:
        f = ctx.field((size, size, size), dtype=np.float32)
        if spec["kind"] == 0:
            f.upload_global(scenes.gyroid(size))
        else:
            f.synthesize(kind=spec["kind"], seed=spec["seed"], octaves=spec["octaves"], gain=spec["gain"])
        out.append(f)
    return out


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
    w = workload(args)
    size, (width, height) = w["size"], w["image"]
    ctx = api.Context(device=local, rank=rank, size=world, nccl_id=ident, dims=(1, 1, world))
    mode = api.lib().c3d_context_exchange_mode(ctx.handle)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    fields = make_fields(ctx, w)
    grids = scenes.uniform_grids(size, size, size)
    scene = scenes.scene(width, height)
    lines = scenes.box_lines(nitems=128, width=4.0)
    contours = [dict(field=fields[f], grids=grids, threshold=v, color=COLORS[i % len(COLORS)]) for i, (f, v) in enumerate(w["contours"])]
    nct = len(contours)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def frame(**kw):
        # the reference extends (= exchanges the halos of) every object's array in every frame (contour/extend.c:211-213):
        # the fields' halos are declared stale, so that the exchange runs inside the frame
        if world > 1:
            for f in fields:
                f.mark_dirty()
        return ctx.render(scene, contours, lines, **kw)

    # ---- device-resident arm ----
    stage_sum, launches = {}, 0
    detail_sum = [0.0] * 11
    sweep_ms, sweep_bytes = [], []
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local if rank == 0 else None) as clocks:          # nvidia-smi needs ~0.2 s to start: it runs through the warm-up too
        for _ in range(args.warmup):
            frame(want_rgb=False)
        # (the link counters are read OUTSIDE the barrier-to-barrier region: the tool takes a different time on every rank,
        # and a rank that enters the timed loop late is waited for by all the others in the first frame)
        nvl0 = nvlink_bytes(local) if world > 1 else None
        barrier()
        clocks.mark_start()
        start.record(stream)
        for _ in range(args.steps):
            frame(want_rgb=False)
            t = ctx.timings()
            launches += t["launches"]
            sweep_ms += t["sweep_each"]
            sweep_bytes += t["sweep_bytes"]
            for k in ("halo", "clear", "sweep", "expand", "raster", "lines", "resolve", "reduce", "shade", "pack", "download", "total"):
                stage_sum[k] = stage_sum.get(k, 0.0) + t[k]
            detail_sum = [a + b for a, b in zip(detail_sum, t["detail"])]
        end.record(stream)
        barrier()
        clocks.mark_end()
        nvl1 = nvlink_bytes(local) if world > 1 else None
        time.sleep(0.05)
    elapsed_ms = start.elapsed_time(end)
    tt = torch.tensor([elapsed_ms], device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    elapsed_ms = float(tt.item())
    active = t["active_cubes"]
    # one extra, untimed frame: the image itself (must not depend on the number of ranks)
    img = frame(want_rgb=True)
    image_md5 = hashlib.md5(img.tobytes()).hexdigest() if rank == 0 else None
    covered = float((img != np.array(scene["bg"], dtype=np.uint8)).any(axis=2).mean()) if rank == 0 else None
    del img
    # where each rank spent the frame
    per_rank = None
    if world > 1:
        mine = {k: round(v / args.steps, 4) for k, v in stage_sum.items()}
        mine["cut_cubes_listed"] = int(sum(active[:nct]))
        mine["cubes_walked_by_raster"] = int(t["list_entries"])
        mine["culled_cubes"] = int(t["culled_cubes"])
        mine["tile_share"] = [round(v, 4) for v in t["tile_share"][:world]]
        names = ("barrier0", "halo_pull", "stage_clear_sweep", "to_outboxes", "barrier1", "merge_lists", "raster_lines_select", "barrier2", "work_pull", "shading", "barrier3")
        mine["detail"] = {k: round(v / args.steps, 4) for k, v in zip(names, detail_sum)}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        per_rank = gathered
    voxels = size ** 3 * nct
    value = voxels * args.steps / (elapsed_ms * 1e-3) / 1e9
    # NVLink traffic of the timed frames by the driver's link counters (every rank reads its own GPU's)
    nvlink = None
    if world > 1:
        mine = None
        if nvl0 is not None and nvl1 is not None:
            mine = {"tx_bytes_per_step": (nvl1[0] - nvl0[0]) / args.steps, "rx_bytes_per_step": (nvl1[1] - nvl0[1]) / args.steps}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if all(g is not None for g in gathered):
            ms = elapsed_ms / args.steps
            nvlink = {"per_rank": gathered, "source": "nvidia-smi nvlink -gt d (driver link counters, KiB granularity) around the timed frames",
                      "max_rx_GBps_over_the_frame": max(g["rx_bytes_per_step"] for g in gathered) / (ms * 1e-3) / 1e9,
                      "total_rx_bytes_per_step": sum(g["rx_bytes_per_step"] for g in gathered)}

    # ---- the exchange steps on their own (N > 1) ----
    collectives = None
    if world > 1:
        iters = 10
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fields[0].exchange()
        barrier()
        h0.record(stream)
        for _ in range(iters):
            fields[0].exchange()
        h1.record(stream)
        barrier()
        halo_ms = h0.elapsed_time(h1) / iters
        faces = (1 if rank > 0 else 0) + (1 if rank < world - 1 else 0)           # z slabs: interior faces of this rank
        halo_bytes = faces * 2 * fields[0].ext[0] * fields[0].ext[1] * 4          # 2 planes pulled per face, fp32
        mine = {"halo_ms": halo_ms, "halo_bytes": halo_bytes}
        if mode == 1:
            ar_ms, rd_ms = ctx.collectives_probe(width, height, iters)
            mine.update(depth_allreduce_ms=ar_ms, rgba_reduce_ms=rd_ms)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        hm = max(g["halo_ms"] for g in gathered)
        hb = max(g["halo_bytes"] for g in gathered)
        collectives = {"halo_exchange": {"bytes_in_per_interior_rank": hb, "ms": hm, "GBps_in_per_rank": hb / (hm * 1e-3) / 1e9,
                                         "note": "two flag barriers + the pull of 2 planes per interior face, timed alone (CUDA events, max over ranks)"}}
        if mode == 1:
            npix = width * height
            ar = max(g["depth_allreduce_ms"] for g in gathered)
            rd = max(g["rgba_reduce_ms"] for g in gathered)
            collectives["depth_allreduce"] = {"bytes": 8 * npix, "ms": ar, "algbw_GBps": 8 * npix / (ar * 1e-3) / 1e9,
                                              "busbw_GBps": 2.0 * (world - 1) / world * 8 * npix / (ar * 1e-3) / 1e9}
            collectives["rgba_reduce"] = {"bytes": 4 * npix, "ms": rd, "algbw_GBps": 4 * npix / (rd * 1e-3) / 1e9}

    # ---- e2e arms: contour3d_execute with host / device double arrays, PPM to /dev/shm ----
    e2e = e2e_dev = None
    if not args.no_e2e:
        e2e, e2e_dev = e2e_arms(args, w, ctx, fields, grids, scene, lines, rank, world, barrier)

    # ---- roofline of the sweep ----
    peak, peak_src = peaks()
    sweep_avg_ms = float(np.mean(sweep_ms))
    achieved = float(np.mean(sweep_bytes)) / (sweep_avg_ms * 1e-3) / 1e9
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(tfile) and args.config == 3 and world == 1:
        traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
    nthr = t["sweep_thresholds"][0]
    roofline = {"kernel": f"k_sweep_tiles<float,{nthr},...> (c3d_sweep.cu; one pass classifies {nthr} iso-values)", "bound": "hbm", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_launch": float(np.mean(sweep_bytes)), "ms_per_launch": sweep_avg_ms, "launches_per_step": len(sweep_ms) / args.steps,
                "share_of_step": stage_sum["sweep"] / stage_sum["total"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, w)

    if rank == 0:
        line = {
            "metric": "Gvoxel/s", "value": value, "unit": "Gvoxel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "frames_per_s": args.steps / (elapsed_ms * 1e-3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 field / f64 geometry", "data": "synthetic",
            "config": {"workload": f"{w['name']} {[v for _, v in w['contours']]}, {width}x{height}, 8 box lines, z slabs dims={{1,1,{world}}}",
                       "l2": f"inputs larger than L2 ({4 * size ** 3 / 1e9:.1f} GB per field vs 126 MB)", "timing": "CUDA events on the library stream, max over ranks",
                       "exchange": {0: "one rank", 1: "NCCL: halo send/recv, whole-image all-reduce + reduce",
                                    2: "peer memory (CUDA IPC over NVLink): halo pull, screen-partitioned raster, shading at the cube's owner, colours stored into rank 0's image"}[mode],
                       "halo_in_step": world > 1},
            "clocks": clocks.summary(),
            "e2e": e2e,
            "e2e_device": e2e_dev,
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "image_md5": image_md5, "image_covered_fraction": covered,
            "stages_ms": {k: v / args.steps for k, v in stage_sum.items()},
            "stages_ms_per_rank": per_rank,
            "collectives": collectives,
            "nvlink_counters": nvlink,
            "active_cubes": active, "culled_cubes": t["culled_cubes"], "culled_triangles": t["culled_triangles"], "shade_spills": t["shade_spills"],
        }
        print(json.dumps(line))
    barrier()
    for f in fields:
        f.destroy()
    ctx.destroy()
    if world > 1:
        dist.destroy_process_group()
    return 0


def e2e_arms(args, w, ctx, fields, grids, scene, lines, rank, world, barrier):
    """contour3d_execute (the public drop-in) with HOST double arrays from pinned memory, then with DEVICE double arrays"""
    import torch
    from contour3d_b200 import api

    size = w["size"]
    m = fields[0].mysizes
    local_bytes = 8 * m[0] * m[1] * m[2] * len(fields)
    if local_bytes > args.e2e_max_gb * 1e9:
        note = {"skipped": f"{local_bytes / 1e9:.1f} GB of host doubles per rank exceed --e2e-max-gb {args.e2e_max_gb}"}
        return note, note
    host, pins = [], []
    for f in fields:
        # host copy of my pencil, widened exactly to double (what a solver would hold), in pinned memory
        ext = f.download_ext()
        lo = [f.offsets[d] - f.ext_off[d] for d in range(3)]
        local = ext[lo[2]:lo[2] + m[2], lo[1]:lo[1] + m[1], lo[0]:lo[0] + m[0]]
        pinned = torch.empty(local.shape, dtype=torch.float64, pin_memory=True)
        pinned.numpy()[...] = local
        del ext, local
        pins.append(pinned)
        host.append(pinned.numpy())
    info = api.sdecomp_construct((1, 1, world))
    contours = [dict(field=f, threshold=v, color=COLORS[i % len(COLORS)]) for i, (f, v) in enumerate(w["contours"])]
    out = f"/dev/shm/contour3d_b200_bench_{os.getpid()}.ppm"
    steps = max(1, min(args.steps, args.e2e_steps))
    voxels = size ** 3 * len(contours)
    npix3 = int(3 * scene["width"] * scene["height"])

    def timed(arrays, reps):
        rc = api.execute(scene, grids, arrays, contours, lines, out, sdecomp_info=info)
        assert rc == 0, "contour3d_execute failed"
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            rc = api.execute(scene, grids, arrays, contours, lines, out, sdecomp_info=info)
            assert rc == 0
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device="cuda")
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    dt = timed(host, steps)
    md5_host = hashlib.md5(open(out, "rb").read()).hexdigest() if rank == 0 else None
    e2e = {"value": voxels * steps / dt / 1e9, "unit": "Gvoxel/s", "frames_per_s": steps / dt, "ms_per_step": dt / steps * 1e3, "steps": steps,
           "h2d_bytes_per_step": int(sum(h.nbytes for h in host) + 3 * 8 * size), "d2h_bytes_per_step": npix3 if rank == 0 else 0,
           "ppm_md5": md5_host,
           "path": "contour3d_execute(host double arrays, pinned) -> PPM in /dev/shm; wall clock around the call, H2D of the array and D2H of the image inside"}
    # the same call as a GPU-resident solver makes it: the arrays already live on the device
    dev = [p.cuda(non_blocking=False) for p in pins]
    torch.cuda.synchronize()
    dsteps = max(steps, min(args.steps, 10))
    dt = timed([int(d.data_ptr()) for d in dev], dsteps)
    md5_dev = hashlib.md5(open(out, "rb").read()).hexdigest() if rank == 0 else None
    e2e_dev = {"value": voxels * dsteps / dt / 1e9, "unit": "Gvoxel/s", "frames_per_s": dsteps / dt, "ms_per_step": dt / dsteps * 1e3, "steps": dsteps,
               "h2d_bytes_per_step": int(3 * 8 * size), "d2h_bytes_per_step": npix3 if rank == 0 else 0, "ppm_md5": md5_dev,
               "path": "contour3d_execute(DEVICE double arrays) -> PPM in /dev/shm; wall clock around the call; one rank: the array is swept where it lies, "
                       "several ranks: copied device-to-device into the array with halo room"}
    if rank == 0 and os.path.exists(out):
        os.unlink(out)
    del dev
    return e2e, e2e_dev


def cpu_baseline(args, w):
    """the unmodified reference on a bounded crop, and the GPU path on the same crop: pixel-by-pixel parity at benchmark data"""
    from contour3d_b200 import api, scenes
    from oracle import oracle, reference

    size, (width, height) = w["size"], w["image"]
    cores = host_cores()
    nct = len(w["contours"])
    n = min(cpu_sample_size(cores, nct, size), 384)
    o = (size - n) // 2
    crops = [host_window(w, f, n, (o, o, o)) for f in range(len(w["fields"]))]
    work = n ** 3 * nct
    with tempfile.TemporaryDirectory(dir="/dev/shm") as tmp:
        if reference.available():
            case = os.path.join(tmp, "sample.bin")
            grids, contours = sample_case(w, crops, n, case)
            ranks = min(cores, n // 4)
            ref_ppm = os.path.join(tmp, "ref.ppm")
            res = run_reference_on(case, ref_ppm, 1, 1, ranks)
            secs = res["seconds_mean"]
            # the same crop case through the drop-in on the GPU
            ours_ppm = os.path.join(tmp, "ours.ppm")
            arrays = [c.astype(np.float64) for c in crops]
            rc = api.execute(scenes.scene(width, height), grids, arrays, contours, scenes.box_lines(nitems=128, width=4.0), ours_ppm)
            parity = {"error": "contour3d_execute failed"}
            if rc == 0:
                a, b = api.read_ppm(ours_ppm), api.read_ppm(ref_ppm)
                d = np.abs(a.astype(np.int32) - b.astype(np.int32))
                parity = {"pixels": int(a.shape[0] * a.shape[1]), "pixels_differing": int((d.max(axis=2) > 0).sum()), "max_lsb": int(d.max()),
                          "covered_fraction": float((b != 0).any(axis=2).mean()),
                          "what": f"centred {n}^3 crop x {nct} isovalues, {width}x{height}: contour3d_execute on the GPU vs the unmodified reference's PPM"}
            return {"value": work / secs / 1e9, "unit": "Gvoxel/s", "cores": ranks, "kind": "reference", "seconds": secs,
                    "sample": f"unmodified reference (oracle/_ref), {ranks} ranks y-split, centred {n}^3 crop x {nct} isovalues, {width}x{height}, 1 frame after 1 warm-up",
                    "crop_parity": parity}
    # the restatement, threads emulating ranks
    g = scenes.uniform_grids(size, size, size)
    grids = [g[d][o:o + n] for d in range(3)]
    cs = [dict(field=crops[f], grids=grids, threshold=v, color=COLORS[i % len(COLORS)]) for i, (f, v) in enumerate(w["contours"])]
    t0 = time.perf_counter()
    oracle.frame(scenes.scene(width, height), cs, scenes.box_lines(nitems=128, width=4.0), nthreads=cores)
    secs = time.perf_counter() - t0
    return {"value": work / secs / 1e9, "unit": "Gvoxel/s", "cores": cores, "kind": "port", "seconds": secs,
            "sample": f"oracle restatement, {cores} threads y-split, centred {n}^3 crop x {nct} isovalues, {width}x{height}, 1 frame"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, help="BASELINE.json configuration: 2, 3 (default, the metric's), 4 or 5")
    ap.add_argument("--size", type=int, default=0)
    ap.add_argument("--image", type=lambda s: tuple(int(v) for v in s.lower().split("x")), default=None)
    ap.add_argument("--isovalues", type=lambda s: tuple(float(v) for v in s.split(",")), default=None)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-max-gb", type=float, default=20.0, help="skip the e2e arms when a rank's host doubles exceed this")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-full", default="auto", choices=["auto", "yes", "no"], help="reference arm: also run ONE frame of the whole job")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    return reference_arm(args) if args.impl == "reference" else ours(args)


if __name__ == "__main__":
    sys.exit(main())
