#!/usr/bin/env python
"""Benchmark of the georeference + regrid hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on host cores

One "step" = one full pass of the hot path over one frame per GPU:
    getMapping (TAN WCS -> rays -> WGS84+110 km intersection -> lat/lon, MLat/MLT, elevation for
    all pixel corners and centres, mask sanitisation) + resample(arcsecPerPx=100, 'mean').
N == 1 runs BASELINE.json configs[1] (one synthetic ISS Nikon D3S frame, 4256x2832, FP64, MLat/MLT
outputs included); N > 1 runs one frame of the configs[3] sequence per rank and step (frame-sharded,
no data-path collective, weak scaling).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "georeferenced+resampled Mpix/s"
UNIT = "Mpix/s"
ALGO_BYTES_PER_PIXEL_GEOREF = 72.0      # 4 f64 per corner + 5 f64 per centre written
ALGO_FLOP_PER_PIXEL_GEOREF = 290.0      # algorithmic FP64 ops per pixel
ARCSEC_PER_PX = 100


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None,
                    help="timed steps (default: 200 for the B200 arm -- a step is ~0.5 ms -- and 2 for --impl reference)")
    ap.add_argument("--warmup", type=int, default=None, help="untimed steps (default: 10 / 1)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=4256)
    ap.add_argument("--height", type=int, default=2832)
    ap.add_argument("--fast-center", action="store_true", help="fastCenterCalculation=True variant")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-scale", type=int, default=1,
                    help="linear down-scale of the frame used for the CPU baseline sample")
    args = ap.parse_args()
    ref = args.impl == "reference"
    if args.steps is None:
        args.steps = 2 if ref else 200
    if args.warmup is None:
        args.warmup = 1 if ref else 10
    return args


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "25"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, smmax, reasons = [], [], set()
        try:
            with open(self.path) as fh:
                for line in fh:
                    p = [x.strip() for x in line.split(",")]
                    if len(p) < 9:
                        continue
                    try:
                        sm.append(float(p[1]))
                        smmax.append(float(p[2]))
                    except ValueError:
                        continue
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                         p[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = [s for s in sm if s > 0.5 * max(sm)] or sm
            out = {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(smmax)),
                   "reasons": sorted(reasons), "samples": len(sm)}
        return out


# --------------------------------------------------------------------- CPU (oracle) legs
def _oracle_frame(args):
    """One full pass of the reference algorithm (numpy restatement in oracle/) over a frame."""
    width, height, seed, fast = args
    import io
    import contextlib
    from auromat_b200 import synthetic
    import oracle.auromat_oracle as O
    hdr = synthetic.issHeader(width, height)
    img = synthetic.issImage(width, height, seed)
    t, cam = synthetic.headerTimeAndCamera(hdr)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        geo = O.georeference(hdr, cam, t, 110, fast_center=fast)
        if not fast:
            mk, mc = O.sanitize_masks(np.isnan(geo['lats']), np.isnan(geo['latsCenter']))
            for n in ('lats', 'lons', 'mlat', 'mlt'):
                geo[n][mk] = np.nan
            for n in ('latsCenter', 'lonsCenter', 'mlatCenter', 'mltCenter', 'elevation'):
                geo[n][mc] = np.nan
        O.resample_frame(geo, img, 110, arcsec_per_px=ARCSEC_PER_PX)
    return time.perf_counter() - t0


def cpu_baseline_single(width, height, fast, scale):
    w, h = width // scale, height // scale
    dt = _oracle_frame((w, h, 0, fast))
    return {"value": w * h / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "1 frame %dx%d, numpy oracle port of the reference path, %.1f s" % (w, h, dt)}


def run_reference(args, rank, world):
    """`--impl reference`: the reference algorithm (oracle port; the reference itself is pure
    Python + absent third-party wheels and cannot travel to the GPU box) on all host cores,
    one reduced frame per worker process and step."""
    if rank != 0:
        return
    import multiprocessing as mp
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    scale = 4
    w, h = args.width // scale, args.height // scale
    steps, warmup = max(1, min(args.steps, 5)), max(0, min(args.warmup, 1))
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        for _ in range(warmup):
            pool.map(_oracle_frame, [(w, h, i, args.fast_center) for i in range(workers)])
        t0 = time.perf_counter()
        for s in range(steps):
            pool.map(_oracle_frame, [(w, h, s * workers + i, args.fast_center) for i in range(workers)])
        dt = time.perf_counter() - t0
    value = steps * workers * w * h / dt / 1e6
    sample = "%d worker processes x 1 frame %dx%d per step (1/%d of the %dx%d frame's pixels, same field of view)" % (
        workers, w, h, scale * scale, args.width, args.height)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "frames_per_s": steps * workers / dt * (w * h) / (args.width * args.height),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, n_gpus):
    return {
        "workload": ("BASELINE configs[1]: one synthetic ISS Nikon D3S frame %dx%d, TAN WCS, 110 km, getMapping "
                     "(corners+centres, lat/lon+MLat/MLT+elevation) + resample(arcsecPerPx=100, mean)"
                     % (args.width, args.height)) if n_gpus == 1 else
                    ("BASELINE configs[3]: synthetic ISS sequence, one %dx%d frame per GPU and step, frame i -> "
                     "rank i mod N, no collective" % (args.width, args.height)),
        "frame": [args.width, args.height], "altitude_km": 110, "arcsec_per_px": ARCSEC_PER_PX,
        "fast_center": bool(args.fast_center), "outputs": "lat/lon/MLat/MLT corners+centres, elevation, resampled RGB+elevation",
        "parallelism": "frames x%d" % n_gpus,
        "api": "auromat_b200.pipeline.resampleSequence (getMapping + resample per frame, pipelined: georeferencing of 3 frames enqueued ahead)",
        "l2": "per-step working set ~0.9 GB (72 B/px planes + image) exceeds the 126 MB L2; no explicit flush",
    }


# ------------------------------------------------------------------------------ GPU arm
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from auromat_b200 import synthetic
    from auromat_b200.mapping.spacecraft import getMapping
    from auromat_b200.runtime import get_context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        from auromat_b200.parallel import bindToLocalCpus
        bindToLocalCpus(local_rank)          # NUMA-local pinned buffers and launch thread
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = get_context(local_rank)
    W, H = args.width, args.height
    npx = W * H
    total = args.warmup + args.steps

    if world == 1:
        headers = [synthetic.issHeader(W, H)] * total
    else:
        # frames of the configs[3] sequence, frame i -> rank i mod N; a cyclic window of 32 frames keeps
        # the per-GPU work (fraction of pixels that see the Earth) the same for every N (weak scaling)
        seq = synthetic.sequenceHeaders(32, W, H)
        headers = [seq[(s * world + rank) % 32] for s in range(total)]
    img_host = torch.from_numpy(synthetic.issImage(W, H, seed=1000 + rank)).pin_memory()
    img_dev = img_host.to(ctx.torch_device)

    from auromat_b200.pipeline import resampleSequence

    # `value`: inputs resident in HBM -- the public sequence API (getMappingSequence + ResampleProvider
    # of the reference, pipelined) over K frames; every frame runs the full path from scratch.
    def run_device(hdrs):
        last = None
        for f in resampleSequence([img_dev] * len(hdrs), hdrs, arcsecPerPx=ARCSEC_PER_PX, magnetic=True,
                                  fastCenterCalculation=args.fast_center, toHost=False, device=local_rank,
                                  ringBuffers=True):
            last = f
        return last

    # `e2e`: same call with HOST buffers: pinned image in (H2D every frame), resampled image, mask and
    # elevation out (D2H every frame into pinned buffers)
    transfer = {}

    def run_e2e(hdrs, sparse=True):
        last = None
        transfer.clear()
        for f in resampleSequence([img_host.numpy()] * len(hdrs), hdrs, arcsecPerPx=ARCSEC_PER_PX, magnetic=True,
                                  fastCenterCalculation=args.fast_center, toHost=True, device=local_rank,
                                  ringBuffers=True, transferStats=transfer, sparseUpload=sparse):
            last = f
        transfer['frames'] = len(hdrs)
        last._finish()                 # the last frame's results are in its pinned host buffers
        return None, None, last

    def run_e2e_full_upload(hdrs):
        return run_e2e(hdrs, sparse=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        fn(headers[:args.warmup])
        barrier()
        launches0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        out = fn(headers[args.warmup:args.warmup + args.steps])
        t_host = time.perf_counter() - t_host
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ctx.launch_count - launches0
        if os.environ.get("AMT_BENCH_RANKLOG"):
            # per-rank diagnostics (stderr): device time of this rank and the wall time its host loop needed
            sys.stderr.write("[rank %d] %s: device %.4f ms/step, host loop %.4f ms/step, cpus %d\n" % (
                rank, getattr(fn, "__name__", "variant"), ms / args.steps, t_host * 1e3 / args.steps,
                len(os.sched_getaffinity(0))))
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=ctx.torch_device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, out

    # secondary variants (reported under "variants", not the headline): plane-free resampling
    # (no coordinate planes, no MLat/MLT) and fastCenterCalculation=True
    def run_variant(hdrs, **kw):
        last = None
        for f in resampleSequence([img_dev] * len(hdrs), hdrs, arcsecPerPx=ARCSEC_PER_PX, toHost=False,
                                  device=local_rank, ringBuffers=True, **kw):
            last = f
        return last

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, launches, out = timed(run_device)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _, out_e2e = timed(run_e2e)
    d2h = int(sum(b.numel() * b.element_size() for b in out_e2e[2]._host))
    # image bytes actually copied per frame: only the row range that holds georeferenced pixels is
    # uploaded (pipeline.resampleSequence(sparseUpload=True)); the complete frame would be h2d_full
    h2d_full = int(img_host.numel() * img_host.element_size())
    h2d = int(transfer['h2d_bytes'] // max(1, transfer['frames']))
    variants = {}
    if world == 1:
        # best of 3 timed passes each (secondary numbers; the first pass also warms the allocator for
        # the variant's own buffer sizes)
        ms_pf = min(timed(lambda h: run_variant(h, magnetic=False, coordinates=False))[0] for _ in range(3))
        ms_fc = min(timed(lambda h: run_variant(h, magnetic=True, fastCenterCalculation=True))[0] for _ in range(3))
        ms_full = min(timed(run_e2e_full_upload)[0] for _ in range(2))
        variants = {
            "e2e_full_image_upload": {"value": args.steps * npx / (ms_full * 1e-3) / 1e6, "unit": UNIT,
                                      "ms_per_step": ms_full / args.steps, "h2d_bytes_per_step": int(
                                          img_host.numel() * img_host.element_size()),
                                      "note": "e2e with sparseUpload=False: the complete 36 MB frame is copied every "
                                              "step although the rows above the limb never influence the result; "
                                              "PCIe-bound"},
            "plane_free_resample_only": {"value": args.steps * npx / (ms_pf * 1e-3) / 1e6, "unit": UNIT,
                                         "ms_per_step": ms_pf / args.steps,
                                         "note": "resampleSequence(coordinates=False): hit bitmaps + outline stats + "
                                                 "fused georeference/binning kernel, no planes, no MLat/MLT"},
            "fast_center": {"value": args.steps * npx / (ms_fc * 1e-3) / 1e6, "unit": UNIT,
                            "ms_per_step": ms_fc / args.steps, "note": "fastCenterCalculation=True, all 9 planes"},
        }
    # dominant kernel alone: the fused georeference kernel (all 9 planes)
    m = getMapping(img_dev, headers[0], fastCenterCalculation=args.fast_center, identifier="roofline")
    frame = m.frameConstants
    nk, nc = (W + 1) * (H + 1), npx
    planes = {n: ctx.empty(nk if n.endswith('_k') else nc, torch.float64)
              for n in ('lat_k', 'lon_k', 'mlat_k', 'mlt_k', 'lat_c', 'lon_c', 'mlat_c', 'mlt_c', 'elev_c')}
    # + the validity bitmaps: exactly the launch the sequence pipeline issues
    planes['valid_k'], planes['valid_c'] = ctx.new_bitmaps(W, H)
    for _ in range(3):
        ctx.georef(frame, planes)
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ctx.georef(frame, planes)
    e1.record()
    torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / reps
    fp64_peak = ctx.measure_fp64_peak() if rank == 0 else 0.0     # warp-lane DFMA/s (2 FLOP each)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, peak_kind = measured_peaks()
    value = args.steps * world * npx / (ms_dev * 1e-3) / 1e6
    e2e_value = args.steps * world * npx / (ms_e2e * 1e-3) / 1e6
    achieved = ALGO_BYTES_PER_PIXEL_GEOREF * npx / (k_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
        "frames_per_s": args.steps * world / (ms_dev * 1e-3),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "h2d_bytes_full_frame": h2d_full,
                "ms_per_step": ms_e2e / args.steps, "frames_per_s": args.steps * world / (ms_e2e * 1e-3)},
        "gpu_launches": int(launches),
        "variants": variants,
        "clocks": clocks,
        "roofline": {
            "kernel": "k_georef_tiles" if args.fast_center else "k_georef_points",
            "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "peak_kind": peak_kind,
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel for this frame size from the
            # ncu --set full capture in profiles/r01_kernels_v8_ncu.txt (writes only; part of the
            # last planes is still in L2 when the kernel ends, hence < algorithmic bytes)
            "traffic": 812.4e6 if (W, H) == (4256, 2832) and not args.fast_center else None,
            "kernel_ms": k_ms, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PIXEL_GEOREF * npx,
            # the kernel is FP64-pipe / issue bound, not HBM bound: algorithmic FP64 rate
            # (290 reference-formula ops per pixel) against the DFMA peak measured just now;
            # executed-instruction pipe utilisation is in profiles/ (ncu sm__pipe_fp64_cycles_active)
            "fp64": {"algorithmic_tflops": ALGO_FLOP_PER_PIXEL_GEOREF * npx / (k_ms * 1e-3) / 1e12,
                     "peak_tflops_measured": 2 * fp64_peak / 1e12,
                     "dfma_issue_peak_ginst": fp64_peak / 1e9,
                     "ncu_pipe_fp64_pct": 64.4, "ncu_issue_active_pct": 65.4,
                     "ncu_source": "profiles/r01_kernels_v8_ncu.txt"},
        },
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_single(W, H, args.fast_center, args.cpu_baseline_scale)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
