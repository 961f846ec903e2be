#!/usr/bin/env python
"""Benchmark of the area-average interpolation hot path (BASELINE.json metric: output Mpixels/s, device-timed).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 4] [--impl ours|reference]

A "step" is one pass of the hot path over the whole workload (default: BASELINE config 4, the 16384x16384
float32 slice, 0.37x, 17.3 deg -> 7591x7591 canvas).  With N > 1 (torchrun, one process per GPU) the canvas is
split into row bands balanced by covered pixels; every rank holds its source halo + its band; there is no
data-path collective (bands are independent) -- only the barrier and the max-over-ranks of the timing.

value     = canvas Mpixels/s with the source resident in HBM, CUDA events around exactly K steps (max over ranks)
e2e       = the same metric through the host-buffer C-ABI call (aai_run_host_band): pinned host source ->
            device, kernel, device -> pinned host canvas, all inside the timed region
roofline  = the dominant kernel against its bound (FP32 pipe for the rotated clip path, HBM for the separable path)
cpu_baseline = the reference's own CPU implementation (compiled upstream Source.cpp, 1 thread as shipped) on a
            bounded replica of the workload, timed on this box (rank 0, N = 1)

--impl reference times the upstream CPU implementation on all host cores (one process per core, one replica image
each per step).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "output Mpixels/s (device-timed)"
UNIT = "Mpix/s"

# BASELINE.json configs (SURVEY.md §8d fixes isocentres and generators); index = config number
CONFIGS = {
    1: dict(label="cfg1: 512x512 u8 grayscale, 0.5x, 0 deg", w=512, h=512, dtype="uint8", ch=1, ratio=0.5,
            angle=0.0, iso=(256.0, 256.0), batch=1, replica=512, baseline_replica=512),
    2: dict(label="cfg2: 2048x2048 u8 grayscale, 0.37x, 30 deg", w=2048, h=2048, dtype="uint8", ch=1, ratio=0.37,
            angle=30.0, iso=(1024.0, 1024.0), batch=1, replica=512, baseline_replica=2048),
    3: dict(label="cfg3: 8192x8192 RGB u8, 1.7x, 45 deg", w=8192, h=8192, dtype="uint8", ch=3, ratio=1.7, angle=45.0,
            iso=(4095.5, 4095.5), batch=1, replica=128, baseline_replica=384),
    4: dict(label="cfg4: 16384x16384 float32, 0.37x, 17.3 deg", w=16384, h=16384, dtype="float32", ch=1, ratio=0.37,
            angle=17.3, iso=(8192.0, 8192.0), batch=1, replica=512, baseline_replica=2048),
    5: dict(label="cfg5: 256 x 4096x4096 float32, 0.5x, 0 deg", w=4096, h=4096, dtype="float32", ch=1, ratio=0.5,
            angle=0.0, iso=(2048.0, 2048.0), batch=256, replica=512, baseline_replica=1024),
}


def replica_of(cfg, side):
    """Same ratio / angle / isocentre rule on a smaller source (per-pixel CPU cost is size independent)."""
    iso = side / 2.0 - (cfg["w"] / 2.0 - cfg["iso"][0])  # same offset from the image centre
    return dict(w=side, h=side, ratio=cfg["ratio"], angle=cfg["angle"], iso=(iso, iso))


def flops_per_covered_pixel(side, c, s, channels):
    """SURVEY.md §8d contract figure: F = 48 + 128 N_bnd + 2 N_int (+2 per extra channel and touched cell)."""
    n_int = max(0.0, side - c - s) ** 2
    n_all = side * side + 2 * side * (c + s) + 1
    n_bnd = n_all - n_int
    return 48 + 128 * n_bnd + 2 * n_int + 2 * (channels - 1) * n_all


# ---- reference arm -----------------------------------------------------------------------------------------------

def _ref_worker(args):
    side, ratio, angle, iso, seed, kind = args
    sys.path.insert(0, ROOT)
    from area_average_interpolation_b200.synthetic import synthetic_image
    from oracle import port, ref

    src = synthetic_image(side, side, np.float64, seed)
    t0 = time.perf_counter()
    if kind == "reference":
        ok, msg, dst, _, _ = ref.run(src, 1.0, ratio, iso, angle, mode=1)
        assert ok, msg
    else:
        st, dst, _ = port.run(src, 1.0, ratio, iso, angle, mode=1, threads=1)
        assert st == 0
    return dst.size, time.perf_counter() - t0


def cpu_kind():
    from oracle import port, ref

    if ref.available:
        return "reference"
    port.lib()
    return "port"


def reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp

    kind = cpu_kind()
    cores = os.cpu_count() or 1
    rep = replica_of(cfg, cfg["replica"])
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        tasks = [(rep["w"], rep["ratio"], rep["angle"], rep["iso"], 20201 + k, kind) for k in range(cores)]
        for _ in range(args.warmup):
            pool.map(_ref_worker, tasks)
        t0 = time.perf_counter()
        pixels = 0
        for _ in range(args.steps):
            pixels += sum(n for n, _ in pool.map(_ref_worker, tasks))
        dt = time.perf_counter() - t0
    value = pixels / dt / 1e6
    sample = (f"{cores} processes x one {rep['w']}x{rep['h']} float64 replica of {cfg['label']} per step "
              f"(same ratio/angle; per-pixel CPU cost is size independent)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["label"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---- clocks ------------------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_uuid):
        self.proc = None
        self.path = None
        self.uuid = device_uuid

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            cmd = ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"]
            if self.uuid:
                cmd += ["-i", self.uuid]
            self.proc = subprocess.Popen(cmd, stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


# ---- our arm -----------------------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(torch, local):
    """One process per GPU: run on the CPUs of the GPU's own NUMA node, so that the pinned host buffers (first touch)
    and the copy threads sit next to its PCIe root port.  With 8 ranks uploading at once, buffers on the far socket cap
    the upload at ~23 GB/s per rank (measured; 35-50 GB/s from the near socket).  Best effort: any failure leaves the
    affinity as it was."""
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        devid = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0"
        node = int(open(os.path.join(path, "numa_node")).read().strip())
        cpulist = open(os.path.join(path, "local_cpulist")).read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if node >= 0 and cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception as exc:  # no sysfs entry, restricted container, ...
        print(f"[bench] NUMA binding skipped: {exc}", file=sys.stderr)
    return None


def our_arm(args, cfg):
    import torch
    import torch.distributed as dist

    import area_average_interpolation_b200 as aai
    from area_average_interpolation_b200.sharding import band_for_rank, batch_slice
    from area_average_interpolation_b200.synthetic import synthetic_image

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.item()

    if args.arith == "auto":
        arith = aai.ARITH_F64 if cfg["dtype"] == "float64" else aai.ARITH_F32
    else:
        arith = aai.ARITH_F32 if args.arith == "f32" else aai.ARITH_F64
    np_dt = np.dtype(cfg["dtype"])
    t_dt = {"uint8": torch.uint8, "float32": torch.float32, "float64": torch.float64}[cfg["dtype"]]
    out_dt = torch.float64 if cfg["dtype"] == "float64" else torch.float32
    W, H, CH = cfg["w"], cfg["h"], cfg["ch"]
    plan = aai.make_plan(W, H, 1.0, cfg["ratio"], cfg["iso"], cfg["angle"])
    assert plan.status == 0, plan.message
    seed = 20201 + args.config
    stream = torch.cuda.current_stream().cuda_stream
    shape_tail = (CH,) if CH > 1 else ()

    peer = None
    if cfg["batch"] == 1:
        # one image, canvas row bands
        band = band_for_rank(plan, rank, world)
        host_dst = torch.empty((band.rows, plan.dst_w) + shape_tail, dtype=out_dt, pin_memory=True)
        dev_dst = torch.empty((band.rows, plan.dst_w) + shape_tail, dtype=out_dt, device=dev)
        di = aai.tensor_image(dev_dst, y0=band.row0, height=plan.dst_h)
        hdi = aai.tensor_image(host_dst, y0=band.row0, height=plan.dst_h)
        my_pixels = band.rows * plan.dst_w
        launches_per_step = 1
        d2h = band.rows * plan.dst_w * CH * host_dst.element_size()
        if world > 1 and not args.no_peer:
            # every source row crosses PCIe once (its owner uploads it); halos are pulled over NVLink (CUDA IPC)
            try:
                from area_average_interpolation_b200.sharding import PeerSource

                def gather(obj):
                    out = [None] * world
                    dist.all_gather_object(out, obj)
                    return out

                peer = PeerSource(plan, aai._NP_TO_AAI[np_dt], CH, rank, world, local, band, gather)
            except Exception as exc:  # IPC unavailable: every rank uploads its own halo from the host
                print(f"[bench] rank {rank}: peer source exchange unavailable ({exc}); uploading halos from host",
                      file=sys.stderr)
                peer = None
            ok = torch.tensor([1 if peer is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0 and peer is not None:
                peer.close()
                peer = None
        if peer is not None:
            o0, o1 = peer.owned_rows()
            host_src = torch.empty((o1 - o0, W) + shape_tail, dtype=t_dt, pin_memory=True)
            synthetic_image(W, H, np_dt, seed, channels=CH, y0=o0, rows=o1 - o0, out=host_src.numpy())
            hsi = aai.tensor_image(host_src, y0=o0, height=H)
            si = peer.full
            resident_bytes = (band.src_y1 - band.src_y0) * W * CH * np_dt.itemsize

            def sync():
                torch.cuda.current_stream().synchronize()
                dist.barrier()

            phases = os.environ.get("AAI_BENCH_PHASES")

            def e2e_step():
                t = [time.perf_counter()]
                peer.upload_owned(hsi, stream)
                if phases:
                    torch.cuda.current_stream().synchronize(); t.append(time.perf_counter())
                sync()                      # every owner's rows are on its device
                if phases:
                    t.append(time.perf_counter())
                moved = peer.pull_halo(stream)      # NVLink peer copies of the rows this band needs but does not own
                if phases:
                    torch.cuda.current_stream().synchronize(); t.append(time.perf_counter())
                aai.run_device(plan, si, di, band.row0, band.row1, arith=arith, device=local, stream=stream)
                aai.image_download(hdi, di, local, stream)
                if phases:
                    torch.cuda.current_stream().synchronize(); t.append(time.perf_counter())
                sync()                      # peers have finished reading before the next upload overwrites
                if phases:
                    t.append(time.perf_counter())
                    print(f"[phases] rank {rank}: upload {1e3*(t[1]-t[0]):.2f} barrier {1e3*(t[2]-t[1]):.2f} pull "
                          f"{1e3*(t[3]-t[2]):.2f} ({moved/1e6:.0f} MB) kernel+d2h {1e3*(t[4]-t[3]):.2f} barrier "
                          f"{1e3*(t[5]-t[4]):.2f} ms", file=sys.stderr)

            e2e_step()  # populates this rank's halo for the device-timed loop
            h2d = (o1 - o0) * W * CH * np_dt.itemsize
        else:
            halo_rows = band.src_y1 - band.src_y0
            host_src = torch.empty((max(halo_rows, 1), W) + shape_tail, dtype=t_dt, pin_memory=True)
            if halo_rows:
                synthetic_image(W, H, np_dt, seed, channels=CH, y0=band.src_y0, rows=halo_rows,
                                out=host_src.numpy()[:halo_rows])
            dev_src = host_src.to(dev)
            si = aai.tensor_image(dev_src[:halo_rows] if halo_rows else dev_src, y0=band.src_y0, height=H)
            hsi = aai.tensor_image(host_src[:halo_rows] if halo_rows else host_src, y0=band.src_y0, height=H)
            resident_bytes = dev_src.numel() * dev_src.element_size()

            def e2e_step():
                aai.run_host_band(plan, hsi, hdi, band.row0, band.row1, arith=arith, device=local, stream=stream,
                                  synchronize=False)

            h2d = halo_rows * W * CH * np_dt.itemsize

        def step():
            aai.run_device(plan, si, di, band.row0, band.row1, arith=arith, device=local, stream=stream)

    else:
        # batch of independent images: whole images per rank
        lo, hi = batch_slice(cfg["batch"], rank, world)
        n_img = hi - lo
        distinct = min(n_img, 4)  # the host keeps a few distinct synthetic slices; the device batch is resident
        host_src = torch.empty((distinct, H, W), dtype=t_dt, pin_memory=True)
        for k in range(distinct):
            synthetic_image(W, H, np_dt, seed + 1000 * (lo + k), out=host_src.numpy()[k])
        host_dst = torch.empty((distinct, plan.dst_h, plan.dst_w), dtype=out_dt, pin_memory=True)
        dev_src = torch.empty((n_img, H, W), dtype=t_dt, device=dev)
        for k in range(n_img):
            dev_src[k].copy_(host_src[k % distinct], non_blocking=True)
        dev_dst = torch.empty((n_img, plan.dst_h, plan.dst_w), dtype=out_dt, device=dev)
        sis = [aai.tensor_image(dev_src[k]) for k in range(n_img)]
        dis = [aai.tensor_image(dev_dst[k]) for k in range(n_img)]
        hsis = [aai.tensor_image(host_src[k]) for k in range(distinct)]
        hdis = [aai.tensor_image(host_dst[k]) for k in range(distinct)]
        my_pixels = n_img * plan.dst_w * plan.dst_h
        launches_per_step = 1  # the equally strided batch is one launch (rank-3 TMA tensor map)

        def step():
            aai.run_device_batch(plan, sis, dis, arith=arith, device=local, stream=stream)

        def e2e_step():
            for k in range(n_img):
                aai.run_host_band(plan, hsis[k % distinct], hdis[k % distinct], 0, plan.dst_h, arith=arith,
                                  device=local, stream=stream, synchronize=False)

        h2d = n_img * H * W * np_dt.itemsize
        d2h = n_img * plan.dst_w * plan.dst_h * host_dst.element_size()
        resident_bytes = dev_src.numel() * dev_src.element_size()
        band = None

    total_pixels = sum_over_ranks(float(my_pixels))

    # ---- device-timed region -------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    uuid = None
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        pass
    sampler = ClockSampler(uuid if rank == 0 else None)
    if rank == 0:
        sampler.start()
    # untimed pre-load so that the clock sampler sees the loaded state even when K steps are only milliseconds
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 0.6:
        step()
        torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    before = aai.launch_count()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = aai.launch_count() - before
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = total_pixels / (ms_step * 1e-3) / 1e6

    # ---- end-to-end region (host buffers through the C ABI) --------------------------------------------------
    for _ in range(2):
        e2e_step()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        e2e_step()
    e3.record()
    barrier()
    e2e_ms = max_over_ranks(e2.elapsed_time(e3)) / args.steps
    clocks = sampler.stop() if rank == 0 else {}
    h2d_total, d2h_total = sum_over_ranks(float(h2d)), sum_over_ranks(float(d2h))

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    if cfg["batch"] == 1:
        nz = (dev_dst if CH == 1 else dev_dst[..., 0]) != 0
        covered = sum_over_ranks(float(nz.sum().item()))
    else:
        covered = total_pixels
    alg_bytes = (W * H * CH * np_dt.itemsize + plan.dst_w * plan.dst_h * CH * host_dst.element_size()) * cfg["batch"]
    kernel_ms = ms_step
    if plan.axis_aligned:
        per_launch = alg_bytes / world
        achieved = per_launch / (kernel_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "traffic": None, "peak_source": hbm_src,
                    "algorithmic": "each source byte read once + each canvas byte written once (SURVEY 8d)"}
    else:
        F = flops_per_covered_pixel(plan.side, plan.cos_t, plan.sin_t, CH)
        nominal = n_sm * 128 * 2 * sm_max * 1e6 / 1e12
        try:  # measured on this device right now (register-only FFMA probe in the library)
            measured = aai.measure_fp32_tflops(local)
        except Exception as exc:
            print(f"[bench] FP32 probe failed: {exc}", file=sys.stderr)
            measured = 0.0
        fp32_peak = measured if 0.5 * nominal < measured < 1.1 * nominal else nominal
        achieved = F * covered / world / (kernel_ms * 1e-3) / 1e12
        roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp32_peak, "traffic": None,
                    "peak_source": (f"measured on this device by the library's FFMA probe ({measured:.2f} TFLOP/s; nominal "
                                    f"{n_sm} SMs x 128 lanes x 2 x {sm_max:.0f} MHz = {nominal:.2f}; MEASURED_PEAKS.json "
                                    "holds no FP32 figure)" if fp32_peak == measured else
                                    f"nominal {n_sm} SMs x 128 FP32 lanes x 2 x {sm_max:.0f} MHz (FFMA probe unavailable; "
                                    "MEASURED_PEAKS.json holds no FP32 figure)"),
                    "peak_nominal": nominal,
                    "algorithmic": f"{F:.0f} flop per covered canvas pixel x {covered / world:.0f} covered pixels per "
                                   "launch (SURVEY 8d contract figure)",
                    "hbm": {"achieved": alg_bytes / world / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "peak_source": hbm_src}}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "latest.json"))).get(f"cfg{args.config}", {})
        traffic = prof.get("dram_bytes_per_launch")
        if traffic is not None and "per" in prof:  # profiled per slice of a batched launch: scale to this rank's launch
            traffic = int(traffic * cfg["batch"] / world)
        roofline["traffic"] = traffic
    except Exception:
        pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if arith == aai.ARITH_F32 else "f64", "data": "synthetic",
        "config": {"workload": cfg["label"], "canvas": f"{plan.dst_w}x{plan.dst_h}", "scale": plan.scale,
                   "footprint_side": plan.side, "sharding": ("row bands balanced by covered pixels" if cfg["batch"] == 1
                                                             else "whole images per rank"),
                   "timing": f"inputs larger than L2 ({resident_bytes / 1e6:.0f} MB resident source per rank)"
                   if resident_bytes > 130e6 else "L2-resident input (source smaller than L2); latency bound",
                   "covered_pixels": covered, "numa_node_rank0": numa,
                   "e2e_source": ("each source row uploaded once by its owner rank, halos pulled over NVLink "
                                  "(CUDA IPC peer copies, no NCCL on the data path)" if peer is not None else
                                  "each rank uploads its band's source halo from pinned host memory (chunk-pipelined)")},
        "clocks": {k: clocks.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples")},
        "e2e": {"value": total_pixels / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d_total,
                "d2h_bytes_per_step": d2h_total, "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind = cpu_kind()
        rep = replica_of(cfg, cfg["baseline_replica"])
        n, sec = _ref_worker((rep["w"], rep["ratio"], rep["angle"], rep["iso"], seed, kind))
        line["cpu_baseline"] = {
            "value": n / sec / 1e6, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"one {rep['w']}x{rep['h']} float64 replica of the workload (same ratio/angle), {sec:.1f} s, "
                      "single thread as the reference ships"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if peer is not None:
        barrier()
        peer.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arith", default="auto", choices=["auto", "f64", "f32"],
                    help="auto: FP32 kernel for float32/8-bit images, FP64 kernel for float64 images")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-peer", action="store_true",
                    help="N>1: every rank uploads its whole halo from the host instead of NVLink peer copies")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    # stdout carries exactly ONE JSON line: libraries that write to file descriptor 1 (NCCL prints its version banner
    # there under torchrun) are sent to stderr; the JSON line is written to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(json_fd, "w")
    if args.impl == "reference":
        return reference_arm(args, cfg)
    return our_arm(args, cfg)


if __name__ == "__main__":
    sys.exit(main())
