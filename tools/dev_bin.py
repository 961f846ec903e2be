"""Developer tool (GPU box): fast mode by source-side binning (aai_kernels_bin.cu) against the gather kernel and the
FP64 fast kernel -- holes (pixels nobody wrote), error against FP64, row bands / stacks bitwise, timing on config 4.
    python tools/dev_bin.py [--time] [--lib other.so]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import area_average_interpolation_b200 as aai

ap = argparse.ArgumentParser()
ap.add_argument("--time", action="store_true")
ap.add_argument("--lib", default="")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--sweep", action="store_true", help="time both kernels over ratios / angles on an 8192^2 source")
ap.add_argument("--skip-checks", action="store_true")
args = ap.parse_args()
if args.lib:
    aai.LIB_PATH = args.lib
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
SENT = -777.0


def run(plan, src, arith, dtype=torch.float32, rows=None, y0=None, height=None):
    dst = torch.full((plan.dst_h, plan.dst_w), SENT, dtype=torch.float32, device=dev).to(dtype)
    si = aai.tensor_image(src) if y0 is None else aai.tensor_image(src, y0=y0, height=height)
    if rows is None:
        aai.run_device(plan, si, aai.tensor_image(dst), mode=aai.MODE_FAST, arith=arith, stream=st)
    else:
        aai.run_device(plan, si, aai.tensor_image(dst), rows[0], rows[1], mode=aai.MODE_FAST, arith=arith, stream=st)
    torch.cuda.synchronize()
    return dst


CASES = [
    (1500, 1100, 0.37, 17.3, (750.0, 550.0)),
    (900, 700, 0.37, 30.0, (450.0, 350.0)),
    (640, 480, 0.6, 61.0, (320.0, 240.0)),
    (800, 600, 0.23, 40.0, (400.0, 300.0)),
    (500, 400, 0.45, 12.0, (10.0, 390.0)),
    (333, 517, 0.7, 45.0, (166.0, 258.0)),
    (257, 131, 0.3, 83.0, (128.0, 65.0)),
    (2048, 2048, 0.37, 17.3, (1024.0, 1024.0)),
]
bad = 0
for (w, h, ratio, angle, iso) in ([] if args.skip_checks else CASES):
    g = torch.Generator(device="cpu").manual_seed(w * 7 + h)
    src = (torch.rand((h, w), generator=g, dtype=torch.float32) * 4096).to(dev)
    plan = aai.make_plan(w, h, 1.0, ratio, iso, angle)
    a = run(plan, src, aai.ARITH_F32_BINNED)
    b = run(plan, src, aai.ARITH_F32)
    c = run(plan, src, aai.ARITH_F64)
    holes = int((a == SENT).sum())
    an, bn, cn = a.cpu().numpy().astype(np.float64), b.cpu().numpy().astype(np.float64), c.cpu().numpy().astype(np.float64)
    err = np.abs(an - cn) / np.maximum(np.abs(cn), 16.0)
    errg = np.abs(bn - cn) / np.maximum(np.abs(cn), 16.0)
    nbad = int((err > 1e-5).sum())
    line = (f"{w}x{h} r={ratio} a={angle}: canvas {plan.dst_w}x{plan.dst_h} holes {holes} max err bin {err.max():.2e} "
            f"(gather {errg.max():.2e}) bad {nbad} identical-to-gather {(an == bn).mean():.4f}")
    if nbad:
        ys, xs = np.nonzero(err > 1e-5)
        line += " first bad: " + ", ".join(f"({x},{y}) {an[y, x]:.3f} vs {cn[y, x]:.3f}" for y, x in list(zip(ys, xs))[:6])
    # u8 canvas
    a8 = run(plan, src / 16.0, aai.ARITH_F32_BINNED, dtype=torch.uint8)
    c8 = run(plan, src / 16.0, aai.ARITH_F64, dtype=torch.uint8)
    d8 = (a8.to(torch.int16) - c8.to(torch.int16)).abs()
    line += f" | u8 canvas: differing {int((d8 > 0).sum())} max {int(d8.max())}"
    # row bands with halo-only sources, bitwise
    from area_average_interpolation_b200.sharding import all_bands
    bands = torch.full_like(a, -2.0)
    for band in all_bands(plan, 3):
        halo = src[band.src_y0:band.src_y1].contiguous()
        aai.run_device(plan, aai.tensor_image(halo, y0=band.src_y0, height=h), aai.tensor_image(bands), band.row0, band.row1,
                       mode=aai.MODE_FAST, arith=aai.ARITH_F32_BINNED, stream=st)
    torch.cuda.synchronize()
    line += f" | bands bitwise {bool(torch.equal(bands, a))}"
    # chunks of 37 rows over the full source
    ch = torch.full_like(a, -3.0)
    for r0 in range(0, plan.dst_h, 37):
        aai.run_device(plan, aai.tensor_image(src), aai.tensor_image(ch), r0, min(r0 + 37, plan.dst_h), mode=aai.MODE_FAST,
                       arith=aai.ARITH_F32_BINNED, stream=st)
    torch.cuda.synchronize()
    line += f" chunks bitwise {bool(torch.equal(ch, a))}"
    # stack of 3 slices
    srcs = torch.stack([src, src.flip(0), src * 0.5])
    dsts = torch.full((3, plan.dst_h, plan.dst_w), -4.0, dtype=torch.float32, device=dev)
    aai.run_device_batch(plan, [aai.tensor_image(srcs[k]) for k in range(3)], [aai.tensor_image(dsts[k]) for k in range(3)],
                         mode=aai.MODE_FAST, arith=aai.ARITH_F32_BINNED, stream=st)
    torch.cuda.synchronize()
    s1 = run(plan, srcs[1].contiguous(), aai.ARITH_F32_BINNED)
    line += f" stack bitwise {bool(torch.equal(dsts[0], a) and torch.equal(dsts[1], s1))}"
    ring, staged = run(plan, src, aai.ARITH_F32_RING), run(plan, src, aai.ARITH_F32_STAGED)
    ring_ok = bool(torch.equal(ring, staged)) and int((ring == SENT).sum()) == 0
    line += f" | ring == staged {ring_ok}"
    print(line, flush=True)
    bad += nbad + holes + (0 if ring_ok else 1)

if args.time:
    w = h = 16384
    src = torch.rand((h, w), dtype=torch.float32, device=dev) * 4096
    plan = aai.make_plan(w, h, 1.0, 0.37, (8192.0, 8192.0), 17.3)
    dst = torch.empty((plan.dst_h, plan.dst_w), dtype=torch.float32, device=dev)
    si, di = aai.tensor_image(src), aai.tensor_image(dst)
    for name, arith in [("bin", aai.ARITH_F32_BINNED), ("gather", aai.ARITH_F32), ("staged", aai.ARITH_F32_STAGED),
                        ("ring", aai.ARITH_F32_RING), ("gather", aai.ARITH_F32), ("ring", aai.ARITH_F32_RING)]:
        for _ in range(3):
            aai.run_device(plan, si, di, mode=aai.MODE_FAST, arith=arith, stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            aai.run_device(plan, si, di, mode=aai.MODE_FAST, arith=arith, stream=st)
        e1.record()
        torch.cuda.synchronize()
        print(f"cfg4 fast mode, {name}: {e0.elapsed_time(e1) / args.steps:.4f} ms", flush=True)
    a = torch.empty_like(dst)
    aai.run_device(plan, si, aai.tensor_image(a), mode=aai.MODE_FAST, arith=aai.ARITH_F32_BINNED, stream=st)
    c = torch.empty_like(dst)
    aai.run_device(plan, si, aai.tensor_image(c), mode=aai.MODE_FAST, arith=aai.ARITH_F64, stream=st)
    torch.cuda.synchronize()
    err = ((a - c).abs() / torch.clamp(c.abs(), min=16.0)).max().item()
    print(f"cfg4 full canvas: max err vs FP64 fast kernel {err:.2e}")
if args.sweep:
    w = h = 8192
    src = torch.rand((h, w), dtype=torch.float32, device=dev) * 4096
    for ratio in (0.6, 0.45, 0.37, 0.33, 0.29, 0.27):
        for angle in (8.0, 17.3, 30.0, 45.0, 70.0):
            plan = aai.make_plan(w, h, 1.0, ratio, (4096.0, 4096.0), angle)
            dst = torch.empty((plan.dst_h, plan.dst_w), dtype=torch.float32, device=dev)
            si, di = aai.tensor_image(src), aai.tensor_image(dst)
            t = {}
            for name, arith in [("bin", aai.ARITH_F32_BINNED), ("gather", aai.ARITH_F32)]:
                for _ in range(2):
                    aai.run_device(plan, si, di, mode=aai.MODE_FAST, arith=arith, stream=st)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    aai.run_device(plan, si, di, mode=aai.MODE_FAST, arith=arith, stream=st)
                e1.record()
                torch.cuda.synchronize()
                t[name] = e0.elapsed_time(e1) / 10
            print(f"sweep 8192^2 ratio {ratio} angle {angle}: canvas {plan.dst_w}^2  bin {t['bin']:.4f} ms  gather {t['gather']:.4f} ms  "
                  f"bin/gather {t['bin'] / t['gather']:.2f}", flush=True)
print("DEV_BIN", "OK" if bad == 0 else f"FAILED ({bad})")
