"""Developer tool: per-band kernel times of the row-band partition on ONE GPU (bands run one after another),
to tune the partitioner's cost model.  python tools/dev_bands.py [--config 4] [--parts 8]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import area_average_interpolation_b200 as aai
from bench import CONFIGS

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--parts", type=int, nargs="+", default=[8])
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--json", default="", help="append per-band features + times as JSON lines (cost-model fit)")
ap.add_argument("--lib", default="", help="another build of libaai_b200.so (A/B variants)")
ap.add_argument("--arith", default="f32", choices=["f32", "f64"])
ap.add_argument("--mode", type=int, default=1, choices=[1, 2], help="1 area average, 2 fast mode")
args = ap.parse_args()
if args.lib:
    aai.LIB_PATH = args.lib
cfg = CONFIGS[args.config]
dev = torch.device("cuda:0")
plan = aai.make_plan(cfg["w"], cfg["h"], 1.0, cfg["ratio"], cfg["iso"], cfg["angle"])
tail = (cfg["ch"],) if cfg["ch"] > 1 else ()
if cfg["dtype"] == "uint8":
    src = torch.randint(0, 256, (cfg["h"], cfg["w"]) + tail, dtype=torch.uint8, device=dev)
else:
    src = torch.rand((cfg["h"], cfg["w"]) + tail, dtype=torch.float32, device=dev) * 4096
dst = torch.empty((plan.dst_h, plan.dst_w) + tail, dtype=torch.float32, device=dev)
si, di = aai.tensor_image(src), aai.tensor_image(dst)
st = torch.cuda.current_stream().cuda_stream
ARITH = aai.ARITH_F32 if args.arith == "f32" else aai.ARITH_F64
WEIGHT = aai.band_empty_weight(plan, args.mode, ARITH)


def timed(r0, r1):
    for _ in range(3):
        aai.run_device(plan, si, di, r0, r1, mode=args.mode, arith=ARITH, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        aai.run_device(plan, si, di, r0, r1, mode=args.mode, arith=ARITH, stream=st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps


def features(r0, r1):
    """Per-band cost-model inputs: rows, canvas pixels, covered pixels, warps (16 x 2 pixel tiles) that hold a covered
    pixel, span ends (border pixels sit there)."""
    import ctypes as C
    lib = aai.lib()
    cov = warps = ends = 0
    for y in range(r0 - r0 % 2, r1, 2):
        lo, hi = None, None
        for yy in (y, y + 1):
            if yy < r0 or yy >= r1:
                continue
            c = aai.covered_pixels(plan, yy, yy + 1)
            cov += c
            if c:
                ends += 2
        # span of the row pair from the plan's covered count is not exposed; approximate by the wider of the two rows
        c2 = max(aai.covered_pixels(plan, max(y, r0), max(y, r0) + 1), aai.covered_pixels(plan, min(y + 1, r1 - 1), min(y + 1, r1 - 1) + 1))
        warps += (c2 + 15) // 16 + (1 if c2 else 0)
    return dict(rows=r1 - r0, pixels=(r1 - r0) * plan.dst_w, covered=cov, warps=warps, ends=ends)


whole = timed(0, plan.dst_h)
print(f"whole canvas: {whole:.4f} ms")
for n in args.parts:
    b = aai.partition_rows(plan, n, WEIGHT)
    ts = [timed(b[k], b[k + 1]) for k in range(n)]
    cov = [aai.covered_pixels(plan, b[k], b[k + 1]) for k in range(n)]
    print(f"{n} bands: rows {[b[k + 1] - b[k] for k in range(n)]}")
    print(f"   covered Mpx {[round(c / 1e6, 2) for c in cov]}")
    print(f"   ms {[round(t, 4) for t in ts]}  max {max(ts):.4f}  sum {sum(ts):.4f}  ideal {whole / n:.4f}  "
          f"speed-up {whole / max(ts):.2f}x")
    if args.json:
        import json
        with open(args.json, "a") as f:
            for k in range(n):
                d = features(b[k], b[k + 1])
                d.update(ms=ts[k], parts=n, band=k, config=args.config)
                f.write(json.dumps(d) + "\n")
