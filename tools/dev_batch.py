"""Developer timing: a stack of slices under ONE rotated plan, per-slice launches vs one batched launch
(aai_run_device_batch, grid.z = slice).  python tools/dev_batch.py [--side 512] [--n 64]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import area_average_interpolation_b200 as aai

ap = argparse.ArgumentParser()
ap.add_argument("--side", type=int, nargs="*", default=[256, 512, 2048])
ap.add_argument("--n", type=int, default=64)
ap.add_argument("--ratio", type=float, default=0.37)
ap.add_argument("--angle", type=float, default=17.3)
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
st = torch.cuda.current_stream().cuda_stream
for side in args.side:
    plan = aai.make_plan(side, side, 1.0, args.ratio, (side / 2.0, side / 2.0), args.angle)
    src = torch.rand((args.n, side, side), dtype=torch.float32, device="cuda") * 4096
    dst = torch.empty((args.n, plan.dst_h, plan.dst_w), dtype=torch.float32, device="cuda")
    ref = torch.empty_like(dst)
    sis = [aai.tensor_image(src[k]) for k in range(args.n)]
    dis = [aai.tensor_image(dst[k]) for k in range(args.n)]
    ris = [aai.tensor_image(ref[k]) for k in range(args.n)]

    def per_slice():
        for k in range(args.n):
            aai.run_device(plan, sis[k], ris[k], arith=aai.ARITH_F32, stream=st)

    def batched():
        aai.run_device_batch(plan, sis, dis, arith=aai.ARITH_F32, stream=st)

    out = []
    for f in (per_slice, batched):
        for _ in range(3):
            f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            f()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / args.steps)
    same = torch.equal(dst, ref)
    mpix = args.n * plan.dst_w * plan.dst_h / 1e6
    print(f"{args.n} x {side}^2 f32, {args.ratio}x, {args.angle} deg -> {plan.dst_w}x{plan.dst_h}: per-slice launches "
          f"{out[0]:.3f} ms ({mpix / out[0] * 1e3:.0f} Mpix/s), one batched launch {out[1]:.3f} ms "
          f"({mpix / out[1] * 1e3:.0f} Mpix/s), bitwise identical: {same}")
