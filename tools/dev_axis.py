"""Developer tool: device time of the axis-aligned kernels on one 4096^2 slice -- TMA kernel (0 deg, 1 channel) against the
FP32 direct-tap kernel (quadrant pre-rotation / RGB / expansion) and the FP64 direct-tap kernel it replaced.
    python tools/dev_axis.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import area_average_interpolation_b200 as aai

dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream


def timed(plan, src, dst, arith, steps=30):
    si, di = aai.tensor_image(src), aai.tensor_image(dst)
    for _ in range(3):
        aai.run_device(plan, si, di, arith=arith, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        aai.run_device(plan, si, di, arith=arith, stream=st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


W = 4096
for label, ratio, angle, dtype, ch in [("0 deg f32 (TMA kernel)", 0.5, 0.0, torch.float32, 1),
                                       ("90 deg f32", 0.5, 90.0, torch.float32, 1),
                                       ("180 deg f32", 0.5, 180.0, torch.float32, 1),
                                       ("0 deg RGB u8", 0.5, 0.0, torch.uint8, 3),
                                       ("270 deg RGB u8", 0.5, 270.0, torch.uint8, 3),
                                       ("0 deg f32 scale 2 (ratio 1.0)", 1.0, 0.0, torch.float32, 1),
                                       ("0 deg f32 0.37x (4 taps)", 0.37, 0.0, torch.float32, 1),
                                       ("90 deg f32 0.37x (4 taps)", 0.37, 90.0, torch.float32, 1)]:
    plan = aai.make_plan(W, W, 1.0, ratio, (W / 2, W / 2), angle)
    tail = (ch,) if ch > 1 else ()
    src = (torch.randint(0, 256, (W, W) + tail, dtype=torch.uint8, device=dev) if dtype == torch.uint8
           else torch.rand((W, W) + tail, dtype=torch.float32, device=dev) * 4096)
    dst = torch.empty((plan.dst_h, plan.dst_w) + tail, dtype=dtype, device=dev)
    t32 = timed(plan, src, dst, aai.ARITH_F32)
    dst64 = torch.empty((plan.dst_h, plan.dst_w) + tail, dtype=torch.float32, device=dev)
    t64 = timed(plan, src, dst64, aai.ARITH_F64)
    byts = src.numel() * src.element_size() + dst.numel() * dst.element_size()
    print(f"{label:34s} canvas {plan.dst_w}x{plan.dst_h}: FP32 path {t32*1e3:8.1f} us ({byts/t32/1e6:7.0f} GB/s algorithmic), "
          f"FP64 path {t64*1e3:8.1f} us")
