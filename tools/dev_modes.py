"""Developer tool: device-timed kernels of every interpolation mode / arithmetic on one BASELINE config.
    python tools/dev_modes.py [--config 4]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import area_average_interpolation_b200 as aai
from bench import CONFIGS

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
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
for mode, mname in [(1, "area average"), (2, "fast"), (3, "exact")]:
    for arith, aname in [(aai.ARITH_F32, "f32"), (aai.ARITH_F64, "f64")]:
        for _ in range(2):
            aai.run_device(plan, si, di, mode=mode, arith=arith, stream=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            aai.run_device(plan, si, di, mode=mode, arith=arith, stream=st)
        e1.record()
        torch.cuda.synchronize()
        print(f"{cfg['label']}: mode {mode} ({mname}), {aname}: {e0.elapsed_time(e1) / args.steps:.3f} ms")
