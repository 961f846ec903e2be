"""Small target for ncu: the dominant kernel of one BASELINE config, device-resident, a few launches.
    python tools/profile_target.py --config 4 --steps 3 [--arith f64|f32]
(random device data instead of the seeded host generator: the profile does not depend on pixel values)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import area_average_interpolation_b200 as aai
from bench import CONFIGS

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--arith", default="f64")
ap.add_argument("--mode", type=int, default=1, help="1 area average, 2 fast mode")
ap.add_argument("--dst", default="same", help="canvas element type: same (as the source) | float32")
ap.add_argument("--angle", type=float, default=None, help="override the rotation angle of the config")
ap.add_argument("--ratio", type=float, default=None, help="override the ratio of the config")
ap.add_argument("--lib", default="", help="another build of libaai_b200.so (A/B variants)")
ap.add_argument("--batch", type=int, default=0, help="config 5: slices per launch (aai_run_device_batch)")
args = ap.parse_args()
cfg = CONFIGS[args.config]
if args.lib:
    aai.LIB_PATH = args.lib
dev = torch.device("cuda:0")
plan = aai.make_plan(cfg["w"], cfg["h"], 1.0, cfg["ratio"] if args.ratio is None else args.ratio, cfg["iso"],
                     cfg["angle"] if args.angle is None else args.angle)
tail = (cfg["ch"],) if cfg["ch"] > 1 else ()
if cfg["dtype"] == "uint8":
    src = torch.randint(0, 256, (cfg["h"], cfg["w"]) + tail, dtype=torch.uint8, device=dev)
else:
    src = torch.rand((cfg["h"], cfg["w"]) + tail, dtype=torch.float32, device=dev) * 4096
dst = torch.empty((plan.dst_h, plan.dst_w) + tail, dtype=src.dtype if args.dst == "same" else torch.float32, device=dev)
si, di = aai.tensor_image(src), aai.tensor_image(dst)
arith = {"f32": aai.ARITH_F32, "f32s": aai.ARITH_F32_STAGED, "f32b": aai.ARITH_F32_BINNED, "f32r": aai.ARITH_F32_RING}.get(args.arith, aai.ARITH_F64)
st = torch.cuda.current_stream().cuda_stream
if args.batch:
    srcs = torch.rand((args.batch, cfg["h"], cfg["w"]), dtype=torch.float32, device=dev) * 4096
    dsts = torch.empty((args.batch, plan.dst_h, plan.dst_w), dtype=torch.float32, device=dev)
    sis = [aai.tensor_image(srcs[i]) for i in range(args.batch)]
    dis = [aai.tensor_image(dsts[i]) for i in range(args.batch)]

    def run_device(plan, si, di, arith, stream):
        aai.run_device_batch(plan, sis, dis, arith=arith, stream=stream)

    aai_run = run_device
else:
    def aai_run(plan, si, di, arith, stream):
        aai.run_device(plan, si, di, mode=args.mode, arith=arith, stream=stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
aai_run(plan, si, di, arith=arith, stream=st)
torch.cuda.synchronize()
e0.record()
for _ in range(args.steps):
    aai_run(plan, si, di, arith=arith, stream=st)
e1.record()
torch.cuda.synchronize()
print(f"{cfg['label']}: canvas {plan.dst_w}x{plan.dst_h}, {e0.elapsed_time(e1) / args.steps:.3f} ms per launch")
