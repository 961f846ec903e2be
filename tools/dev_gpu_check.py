"""Developer check run on the GPU box: parity of the CUDA path vs the oracle on a sweep + kernel timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import area_average_interpolation_b200 as aai
from oracle import port

def check(w, h, r, ang, iso, dtype=np.float64, ch=1, arith=0):
    rng = np.random.default_rng(7)
    shape = (h, w) if ch == 1 else (h, w, ch)
    src = rng.uniform(0, 255, size=shape)
    src = src.astype(dtype) if dtype != np.uint8 else np.floor(src).astype(np.uint8)
    op = aai.AreaAverageInterpolation(arith=arith, out_dtype=np.float64)
    got = op.areaAverageInterpolation(src, 1.0, r, iso, ang)
    worst = 0; nbad = 0
    for c in range(ch):
        st, want, wiso = port.run(src, 1.0, r, iso, ang, channel=c)
        g = got.dst if ch == 1 else got.dst[..., c]
        err = np.abs(g - want) / np.maximum(np.abs(want), 1.0)
        worst = max(worst, err.max()); nbad += int((err > 1e-9).sum())
    tol = 1e-9 if arith == 0 else 1e-5
    nbad = 0 if worst <= tol else 1
    print(f"{w}x{h} r={r} ang={ang} iso={iso} {np.dtype(dtype).name} ch={ch} arith={arith}: canvas {got.dst.shape} max rel err {worst:.3e} bad={nbad}", flush=True)
    return nbad

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), aai.device_count())
    bad = 0
    for a in [(64,64,0.5,0,(32,32)), (64,64,0.5,0,(31.5,31.5)), (64,64,0.37,30,(32,32)), (97,61,0.37,17.3,(48,30)),
              (50,70,0.37,117.3,(25,35)), (50,70,0.37,200,(25,35)), (50,70,0.9,305.5,(0,0)), (40,40,2.3,61,(20,20)),
              (33,47,0.2,-12,(100,-5)), (64,64,0.37,90,(32,32)), (32,32,1.0,0,(16,16)), (32,32,2.0,0,(16,16)),
              (48,48,1.7,45,(23.5,23.5)), (64,64,0.37,44.999,(32,32)), (64,64,0.37,45.001,(32,32)),
              (512,512,0.37,17.3,(256,256)), (300,200,0.11,73.0,(150,100)), (256,256,1.7,45,(127.5,127.5))]:
        bad += check(*a)
    for a in [(64,64,0.37,30,(32,32)), (97,61,0.37,17.3,(48,30)), (50,70,0.37,117.3,(25,35)), (50,70,0.9,305.5,(0,0)), (40,40,2.3,61,(20,20)),
              (512,512,0.37,17.3,(256,256)), (300,200,0.11,73.0,(150,100)), (256,256,1.7,45,(127.5,127.5)), (700,500,0.37,30,(350,250)), (128,128,0.6,1.0,(64,64)), (128,128,0.6,5.0,(64,64))]:
        bad += check(*a, dtype=np.float32, arith=1)
    bad += check(128,96,0.37,30,(64,48),np.uint8,3,arith=1)
    bad += check(128,96,0.37,30,(64,48),np.float32)
    bad += check(128,96,0.37,30,(64,48),np.uint8,3)
    print("TOTAL BAD", bad)
    # timing: cfg4 device-resident
    dev = torch.device("cuda:0")
    for (W, r, ang, iso, dt) in [(16384, 0.37, 17.3, (8192, 8192), torch.float32), (4096, 0.5, 0.0, (2048, 2048), torch.float32),
                                 (2048, 0.37, 30.0, (1024,1024), torch.uint8), (4096, 1.7, 45.0, (2047.5,2047.5), torch.uint8)]:
        plan = aai.make_plan(W, W, 1.0, r, iso, ang)
        if dt == torch.uint8:
            src = torch.randint(0, 256, (W, W), dtype=dt, device=dev)
        else:
            src = torch.rand(W, W, dtype=dt, device=dev) * 4096
        dst = torch.empty(plan.dst_h, plan.dst_w, dtype=torch.float32, device=dev)
        si, di = aai.tensor_image(src), aai.tensor_image(dst)
        st = torch.cuda.current_stream().cuda_stream
        for arith in (0, 1):
            for _ in range(2):
                aai.run_device(plan, si, di, arith=arith, stream=st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 5
            e0.record()
            for _ in range(n):
                aai.run_device(plan, si, di, arith=arith, stream=st)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            print(f"W={W} r={r} ang={ang} arith={arith}: canvas {plan.dst_w}x{plan.dst_h} kernel {ms:.3f} ms -> {plan.dst_w*plan.dst_h/ms/1e3:.1f} Mpix/s", flush=True)
