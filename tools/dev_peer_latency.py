"""Developer tool: step latency of the peer group on a SMALL image (copies take microseconds, so the step time is the
cost of the cross-process synchronisation: interprocess CUDA events + the host counters).  Ranks share the visible GPUs
(two processes on one device work: CUDA IPC is per process).
    python tools/dev_peer_latency.py [--world 2] [--side 512] [--steps 200]"""
import argparse
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rank_main(rank, world, side, steps, tmpdir):
    import numpy as np
    import torch

    import area_average_interpolation_b200 as aai
    from test_peer_group_gpu import _file_gather

    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    stream = torch.cuda.current_stream().cuda_stream
    plan = aai.make_plan(side, side, 1.0, 0.37, (side / 2, side / 2), 17.3)
    g = aai.PeerGroup(plan, aai.F32, 1, rank, world, dev, _file_gather(tmpdir, "lat", rank, world))
    o0, o1 = g.owned_rows()
    r0, r1 = g.band()
    src = torch.rand((max(o1 - o0, 1), side), dtype=torch.float32).pin_memory()
    dst = torch.empty((max(r1 - r0, 1), plan.dst_w), dtype=torch.float32).pin_memory()
    si = aai.tensor_image(src[:o1 - o0], y0=o0, height=side)
    di = aai.tensor_image(dst[:r1 - r0], y0=r0, height=plan.dst_h)
    for sync in (False, True):
        for _ in range(10):
            g.run(si, di, arith=aai.ARITH_F32, stream=stream, synchronize=sync)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            g.run(si, di, arith=aai.ARITH_F32, stream=stream, synchronize=sync)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps * 1e3
        print(f"rank {rank}/{world} side {side} synchronize={sync}: {dt:.3f} ms per step, phases {g.last_timing()}", flush=True)
    open(os.path.join(tmpdir, f"done{rank}"), "w").close()
    while not all(os.path.exists(os.path.join(tmpdir, f"done{p}")) for p in range(world)):
        time.sleep(0.01)
    g.close()


if __name__ == "__main__":
    import torch.multiprocessing as mp

    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--side", type=int, default=512)
    ap.add_argument("--steps", type=int, default=200)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(rank_main, args=(a.world, a.side, a.steps, d), nprocs=a.world, join=True)
