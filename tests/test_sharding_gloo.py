"""CPU, world_size 2 over gloo: the one-process-per-GPU work assignment used by bench.py (no data-path collective).
Each rank evaluates ITS band with the CPU oracle standing in for the device; the stitched bands must equal the
whole image bit for bit, and the only collectives are the barrier / max-time / pixel-count reductions of the bench."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import area_average_interpolation_b200 as aai
    from area_average_interpolation_b200.sharding import band_for_rank, batch_slice
    from area_average_interpolation_b200.synthetic import synthetic_image
    from oracle import port as oracle

    w, h, ratio, angle, iso = 90, 70, 0.37, 17.3, (45.0, 35.0)
    plan = aai.make_plan(w, h, 1.0, ratio, iso, angle)
    band = band_for_rank(plan, rank, world)
    # each rank materialises ONLY its halo rows of the source (seekable generator), like a GPU rank would
    halo = synthetic_image(w, h, np.float64, 777, y0=band.src_y0, rows=band.src_y1 - band.src_y0)
    src = np.zeros((h, w))
    src[band.src_y0:band.src_y1] = halo
    st, part, _ = oracle.run(src, 1.0, ratio, iso, angle, rows=(band.row0, band.row1))
    assert st == 0
    np.save(os.path.join(tmpdir, f"band{rank}.npy"), part)
    # bench-style reductions
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    px = torch.tensor([band.rows * plan.dst_w], dtype=torch.int64)
    dist.all_reduce(px, op=dist.ReduceOp.SUM)
    dist.barrier()
    assert t.item() == float(world)
    assert px.item() == plan.dst_w * plan.dst_h
    # batch sharding
    lo, hi = batch_slice(7, rank, world)
    cnt = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    assert cnt.item() == 7
    dist.destroy_process_group()


def test_two_rank_bands_stitch_to_the_whole_image(built, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    import area_average_interpolation_b200 as aai
    from area_average_interpolation_b200.synthetic import synthetic_image
    from oracle import port as oracle

    w, h, ratio, angle, iso = 90, 70, 0.37, 17.3, (45.0, 35.0)
    src = synthetic_image(w, h, np.float64, 777)
    st, full, _ = oracle.run(src, 1.0, ratio, iso, angle)
    stitched = np.concatenate([np.load(tmp_path / f"band{r}.npy") for r in range(world)], axis=0)
    assert stitched.shape == full.shape
    assert np.array_equal(stitched, full)


def test_batch_slices_partition_the_batch():
    sys.path.insert(0, ROOT)
    from area_average_interpolation_b200.sharding import batch_slice

    for n in (1, 7, 256):
        for world in (1, 2, 3, 8):
            parts = [batch_slice(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
