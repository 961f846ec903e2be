"""GPU (-m gpu): the peer group of the C ABI (aai_peer_*) -- ONE host image resampled by several processes, every source
row uploaded once by its owner, halos pulled over NVLink / peer copies on interprocess events -- and the host-buffer batch
entry point.  The multi-process cases run their ranks on as many devices as the box has (two processes share cuda:0 on the
1-GPU test box: CUDA IPC works between processes on one device), so the protocol is exercised everywhere."""
import os
import sys
import time

import numpy as np
import pytest

from common import TOL_F32_REL, f32_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    "rot_f32": dict(w=1500, h=1200, ratio=0.37, angle=17.3, iso=(750.0, 600.0), dtype="float32", ch=1, arith=1),
    "rgb_up_u8": dict(w=400, h=300, ratio=1.7, angle=45.0, iso=(199.5, 149.5), dtype="uint8", ch=3, arith=1),
    "quad1_f64": dict(w=700, h=900, ratio=0.6, angle=117.0, iso=(350.0, 450.0), dtype="float64", ch=1, arith=0),
    "axis_f32": dict(w=1024, h=2048, ratio=0.5, angle=0.0, iso=(512.0, 1024.0), dtype="float32", ch=1, arith=1),
}
STEPS = 3


def _source(case, step):
    from area_average_interpolation_b200.synthetic import synthetic_image

    return synthetic_image(case["w"], case["h"], np.dtype(case["dtype"]), 4242 + 17 * step, channels=case["ch"])


def _file_gather(tmpdir, name, rank, world):
    """Out-of-band exchange through files (the peer group needs no torch.distributed / NCCL / MPI)."""

    def gather(blob):
        with open(os.path.join(tmpdir, f"{name}.{rank}.tmp"), "wb") as f:
            f.write(blob)
        os.replace(os.path.join(tmpdir, f"{name}.{rank}.tmp"), os.path.join(tmpdir, f"{name}.{rank}.blob"))
        out = []
        for p in range(world):
            path = os.path.join(tmpdir, f"{name}.{p}.blob")
            t0 = time.time()
            while not os.path.exists(path):
                if time.time() - t0 > 60:
                    raise TimeoutError(path)
                time.sleep(0.005)
            out.append(open(path, "rb").read())
        return out

    return gather


def _rank_main(rank, world, tmpdir, names):
    sys.path.insert(0, ROOT)
    import torch

    import area_average_interpolation_b200 as aai

    device = rank % max(1, torch.cuda.device_count())
    torch.cuda.set_device(device)
    stream = torch.cuda.current_stream().cuda_stream
    for name in names:
        case = CASES[name]
        np_dt = np.dtype(case["dtype"])
        plan = aai.make_plan(case["w"], case["h"], 1.0, case["ratio"], case["iso"], case["angle"])
        group = aai.PeerGroup(plan, aai._NP_TO_AAI[np_dt], case["ch"], rank, world, device,
                              _file_gather(tmpdir, name, rank, world))
        o0, o1 = group.owned_rows()
        r0, r1 = group.band()
        tail = (case["ch"],) if case["ch"] > 1 else ()
        t_dt = {"uint8": torch.uint8, "float32": torch.float32, "float64": torch.float64}[case["dtype"]]
        # one pinned buffer pair per step: the steps are enqueued back to back WITHOUT synchronising in between, so the
        # cross-rank ordering (no owner overwrites rows a peer is still pulling) is what keeps the results right
        srcs = [torch.empty((max(o1 - o0, 1), case["w"]) + tail, dtype=t_dt, pin_memory=True) for _ in range(STEPS)]
        dsts = [torch.full((max(r1 - r0, 1), plan.dst_w) + tail, 0, dtype=t_dt, pin_memory=True) for _ in range(STEPS)]
        for t in range(STEPS):
            srcs[t][:o1 - o0].copy_(torch.from_numpy(_source(case, t)[o0:o1]))
        for t in range(STEPS):
            group.run(aai.tensor_image(srcs[t][:o1 - o0], y0=o0, height=case["h"]),
                      aai.tensor_image(dsts[t][:r1 - r0], y0=r0, height=plan.dst_h), arith=case["arith"], stream=stream,
                      synchronize=False)
        torch.cuda.synchronize()
        for t in range(STEPS):
            np.save(os.path.join(tmpdir, f"{name}.band{rank}.step{t}.npy"), dsts[t][:r1 - r0].numpy())
        # everyone has finished (files as the barrier) before anyone frees the memory its peers read
        open(os.path.join(tmpdir, f"{name}.{rank}.done"), "w").close()
        t0 = time.time()
        while not all(os.path.exists(os.path.join(tmpdir, f"{name}.{p}.done")) for p in range(world)):
            assert time.time() - t0 < 60
            time.sleep(0.005)
        group.close()


@pytest.fixture(scope="module")
def aai(built):
    import area_average_interpolation_b200 as m

    assert m.device_count() >= 1
    return m


def _expected(aai, case, step):
    src = _source(case, step)
    plan = aai.make_plan(case["w"], case["h"], 1.0, case["ratio"], case["iso"], case["angle"])
    dst = np.empty((plan.dst_h, plan.dst_w) + src.shape[2:], dtype=src.dtype)
    aai.run_host(plan, src, dst, arith=case["arith"])
    return plan, src, dst


@pytest.mark.parametrize("world", [1, 2, 3])
def test_peer_group_reproduces_one_gpu_bitwise(aai, built, tmp_path, world):
    """N processes (on N devices when the box has them, else sharing cuda:0), three steps enqueued back to back with
    different images: every step's stitched bands equal the single-process result bit for bit."""
    import torch.multiprocessing as mp

    names = list(CASES) if world == 2 else ["rot_f32", "quad1_f64"]
    mp.spawn(_rank_main, args=(world, str(tmp_path), names), nprocs=world, join=True)
    for name in names:
        case = CASES[name]
        for t in range(STEPS):
            plan, src, want = _expected(aai, case, t)
            parts = [np.load(tmp_path / f"{name}.band{r}.step{t}.npy") for r in range(world)]
            bounds = aai.partition_rows(plan, world)
            got = np.concatenate([parts[r][:bounds[r + 1] - bounds[r]] for r in range(world)], axis=0)
            assert got.shape == want.shape, (name, t)
            assert np.array_equal(got, want), (name, t, world)


def test_host_batch_pipeline_matches_per_image_runs(aai, built):
    """aai_run_host_batch: a volume of host slices through the ring of batched launches == one aai_run_host per slice,
    bit for bit (axis-aligned TMA path and a rotated plan), and within tolerance of the oracle."""
    import torch
    from oracle import port

    rng = np.random.default_rng(11)
    for (w, h, ratio, angle, iso, n) in [(640, 512, 0.5, 0.0, (320.0, 256.0), 23), (300, 260, 0.37, 17.3, (150.0, 130.0), 7)]:
        plan = aai.make_plan(w, h, 1.0, ratio, iso, angle)
        src = torch.from_numpy(rng.uniform(0, 4096, size=(n, h, w)).astype(np.float32)).pin_memory()
        dst = torch.full((n, plan.dst_h, plan.dst_w), -1.0, dtype=torch.float32).pin_memory()
        before = aai.launch_count()
        aai.run_host_batch(plan, [aai.tensor_image(src[k]) for k in range(n)], [aai.tensor_image(dst[k]) for k in range(n)],
                           arith=aai.ARITH_F32)
        assert before < aai.launch_count() <= before + n
        for k in (0, n // 2, n - 1):
            one = np.empty((plan.dst_h, plan.dst_w), dtype=np.float32)
            aai.run_host(plan, src[k].numpy(), one, arith=aai.ARITH_F32)
            # (the single-image host path may take the other axis-aligned kernel for its pitched buffer: rounding only)
            assert np.allclose(dst[k].numpy(), one, rtol=2e-6, atol=1e-5), k
            st, want, _ = port.run(src[k].numpy(), 1.0, ratio, iso, angle)
            assert f32_err(dst[k].numpy(), want, 4096.0).max() <= TOL_F32_REL
