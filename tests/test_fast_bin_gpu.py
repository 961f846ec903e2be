"""GPU (-m gpu): fast mode by source-side BINNING (aai_kernels_bin.cu; fastAreaAverageInterpolation, Source.cpp:584-911,
main loop 866-907) -- float single-channel images in the source frame (scale 1, quadrant 0) under AAI_ARITH_F32_BINNED (opt-in: measured slower than the default gather kernel, profiles/README.md).

The kernel turns the reference's loop around (every source pixel finds the one footprint it lies in), so these tests
check what that design could get wrong: pixels nobody writes, pixels on tile / band / image borders, the FP32 bin
decision (guard band -> FP64), reproducibility across band partitions, and the value itself against the CPU oracle and
against the default canvas-side gather kernel (AAI_ARITH_F32), which makes the same decisions in a different order."""
import numpy as np
import pytest

from common import TOL_F32_REL, f32_err

pytestmark = pytest.mark.gpu

SENTINEL = -777.0


@pytest.fixture(scope="module")
def aai(built):
    import area_average_interpolation_b200 as m

    assert m.device_count() >= 1, "no CUDA device: the product has no CPU fallback"
    return m


@pytest.fixture(scope="module")
def oracle(built):
    from oracle import port

    return port


def _fast(aai, plan, src, arith, dtype=None, rows=None, image=None):
    import torch

    dst = torch.full((plan.dst_h, plan.dst_w), SENTINEL, dtype=torch.float32, device="cuda")
    if dtype is not None:
        dst = dst.to(dtype)
    si = image if image is not None else aai.tensor_image(src)
    r0, r1 = rows if rows else (0, plan.dst_h)
    aai.run_device(plan, si, aai.tensor_image(dst), r0, r1, mode=aai.MODE_FAST, arith=arith,
                   stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return dst


# (w, h, ratio, angle, isocentre): the binning kernel's preconditions hold for all of them (float, 1 channel, scale 1,
# first quadrant, 3 deg away from the axes, L (cos + sin) < 5)
CASES = [
    (1500, 1100, 0.37, 17.3, (750.0, 550.0)),    # BASELINE config 4's ratio / angle: 4 slots, sheared box
    (900, 700, 0.37, 30.0, (450.0, 350.0)),      # config 2's
    (640, 480, 0.6, 61.0, (320.0, 240.0)),       # theta > 45 deg: no shear
    (800, 600, 0.29, 40.0, (400.0, 300.0)),      # L = 3.45, 5 lattice points per axis: 8 slots
    (500, 400, 0.45, 12.0, (10.0, 390.0)),       # isocentre near a corner: most of the canvas is empty
    (333, 517, 0.7, 45.0, (166.0, 258.0)),       # 45 deg
    (257, 131, 0.3, 83.0, (128.0, 65.0)),        # steep, image smaller than two tiles
    (130, 70, 0.5, 4.0, (64.3, 35.7)),           # near the axis limit, one tile
    (2048, 2048, 0.37, 17.3, (1024.0, 1024.0)),  # config 4 replica: many tiles, every border case
    (1237, 911, 0.41, 72.5, (700.25, 400.75)),   # odd sizes, fractional isocentre
]


@pytest.mark.parametrize("w,h,ratio,angle,iso", CASES)
def test_binning_kernel_matches_oracle_and_gather_kernel(aai, oracle, w, h, ratio, angle, iso):
    import torch

    rng = np.random.default_rng(w * 131 + h)
    host = rng.uniform(0.0, 4096.0, size=(h, w)).astype(np.float32)
    src = torch.from_numpy(host).cuda()
    plan = aai.make_plan(w, h, 1.0, ratio, iso, angle)
    got = _fast(aai, plan, src, aai.ARITH_F32_BINNED)
    assert int((got == SENTINEL).sum()) == 0, "canvas pixels that no kernel wrote"
    gather = _fast(aai, plan, src, aai.ARITH_F32)
    g, ga = got.cpu().numpy().astype(np.float64), gather.cpu().numpy().astype(np.float64)
    st, want, _ = oracle.run(host, 1.0, ratio, iso, angle, mode=2)
    assert st == 0 and want.shape == g.shape
    err = f32_err(g, want, 4096.0)
    bad = err > TOL_F32_REL
    if bad.any():  # centre-on-edge ties are decided by the reference's own rounding noise (T5 mask, as in test_gpu_parity)
        from test_gpu_parity import _conditioning_mask

        mask = _conditioning_mask(oracle, host, dict(src_res=1.0, dst_res=ratio, iso=iso, angle=angle, mode=2), want)
        assert not (bad & ~mask).any(), (int((bad & ~mask).sum()), float(err[~mask].max()))
        assert bad.mean() < 0.005
    # same inside decisions as the gather kernel: the two differ by the FP32 summation order only
    assert f32_err(g, ga, 4096.0).max() <= 2e-6
    # empty canvas pixels are exact zeros in both
    assert np.array_equal(g == 0.0, ga == 0.0) or np.abs(g - ga)[(g == 0.0) != (ga == 0.0)].max() == 0.0


@pytest.mark.parametrize("w,h,ratio,angle,iso", [CASES[0], CASES[3], CASES[6], CASES[9]])
def test_binning_kernel_is_bitwise_reproducible_across_bands_chunks_and_stacks(aai, w, h, ratio, angle, iso):
    """The run sums are keyed by (canvas pixel, source column mod NS) and added in fixed order, so the result does not
    depend on how the canvas is cut into bands (multi-GPU), chunks (host pipeline) or which slice of a stack it is."""
    import torch

    from area_average_interpolation_b200.sharding import all_bands

    rng = np.random.default_rng(w + h)
    src = torch.from_numpy(rng.uniform(0.0, 4096.0, size=(3, h, w)).astype(np.float32)).cuda()
    plan = aai.make_plan(w, h, 1.0, ratio, iso, angle)
    stream = torch.cuda.current_stream().cuda_stream
    whole = _fast(aai, plan, src[0], aai.ARITH_F32_BINNED)
    for parts in (2, 3, 8):
        bands = torch.full_like(whole, -2.0)
        for band in all_bands(plan, parts):
            halo = src[0, band.src_y0:band.src_y1].contiguous()  # each band sees only its halo rows
            aai.run_device(plan, aai.tensor_image(halo, y0=band.src_y0, height=h), aai.tensor_image(bands), band.row0,
                           band.row1, mode=aai.MODE_FAST, arith=aai.ARITH_F32_BINNED, stream=stream)
        torch.cuda.synchronize()
        assert torch.equal(bands, whole), parts
    chunks = torch.full_like(whole, -3.0)
    for r0 in range(0, plan.dst_h, 37):
        aai.run_device(plan, aai.tensor_image(src[0]), aai.tensor_image(chunks), r0, min(r0 + 37, plan.dst_h),
                       mode=aai.MODE_FAST, arith=aai.ARITH_F32_BINNED, stream=stream)
    torch.cuda.synchronize()
    assert torch.equal(chunks, whole)
    stack = torch.full((3, plan.dst_h, plan.dst_w), -4.0, dtype=torch.float32, device="cuda")
    aai.run_device_batch(plan, [aai.tensor_image(src[k]) for k in range(3)], [aai.tensor_image(stack[k]) for k in range(3)],
                         mode=aai.MODE_FAST, arith=aai.ARITH_F32_BINNED, stream=stream)
    torch.cuda.synchronize()
    assert torch.equal(stack[0], whole)
    assert torch.equal(stack[2], _fast(aai, plan, src[2].contiguous(), aai.ARITH_F32_BINNED))


def test_binning_kernel_random_plans_agree_with_gather_kernel(aai):
    """Seeded random geometry (ratio, angle, isocentre, odd sizes): box sizes, shear ranges, tile / image border cases the
    hand-picked shapes may miss.  The gather kernel (checked against the oracle elsewhere) is the reference here: same
    inside decisions, so the two may differ by FP32 summation order only, and no pixel may be left unwritten."""
    import torch

    rng = np.random.default_rng(20261018)
    stream = torch.cuda.current_stream().cuda_stream
    worst = 0.0
    for case in range(40):
        w, h = int(rng.integers(40, 900)), int(rng.integers(40, 700))
        ratio = float(rng.uniform(0.26, 0.7))
        angle = float(rng.uniform(3.5, 86.5))
        iso = (float(rng.uniform(-20, w + 20)), float(rng.uniform(-20, h + 20)))
        src = torch.from_numpy(rng.uniform(0.0, 4096.0, size=(h, w)).astype(np.float32)).cuda()
        plan = aai.make_plan(w, h, 1.0, ratio, iso, angle)
        if plan.dst_w * plan.dst_h > 4_000_000 or plan.dst_w * plan.dst_h == 0:
            continue
        got = _fast(aai, plan, src, aai.ARITH_F32_BINNED)
        ref = _fast(aai, plan, src, aai.ARITH_F32)
        assert int((got == SENTINEL).sum()) == 0, (case, w, h, ratio, angle, iso)
        err = float(f32_err(got.cpu().numpy(), ref.cpu().numpy(), 4096.0).max())
        assert err <= 2e-6, (case, w, h, ratio, angle, iso, err)
        worst = max(worst, err)
        if case % 8 == 0:  # two row bands over the whole source: bitwise the single launch
            bands = torch.full_like(got, -2.0)
            mid = plan.dst_h // 2
            for r0, r1 in ((0, mid), (mid, plan.dst_h)):
                aai.run_device(plan, aai.tensor_image(src), aai.tensor_image(bands), r0, r1, mode=aai.MODE_FAST,
                               arith=aai.ARITH_F32_BINNED, stream=stream)
            torch.cuda.synchronize()
            assert torch.equal(bands, got), (case, w, h, ratio, angle, iso)


def test_binning_kernel_8bit_canvas_and_padded_views(aai, oracle):
    """float source -> 8-bit canvas (round half up), and a source that is a column view of a wider array (pitch not a
    multiple of 16 bytes: the kernel only needs 4-byte alignment)."""
    import torch

    w, h, ratio, angle, iso = 700, 500, 0.37, 17.3, (350.0, 250.0)
    rng = np.random.default_rng(5)
    wide = torch.from_numpy(rng.uniform(0.0, 255.0, size=(h, w + 3)).astype(np.float32)).cuda()
    src = wide[:, 1:w + 1]  # not contiguous, starts 4 bytes into each row
    plan = aai.make_plan(w, h, 1.0, ratio, iso, angle)
    got = _fast(aai, plan, None, aai.ARITH_F32_BINNED, dtype=torch.uint8, image=aai.tensor_image(src))
    st, want, _ = oracle.run(src.cpu().numpy(), 1.0, ratio, iso, angle, mode=2)
    want8 = np.clip(np.floor(want + 0.5), 0, 255)
    diff = np.abs(got.cpu().numpy().astype(np.float64) - want8)
    assert diff.max() <= 1.0  # a value within FP32 rounding of x.5 may round the other way
    near_half = np.abs(want + 0.5 - np.round(want + 0.5)) < 1e-3
    assert not (diff > 0)[~near_half].any()


def test_binning_kernel_full_size_config4_against_fp64_kernel_and_oracle_rows(aai, oracle):
    """BASELINE config 4 at full size (16384^2 float32 -> 7591^2): whole canvas against the FP64 fast kernel, sampled
    rows against the CPU oracle."""
    import torch

    from area_average_interpolation_b200.synthetic import synthetic_image_torch

    w = h = 16384
    ratio, angle, iso = 0.37, 17.3, (8192.0, 8192.0)
    src = synthetic_image_torch(w, h, "float32", 20201, device=torch.device("cuda"))
    plan = aai.make_plan(w, h, 1.0, ratio, iso, angle)
    got = _fast(aai, plan, src, aai.ARITH_F32_BINNED)
    assert int((got == SENTINEL).sum()) == 0
    ref64 = _fast(aai, plan, src, aai.ARITH_F64)
    data_max = float(src.max())
    err = (got - ref64).abs() / torch.clamp(ref64.abs(), min=data_max / 256.0)
    assert float(err.max()) <= TOL_F32_REL
    host = src.cpu().numpy()
    for y in (0, 1234, plan.dst_h // 2, plan.dst_h - 1):
        st, want, _ = oracle.run(host, 1.0, ratio, iso, angle, mode=2, rows=(y, y + 1))
        assert f32_err(got[y:y + 1].cpu().numpy(), want, data_max).max() <= TOL_F32_REL, y
