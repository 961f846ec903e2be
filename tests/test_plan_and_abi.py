"""CPU: host logic of the product (plan, partitioner) and the C-ABI surface.  No GPU compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from common import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def aai(built):
    import area_average_interpolation_b200 as m

    return m


def test_library_exports_every_declared_symbol(aai):
    header = open(os.path.join(ROOT, "include", "aai.h")).read()
    names = sorted(set(re.findall(r"\b(aai_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 14
    lib = C.CDLL(aai.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/aai.h but not exported"


def test_struct_layouts_match_the_header(aai):
    # sizes implied by include/aai.h (4 x int32, 6 x int64, 14 x double) / (ptr, 5 x int64, 2 x int32)
    assert C.sizeof(aai.Plan) == 4 * 4 + 6 * 8 + 14 * 8
    assert C.sizeof(aai.Image) == 8 + 5 * 8 + 2 * 4


def test_plan_matches_golden_sizes_and_isocentres(aai):
    _, meta = load_golden()
    for case in meta["cases"]:
        p = aai.make_plan(case["w"], case["h"], case["src_res"], case["dst_res"], case["iso"], case["angle"])
        assert p.status == 0
        assert [p.dst_h, p.dst_w] == case["dst_shape"], case["name"]
        assert [p.dst_iso_x, p.dst_iso_y] == case["dst_iso"], case["name"]


def test_plan_validation_matches_reference_messages(aai):
    _, meta = load_golden()
    for e in meta["errors"]:
        p = aai.make_plan(e["w"], e["h"], e["src_res"], e["dst_res"], (1.0, 1.0), 10.0)
        assert 1 <= p.status <= 4
        assert p.message == e["message"], e["name"]
    # the operator mirror reports them like the reference: ok=False, message, dst untouched, dstIsocenter untouched
    op = aai.AreaAverageInterpolation()
    r = op.areaAverageInterpolation(np.ones((4, 4)), (1.0, 2.0), 1.0, (2, 2), 0.0, dstIsocenter=(-7.0, -9.0))
    assert (r.ok, r.message, r.dst.size, r.dst_isocenter) == (False, "Assumed X & Y resolution are same.", 0, (-7.0, -9.0))
    r = op.areaAverageInterpolation(np.ones((0, 0)), 1.0, 1.0, (2, 2), 0.0)
    assert (r.ok, r.message) == (False, "There is no data in src array.")


def test_plan_rejects_what_the_reference_cannot_handle(aai):
    assert aai.make_plan(8, 8, 1.0, 1.0, (4, 4), float("nan")).status == aai.ERR_ANGLE
    assert aai.make_plan(8, 8, 1.0, 1.0, (4, 4), float("inf")).status == aai.ERR_ANGLE
    assert aai.make_plan(1 << 31, 8, 1.0, 1.0, (4, 4), 0.0).status == aai.ERR_ARGUMENT


def test_plan_matches_oracle_plan_on_a_sweep(aai):
    from oracle import port

    rng = np.random.default_rng(0)
    for _ in range(1500):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        r = float(rng.choice([0.2, 0.37, 0.5, 0.9, 1.0, 1.7, 2.3, rng.uniform(0.05, 3)]))
        ang = float(rng.choice([0, 90, 180, 270, 360, -90, 45, 30, 17.3, rng.uniform(-400, 800)]))
        iso = (float(rng.uniform(-50, 350)), float(rng.uniform(-50, 350)))
        p, o = aai.make_plan(w, h, 1.0, r, iso, ang), port.plan(w, h, 1.0, r, iso, ang)
        assert (p.status, p.scale, p.quadrant, p.dst_w, p.dst_h, p.dst_iso_x, p.dst_iso_y, p.mod_w, p.mod_h, p.side,
                p.sin_t, p.cos_t) == (o["status"], o["scale"], o["quadrant"], o["dstW"], o["dstH"], o["dstIsoX"],
                                      o["dstIsoY"], o["modW"], o["modH"], o["side"], o["sn"], o["cs"])


def test_axis_aligned_flag(aai):
    for ang, flag in [(0, 1), (90, 1), (180, 1), (270, 1), (360, 1), (-90, 1), (45, 0), (1e-9, 0), (17.3, 0)]:
        assert aai.make_plan(32, 32, 1.0, 0.5, (16, 16), ang).axis_aligned == flag, ang


def test_partition_covers_canvas_and_balances_kernel_cost(aai):
    p = aai.make_plan(16384, 16384, 1.0, 0.37, (8192, 8192), 17.3)
    assert (p.dst_w, p.dst_h) == (7591, 7591)
    total = aai.covered_pixels(p)
    assert 0.60 * p.dst_w * p.dst_h < total < 0.68 * p.dst_w * p.dst_h  # covered fraction 0.638
    for n in (1, 2, 4, 8):
        b = aai.partition_rows(p, n)
        assert b[0] == 0 and b[-1] == p.dst_h and all(b[i] < b[i + 1] for i in range(n))
        loads = [aai.covered_pixels(p, b[i], b[i + 1]) for i in range(n)]
        assert sum(loads) == total
        # the split equalises the measured cost model (covered pixels + a fraction of the empty canvas pixels,
        # aai_band_empty_weight), so the bands at the canvas corners (many rows, short spans) hold fewer covered pixels
        w = aai.band_empty_weight(p, aai.MODE_AREA_AVERAGE, aai.ARITH_F32)
        assert w == 0.17
        cost = [loads[i] + w * ((b[i + 1] - b[i]) * p.dst_w - loads[i]) for i in range(n)]
        assert max(cost) <= 1.01 * sum(cost) / n, (n, cost)
        assert max(loads) <= 1.10 * total / n, (n, loads)
        # the same for the kernels with cheaper / dearer covered pixels
        for mode, arith in ((aai.MODE_FAST, aai.ARITH_F32), (aai.MODE_AREA_AVERAGE, aai.ARITH_F64)):
            wk = aai.band_empty_weight(p, mode, arith)
            bk = aai.partition_rows(p, n, wk)
            assert bk[0] == 0 and bk[-1] == p.dst_h and all(bk[i] < bk[i + 1] for i in range(n))
            lk = [aai.covered_pixels(p, bk[i], bk[i + 1]) for i in range(n)]
            ck = [lk[i] + wk * ((bk[i + 1] - bk[i]) * p.dst_w - lk[i]) for i in range(n)]
            assert max(ck) <= 1.01 * sum(ck) / n, (n, mode, arith, ck)
    assert aai.band_empty_weight(p, aai.MODE_FAST, aai.ARITH_F32) > 0.17 > aai.band_empty_weight(p, aai.MODE_AREA_AVERAGE, aai.ARITH_F64)
    with pytest.raises(aai.AaiError):
        aai.partition_rows(p, 4, -0.5)


def test_band_source_window_contains_every_pixel_the_oracle_reads(aai):
    # brute force on a small rotated case: the band's halo must contain all source pixels with non-zero weight
    from oracle import port

    w, h, r, ang, iso = 61, 47, 0.37, 117.3, (30.0, 20.0)
    p = aai.make_plan(w, h, 1.0, r, iso, ang)
    b = aai.partition_rows(p, 3)
    base = np.zeros((h, w))
    st, full, _ = port.run(np.ones((h, w)), 1.0, r, iso, ang)
    for k in range(3):
        x0, x1, y0, y1 = aai.band_source_window(p, b[k], b[k + 1])
        assert 0 <= x0 <= x1 <= w and 0 <= y0 <= y1 <= h
        # zero everything outside the halo: the band must not change
        masked = base.copy()
        masked[y0:y1, x0:x1] = 1.0
        st, got, _ = port.run(masked, 1.0, r, iso, ang, rows=(b[k], b[k + 1]))
        assert np.array_equal(got, full[b[k]:b[k + 1]]), k


def test_no_cpu_fallback_without_a_gpu(aai):
    if aai.device_count() > 0:
        pytest.skip("a GPU is present")
    op = aai.AreaAverageInterpolation()
    with pytest.raises(aai.AaiError) as ei:
        op.areaAverageInterpolation(np.ones((8, 8)), 1.0, 0.5, (4, 4), 0.0)
    assert ei.value.status in (aai.ERR_NO_DEVICE, aai.ERR_CUDA)


def test_peer_group_and_host_batch_fail_loudly_without_a_gpu(aai):
    """The round-2 entry points (peer group, host batch) validate their arguments on the host and have no CPU fallback."""
    import ctypes as C

    assert aai.peer_upload_chunks(0) == 4 and aai.peer_upload_chunks(99) == 4  # out of range only queries
    assert aai.peer_upload_chunks(2) == 2 and aai.peer_upload_chunks(4) == 4
    plan = aai.make_plan(64, 64, 1.0, 0.5, (32, 32), 10.0)
    h = C.c_void_p()
    assert aai.lib().aai_peer_create(C.byref(plan), 7, 1, 0, 1, 0, C.byref(h)) == aai.ERR_ARGUMENT  # bad dtype
    assert aai.lib().aai_peer_create(C.byref(plan), aai.F32, 1, 3, 2, 0, C.byref(h)) == aai.ERR_ARGUMENT  # rank >= world
    bad = aai.make_plan(64, 64, 1.0, -1.0, (32, 32), 10.0)
    assert aai.lib().aai_peer_create(C.byref(bad), aai.F32, 1, 0, 1, 0, C.byref(h)) == aai.ERR_RESOLUTION_NONPOS
    if aai.device_count() > 0:
        return
    with pytest.raises(aai.AaiError) as ei:
        aai.PeerGroup(plan, aai.F32, 1, 0, 1, 0, lambda b: [b])
    assert ei.value.status in (aai.ERR_NO_DEVICE, aai.ERR_CUDA)
    src = np.zeros((64, 64), dtype=np.float32)
    dst = np.zeros((plan.dst_h, plan.dst_w), dtype=np.float32)
    with pytest.raises(aai.AaiError) as ei:
        aai.run_host_batch(plan, [aai._host_image(src)], [aai._host_image(dst)])
    assert ei.value.status in (aai.ERR_NO_DEVICE, aai.ERR_CUDA, aai.ERR_ARGUMENT)
    assert "no CPU fallback" in ei.value.message or ei.value.status != aai.ERR_ARGUMENT


def test_partition_and_halo_properties_on_random_plans(aai):
    """Host logic of SURVEY 8e on random geometry (all quadrants, up- and downscaling, off-centre isocentres):
    bands tile the canvas, loads add up, a band's halo contains the halo of every sub-band (what the chunk-pipelined
    host path uploads progressively), and masking the source outside a band's halo does not change the band."""
    from oracle import port

    rng = np.random.default_rng(8128)
    for case in range(24):
        w, h = int(rng.integers(24, 90)), int(rng.integers(24, 90))
        ratio = float(rng.choice([0.23, 0.37, 0.5, 0.9, 1.0, 1.7, 2.3]))
        angle = float(rng.choice([0.0, 90.0, 17.3, 30.0, 61.0, 117.3, 200.0, 305.5, -12.0]))
        iso = (float(rng.uniform(-10, w + 10)), float(rng.uniform(-10, h + 10)))
        p = aai.make_plan(w, h, 1.0, ratio, iso, angle)
        assert p.status == 0
        if p.dst_h < 4:
            continue
        n = int(rng.integers(1, min(6, p.dst_h) + 1))
        b = aai.partition_rows(p, n)
        assert b[0] == 0 and b[-1] == p.dst_h and all(b[i] <= b[i + 1] for i in range(n)), (case, b)
        assert sum(aai.covered_pixels(p, b[i], b[i + 1]) for i in range(n)) == aai.covered_pixels(p)
        k = int(rng.integers(0, n))
        r0, r1 = b[k], b[k + 1]
        if r1 <= r0:
            continue
        x0, x1, y0, y1 = aai.band_source_window(p, r0, r1)
        assert 0 <= x0 <= x1 <= w and 0 <= y0 <= y1 <= h
        # sub-bands (pipeline chunks) stay inside the band's halo
        cuts = sorted(set([r0, r1] + [int(v) for v in rng.integers(r0, r1 + 1, size=3)]))
        for a, c in zip(cuts[:-1], cuts[1:]):
            sx0, sx1, sy0, sy1 = aai.band_source_window(p, a, c)
            if sy1 > sy0 and sx1 > sx0:
                assert x0 <= sx0 and sx1 <= x1 and y0 <= sy0 and sy1 <= y1, (case, a, c)
        if case % 3 == 0:  # (the oracle runs are the slow part)
            src = rng.uniform(1.0, 2.0, size=(h, w))
            st, full, _ = port.run(src, 1.0, ratio, iso, angle, rows=(r0, r1))
            masked = np.zeros_like(src)
            masked[y0:y1, x0:x1] = src[y0:y1, x0:x1]
            st2, got, _ = port.run(masked, 1.0, ratio, iso, angle, rows=(r0, r1))
            assert st == 0 and st2 == 0 and np.array_equal(got, full), case


def test_run_entry_points_validate_before_touching_cuda(aai):
    """Argument errors of the run functions are decided on the host before any CUDA call, so they are the same on a box
    without a GPU: shapes that do not match the plan, unknown mode / arithmetic, rows outside the band, a failed plan
    (its status comes back unchanged), an empty batch."""
    plan = aai.make_plan(64, 48, 1.0, 0.5, (32, 24), 10.0)
    src = np.zeros((48, 64), dtype=np.float32)
    dst = np.zeros((plan.dst_h, plan.dst_w), dtype=np.float32)
    si, di = aai._host_image(src), aai._host_image(dst)
    small = aai._host_image(np.zeros((5, 5), dtype=np.float32))

    def status_of(fn, *a, **k):
        with pytest.raises(aai.AaiError) as ei:
            fn(*a, **k)
        return ei.value.status

    assert status_of(aai.run_device, plan, si, small) == aai.ERR_ARGUMENT
    assert status_of(aai.run_device, plan, si, di, mode=4) == aai.ERR_ARGUMENT
    assert status_of(aai.run_device, plan, si, di, arith=7) == aai.ERR_ARGUMENT
    assert status_of(aai.run_device, plan, si, di, row0=3, row1=plan.dst_h + 1) == aai.ERR_ARGUMENT
    assert status_of(aai.run_device_batch, plan, [], []) == aai.ERR_ARGUMENT
    bad = aai.make_plan(64, 48, 1.0, -0.5, (32, 24), 10.0)   # the reference's second validation failure
    assert bad.status == 2
    assert status_of(aai.run_device, bad, si, di) == 2
    assert status_of(aai.run_device_batch, bad, [si, si], [di, di]) == 2
    # a source band that misses rows the canvas band needs is refused
    part = aai.Image(si.data, si.pitch_bytes, si.width, si.height, 0, 8, si.dtype, si.channels)
    assert status_of(aai.run_device, plan, part, di) == aai.ERR_ARGUMENT
    assert "source rows" in aai.last_error()


def test_header_is_plain_c_and_links_from_a_c_program(aai, tmp_path):
    """The drop-in boundary is a C ABI: include/aai.h compiles as pedantic C99 and a C program links the library."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "area_average_interpolation_b200")
    src = tmp_path / "c_client.c"
    src.write_text('#include <stdio.h>\n#include "aai.h"\n'
                   "int main(void) {\n"
                   "    aai_plan p;\n"
                   "    int st = aai_plan_create(512, 512, 1, 1, 0.5, 0.5, 256, 256, 0.0, &p);\n"
                   '    printf("%d %lld %lld %s|\\n", st, (long long)p.dst_w, (long long)p.dst_h, aai_status_string(1));\n'
                   "    return st;\n}\n")
    exe = str(tmp_path / "c_client")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
                    str(src), "-L" + libdir, "-laai_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert out.strip() == "0 256 256 Assumed X & Y resolution are same.|"


def test_python_mirror_constants_equal_the_header_enums(aai, tmp_path):
    """The ctypes mirror spells the header's enum values out by hand: compile include/aai.h and compare."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = {"AAI_MODE_AREA_AVERAGE": aai.MODE_AREA_AVERAGE, "AAI_MODE_FAST": aai.MODE_FAST,
             "AAI_ARITH_F64": aai.ARITH_F64, "AAI_ARITH_F32": aai.ARITH_F32, "AAI_ARITH_F32_STAGED": aai.ARITH_F32_STAGED,
             "AAI_ARITH_F32_BINNED": aai.ARITH_F32_BINNED, "AAI_ARITH_F32_RING": aai.ARITH_F32_RING,
             "AAI_F64": aai.F64, "AAI_F32": aai.F32, "AAI_U8": aai.U8, "AAI_ERR_ARGUMENT": aai.ERR_ARGUMENT}
    src = tmp_path / "enums.c"
    src.write_text('#include <stdio.h>\n#include "aai.h"\nint main(void) {\n' +
                   "".join(f'    printf("{n} %d\\n", (int){n});\n' for n in names) + "    return 0;\n}\n")
    exe = str(tmp_path / "enums")
    subprocess.run(["gcc", "-std=c99", "-I" + os.path.join(root, "include"), str(src), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    got = {line.split()[0]: int(line.split()[1]) for line in out.splitlines()}
    assert got == {n: int(v) for n, v in names.items()}
