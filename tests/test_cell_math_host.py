"""CPU: the kernels' per-cell closed form (csrc/aai_cell.cuh, compiled for the host by tests/cell_math_host.cpp)
against the oracle's literal shape classifier, pair by pair.  This is the arithmetic the GPU executes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def cellmath(built):
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libcellmath.so")
    cxx = os.environ.get("CXX", "g++")
    subprocess.run([cxx, "-O2", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "cell_math_host.cpp")], check=True)
    lib = C.CDLL(so)
    lib.aai_test_pair_areas.argtypes = [C.c_double] * 3 + [C.c_void_p] * 5 + [C.c_longlong]
    lib.aai_test_pair_areas_f32.argtypes = [C.c_double] * 3 + [C.c_void_p] * 6 + [C.c_longlong]
    lib.aai_test_image_f32.argtypes = [C.c_double] * 9 + [C.c_int] * 4 + [C.c_void_p] * 3
    lib.aai_test_footprint_rows_f32.argtypes = [C.c_double] * 5 + [C.c_int] * 3 + [C.c_void_p]
    lib.aai_test_footprint_rows_f32.restype = C.c_float
    lib.aai_test_footprint_edges_f32.argtypes = [C.c_double] * 5 + [C.c_int] * 3 + [C.c_void_p] * 2
    lib.aai_test_footprint_edges_f32.restype = C.c_float
    lib.aai_test_footprint_edges_f64.argtypes = [C.c_double] * 5 + [C.c_int] * 3 + [C.c_void_p] * 2
    lib.aai_test_edge_pair_vs_scalar.argtypes = [C.c_double] * 3 + [C.c_void_p] * 2 + [C.c_longlong]
    lib.aai_test_edge_pair_vs_scalar.restype = C.c_longlong
    lib.aai_test_quadrant_areas.argtypes = [C.c_double] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_longlong]
    lib.aai_test_pair_areas_f32x2.argtypes = [C.c_double] * 3 + [C.c_void_p] * 7 + [C.c_longlong]
    return lib


def _vertices(cx, cy, c, s, h):
    eu, ev = np.array([c, -s]), np.array([s, c])
    ctr = np.array([cx, cy])
    return np.array([ctr - h * eu - h * ev, ctr + h * eu - h * ev, ctr - h * eu + h * ev, ctr + h * eu + h * ev])


@pytest.mark.parametrize("theta,side", [(30.0, 2.7027027), (17.3, 2.7027027), (45.0, 1.7647059), (61.0, 1.7391304),
                                        (1.0, 5.5), (89.0, 3.3), (73.0, 9.0909), (44.999, 2.0), (12.0, 1.41422)])
def test_cell_area_matches_oracle_shapes(cellmath, theta, side):
    from oracle import port

    rng = np.random.default_rng(int(theta * 1000) + 5)
    c, s = np.cos(np.radians(theta)), np.sin(np.radians(theta))
    n = 3000
    cx, cy = rng.uniform(10, 20, n), rng.uniform(10, 20, n)
    i = np.floor(cx + rng.uniform(-side, side, n)).astype(np.int32)
    j = np.floor(cy + rng.uniform(-side, side, n)).astype(np.int32)
    got = np.zeros(n)
    cellmath.aai_test_pair_areas(c, s, side, cx.ctypes.data, cy.ctypes.data, i.ctypes.data, j.ctypes.data,
                                 got.ctypes.data, n)
    want = np.array([port.pair_area(_vertices(cx[k], cy[k], c, s, side / 2), int(i[k]), int(j[k])) for k in range(n)])
    assert (want > 0).sum() > n // 5  # the sample really exercises touched cells
    assert np.abs(got - want).max() < 1e-11


def test_reference_quirk_is_reproduced(cellmath):
    """A left/right edge cutting one corner must give 1/2 (1-a)(1-b), not the true 1/2 a b (Source.cpp:1055-1062)."""
    from oracle import port

    theta, side = 30.0, 2.7027027
    c, s = np.cos(np.radians(theta)), np.sin(np.radians(theta))
    rng = np.random.default_rng(99)
    n = 20000
    cx, cy = rng.uniform(10, 20, n), rng.uniform(10, 20, n)
    i = np.floor(cx + rng.uniform(-side, side, n)).astype(np.int32)
    j = np.floor(cy + rng.uniform(-side, side, n)).astype(np.int32)
    got = np.zeros(n)
    cellmath.aai_test_pair_areas(c, s, side, cx.ctypes.data, cy.ctypes.data, i.ctypes.data, j.ctypes.data,
                                 got.ctypes.data, n)
    # exact polygon clipping for comparison
    def exact(k):
        poly = [(i[k] - .5, j[k] - .5), (i[k] + .5, j[k] - .5), (i[k] + .5, j[k] + .5), (i[k] - .5, j[k] + .5)]
        for (nx, ny) in [(c, -s), (-c, s), (s, c), (-s, -c)]:
            out = []
            for a in range(len(poly)):
                p, q = poly[a], poly[(a + 1) % len(poly)]
                dp = side / 2 - ((p[0] - cx[k]) * nx + (p[1] - cy[k]) * ny)
                dq = side / 2 - ((q[0] - cx[k]) * nx + (q[1] - cy[k]) * ny)
                if dp >= 0:
                    out.append(p)
                if (dp >= 0) != (dq >= 0):
                    t = dp / (dp - dq)
                    out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
            poly = out
            if not poly:
                return 0.0
        return 0.5 * abs(sum(poly[a][0] * poly[(a + 1) % len(poly)][1] - poly[(a + 1) % len(poly)][0] * poly[a][1]
                             for a in range(len(poly))))
    ex = np.array([exact(k) for k in range(2000)])
    differs = np.abs(got[:2000] - ex) > 1e-9
    assert 0.05 < differs.mean() < 0.5  # a sizeable share of touched pairs carries the quirk
    want = np.array([port.pair_area(_vertices(cx[k], cy[k], c, s, side / 2), int(i[k]), int(j[k])) for k in range(2000)])
    assert np.abs(got[:2000] - want).max() < 1e-11


@pytest.mark.parametrize("theta,side", [(17.3, 2.7027027), (30.0, 2.7027027), (45.0, 1.7647059), (61.0, 1.7391304),
                                        (5.0, 3.3), (85.0, 1.5), (73.0, 3.9)])
def test_f32_cell_math_guard_band_catches_every_decision_flip(cellmath, theta, side):
    """FP32 areas agree with the FP64 closed form to ~1e-6 on every pair the guard band does not flag; the pairs it
    flags (redone in FP64 by the kernel) are rare."""
    rng = np.random.default_rng(int(theta * 100))
    c, s = np.cos(np.radians(theta)), np.sin(np.radians(theta))
    n = 1_000_000
    cx, cy = rng.uniform(100, 116, n), rng.uniform(100, 116, n)
    i = np.floor(cx + rng.uniform(-side, side, n) + 0.5).astype(np.int32)
    j = np.floor(cy + rng.uniform(-side, side, n) + 0.5).astype(np.int32)
    a64, a32, flag = np.zeros(n), np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.uint8)
    cellmath.aai_test_pair_areas(c, s, side, cx.ctypes.data, cy.ctypes.data, i.ctypes.data, j.ctypes.data,
                                 a64.ctypes.data, n)
    cellmath.aai_test_pair_areas_f32(c, s, side, cx.ctypes.data, cy.ctypes.data, i.ctypes.data, j.ctypes.data,
                                     a32.ctypes.data, flag.ctypes.data, n)
    keep = flag == 0
    assert np.abs(a32 - a64)[keep].max() < 5e-6
    assert flag.mean() < 2e-4


def test_f32_image_with_exact_symmetry_ties(cellmath):
    """45 degrees, scale 3, half-integer isocentre: the canvas centre pixel has cell centres exactly on the
    footprint's centre lines (v0 == +-0).  Every pixel the guard band does not flag must be within 1e-5."""
    import area_average_interpolation_b200 as aai
    from oracle import port

    w = h = 96
    ratio, angle, iso = 1.7, 45.0, (47.5, 47.5)
    rng = np.random.default_rng(7)
    src = rng.uniform(0, 255, size=(h, w)).astype(np.float32).astype(np.float64)
    p = aai.make_plan(w, h, 1.0, ratio, iso, angle)
    mod = np.ascontiguousarray(np.repeat(np.repeat(src, p.scale, axis=0), p.scale, axis=1))
    out = np.zeros((p.dst_h, p.dst_w))
    flag = np.zeros((p.dst_h, p.dst_w), dtype=np.uint8)
    cellmath.aai_test_image_f32(p.cos_t, p.sin_t, p.side, p.off_ix, p.off_iy, p.iso_x, p.iso_y, p.off_x, p.off_y,
                                mod.shape[1], mod.shape[0], p.dst_w, p.dst_h, mod.ctypes.data, out.ctypes.data,
                                flag.ctypes.data)
    st, want, _ = port.run(src, 1.0, ratio, iso, angle)
    err = np.abs(out - want) / np.maximum(np.abs(want), 1.0)
    assert err[flag == 0].max() <= 1e-5
    assert (flag[want > 0] == 1).mean() < 0.05


@pytest.mark.parametrize("theta,side", [(17.3, 2.7027027), (30.0, 2.7027027), (45.0, 1.7647059), (61.0, 2.0),
                                        (5.0, 3.3), (85.0, 1.5), (73.0, 3.9)])
def test_row_formulation_of_the_quirk_matches_the_per_cell_form(cellmath, theta, side):
    """What the FP32 kernel runs: exact areas for every cell + per-row edge-line corrections (aai_row_quirk_f32)
    must reproduce, cell by cell, the FP64 per-cell form (which is checked against the oracle above)."""
    rng = np.random.default_rng(int(theta * 10) + 1)
    c, s = np.cos(np.radians(theta)), np.sin(np.radians(theta))
    n = int(np.floor(side * (c + s) + 1)) + 2
    tau = 4e-6 * max(1.0, 1.0 / c, 1.0 / s)
    flagged = 0
    trials = 3000
    for _ in range(trials):
        cx, cy = rng.uniform(100, 116, 2)
        i0 = int(np.ceil(cx - (side / 2 * (c + s) + 0.5)))
        j0 = int(np.ceil(cy - (side / 2 * (c + s) + 0.5)))
        got = np.zeros(n * n, dtype=np.float32)
        worst = cellmath.aai_test_footprint_rows_f32(c, s, side, cx, cy, i0, j0, n, got.ctypes.data)
        ii, jj = np.meshgrid(np.arange(i0, i0 + n), np.arange(j0, j0 + n))
        i = ii.ravel().astype(np.int32)
        j = jj.ravel().astype(np.int32)
        want = np.zeros(n * n)
        cxa, cya = np.full(n * n, cx), np.full(n * n, cy)
        cellmath.aai_test_pair_areas(c, s, side, cxa.ctypes.data, cya.ctypes.data, i.ctypes.data, j.ctypes.data,
                                     want.ctypes.data, n * n)
        if worst < tau:
            flagged += 1  # the kernel redoes this pixel in FP64
        else:
            assert np.abs(got - want).max() < 5e-6
    assert flagged < 0.01 * trials


def test_packed_two_cell_form_equals_scalar_form(cellmath):
    rng = np.random.default_rng(5)
    theta, side = 17.3, 2.7027027
    c, s = np.cos(np.radians(theta)), np.sin(np.radians(theta))
    n = 200000
    cx, cy = rng.uniform(100, 116, n), rng.uniform(100, 116, n)
    i = np.floor(cx + rng.uniform(-side, side, n) + 0.5).astype(np.int32)
    j = np.floor(cy + rng.uniform(-side, side, n) + 0.5).astype(np.int32)
    a0, a1 = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
    f0, f1, flag = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=np.uint8)
    cellmath.aai_test_pair_areas_f32x2(c, s, side, cx.ctypes.data, cy.ctypes.data, i.ctypes.data, j.ctypes.data,
                                       a0.ctypes.data, a1.ctypes.data, flag.ctypes.data, n)
    cellmath.aai_test_pair_areas_f32(c, s, side, cx.ctypes.data, cy.ctypes.data, i.ctypes.data, j.ctypes.data,
                                     f0.ctypes.data, f1.ctypes.data, n)
    assert np.array_equal(a0, f0)  # lane .x of the packed form is the scalar form, operation by operation


@pytest.mark.parametrize("theta,side", [(17.3, 2.7027027), (30.0, 2.7027027), (45.0, 1.7647059), (61.0, 2.0),
                                        (5.0, 3.3), (85.0, 1.5), (73.0, 3.9), (40.0, 5.1), (52.0, 4.4), (3.5, 1.42)])
def test_edge_formulation_of_the_quirk_matches_the_per_cell_form(cellmath, theta, side):
    """What the FP32 kernel runs: exact areas for every cell + one pair of corrected cells per minor-axis grid line
    that a left/right edge crosses (aai_edge_quirk_f32), total area = L^2 + corrections.  Must reproduce, cell by cell,
    the FP64 per-cell form (which is checked against the oracle above)."""
    rng = np.random.default_rng(int(theta * 10) + 3)
    c, s = np.cos(np.radians(theta)), np.sin(np.radians(theta))
    n = int(np.floor(side * (c + s) + 1)) + 2
    tau = 4e-6 * max(1.0, 1.0 / c, 1.0 / s)
    flagged = 0
    trials = 3000
    corrected = 0
    for _ in range(trials):
        cx, cy = rng.uniform(100, 116, 2)
        i0 = int(np.ceil(cx - (side / 2 * (c + s) + 0.5)))
        j0 = int(np.ceil(cy - (side / 2 * (c + s) + 0.5)))
        got = np.zeros(n * n, dtype=np.float32)
        total = np.zeros(1, dtype=np.float32)
        worst = cellmath.aai_test_footprint_edges_f32(c, s, side, cx, cy, i0, j0, n, got.ctypes.data, total.ctypes.data)
        ii, jj = np.meshgrid(np.arange(i0, i0 + n), np.arange(j0, j0 + n))
        i = ii.ravel().astype(np.int32)
        j = jj.ravel().astype(np.int32)
        want = np.zeros(n * n)
        cxa, cya = np.full(n * n, cx), np.full(n * n, cy)
        cellmath.aai_test_pair_areas(c, s, side, cxa.ctypes.data, cya.ctypes.data, i.ctypes.data, j.ctypes.data,
                                     want.ctypes.data, n * n)
        if worst < tau:
            flagged += 1  # the kernel redoes this pixel in FP64
        else:
            assert np.abs(got - want).max() < 5e-6, (cx, cy)
            assert abs(float(total[0]) - want.sum()) < 2e-5
            corrected += abs(want.sum() - side * side) > 1e-3
    assert flagged < 0.01 * trials
    assert corrected > 0.3 * min(1.0, side * min(c, s)) * trials  # the quirk really fires


def test_edge_crossings_next_to_a_lattice_corner_are_flagged(cellmath):
    """Five footprints of the config-4 canvas (16384^2, 17.3 deg) whose right edge crosses a vertical grid line within
    1e-8 of a lattice corner: FP32 assigns the crossing to the wrong row, so the guard band must send them to FP64."""
    theta, side = 17.3, 1 / 0.37
    c, s = np.cos(np.radians(theta)), np.sin(np.radians(theta))
    n = int(np.floor(side * (c + s) + 1)) + 2
    tau = 4e-6 * max(1.0, 1.0 / c, 1.0 / s)
    for cx, cy in [(2974.930087200926, 6437.0038590545555), (1111.0160516152882, 10986.27985894395),
                   (10639.07378538135, 8740.46522064712), (6007.979929102068, 12085.16388258702),
                   (14312.099825776248, 15698.092180559619)]:
        i0 = int(np.ceil(cx - (side / 2 * (c + s) + 0.5)))
        j0 = int(np.ceil(cy - (side / 2 * (c + s) + 0.5)))
        got = np.zeros(n * n, dtype=np.float32)
        total = np.zeros(1, dtype=np.float32)
        worst = cellmath.aai_test_footprint_edges_f32(c, s, side, cx, cy, i0, j0, n, got.ctypes.data, total.ctypes.data)
        assert worst < tau


@pytest.mark.parametrize("theta,side", [(17.3, 2.7027027), (30.0, 2.7027027), (45.0, 1.7647059), (61.0, 2.0),
                                        (5.0, 3.3), (85.0, 1.5), (73.0, 3.9), (40.0, 5.1)])
def test_packed_two_edge_routine_equals_the_scalar_one(cellmath, theta, side):
    """aai_edge_quirk_pair_f32 (lane x = ALPHA edge, lane y = BETA edge, what the kernel runs) must reproduce
    aai_edge_quirk_f32 bit for bit: cells, corrections and decision margin."""
    rng = np.random.default_rng(int(theta * 7))
    c, s = np.cos(np.radians(theta)), np.sin(np.radians(theta))
    n = 50000
    hb = side / 2 * (c + s)
    g0m = -(hb + rng.uniform(0.0, 1.0, n))   # left boundary of column 0 / top of row 0 relative to the centre
    g0M = -(hb + rng.uniform(0.0, 1.0, n))
    assert cellmath.aai_test_edge_pair_vs_scalar(c, s, side, g0m.ctypes.data, g0M.ctypes.data, n) == 0


@pytest.mark.parametrize("theta,side", [(17.3, 2.7027027), (30.0, 2.7027027), (45.0, 1.7647059), (61.0, 2.0),
                                        (1.5, 5.905), (89.2, 1.5), (73.0, 3.9), (0.05, 2.0), (44.999, 2.0)])
def test_fp64_edge_formulation_matches_the_per_cell_form(cellmath, theta, side):
    """The unrolled FP64 kernel's formulation (exact areas + per-edge crossing events in FP64, no guard band) against
    the per-cell FP64 form that is checked against the oracle above -- also for angles next to the axes, which only
    the FP64 kernels serve."""
    rng = np.random.default_rng(int(theta * 100) + 9)
    c, s = np.cos(np.radians(theta)), np.sin(np.radians(theta))
    n = int(np.floor(side * (c + s) + 1)) + 2
    bad = 0
    trials = 3000
    for _ in range(trials):
        cx, cy = rng.uniform(100, 116, 2)
        i0 = int(np.ceil(cx - (side / 2 * (c + s) + 0.5)))
        j0 = int(np.ceil(cy - (side / 2 * (c + s) + 0.5)))
        got = np.zeros(n * n)
        total = np.zeros(1)
        cellmath.aai_test_footprint_edges_f64(c, s, side, cx, cy, i0, j0, n, got.ctypes.data, total.ctypes.data)
        ii, jj = np.meshgrid(np.arange(i0, i0 + n), np.arange(j0, j0 + n))
        i = ii.ravel().astype(np.int32)
        j = jj.ravel().astype(np.int32)
        want = np.zeros(n * n)
        cxa, cya = np.full(n * n, cx), np.full(n * n, cy)
        cellmath.aai_test_pair_areas(c, s, side, cxa.ctypes.data, cya.ctypes.data, i.ctypes.data, j.ctypes.data,
                                     want.ctypes.data, n * n)
        amp = max(1.0, 1.0 / c, 1.0 / s)
        if np.abs(got - want).max() > 1e-12 * amp or abs(total[0] - want.sum()) > 1e-11 * amp:
            bad += 1
    assert bad == 0


@pytest.mark.parametrize("ratio,theta", [(1.7, 45.0), (1.7, 10.0), (1.7, 80.0), (1.5, 33.0), (2.0, 61.0), (2.8, 45.0),
                                          (2.3, 17.3), (1.42, 45.0)])
def test_quadrant_weights_equal_the_sum_of_exact_cell_areas(cellmath, ratio, theta):
    """Upscaling path of the FP32 kernel: the four source-pixel weights computed directly (Green form about the corner
    of the source-pixel boundaries) equal the per-source-pixel sums of the exact cell areas; they add up to L^2."""
    S = int(ratio * np.sqrt(2.0) + 1 + 2.2e-16)
    L = S / ratio
    assert S >= 3
    t = np.deg2rad(theta)
    c, s = np.cos(t), np.sin(t)
    rng = np.random.default_rng(int(ratio * 1000 + theta))
    n = 40000
    cx = rng.uniform(10.0, 10.0 + 3 * S, size=n)
    cy = rng.uniform(10.0, 10.0 + 3 * S, size=n)
    # include centres on / next to the boundaries and lattice points
    cx[:200] = np.round(cx[:200] * 2) / 2
    cy[100:300] = np.round(cy[100:300] * 2) / 2
    out = np.zeros((n, 8))
    cellmath.aai_test_quadrant_areas(c, s, L, S, cx.ctypes.data, cy.ctypes.data, out.ctypes.data, n)
    w, r = out[:, :4], out[:, 4:]
    assert np.abs(r.sum(axis=1) - L * L).max() < 1e-9  # the checker itself: exact areas tile the footprint
    print('max abs weight error / L^2:', np.abs(w - r).max() / (L * L))
    assert np.abs(w - r).max() < 4e-7 * L * L, float(np.abs(w - r).max())
    assert (w[r == 0] == 0).all()  # a source pixel the footprint does not touch gets exactly no weight
    assert (r > 1e-3).sum(axis=1).max() == 4 and (r > 1e-3).sum(axis=1).min() == 1  # all quadrant patterns occur
