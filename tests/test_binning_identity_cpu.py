"""CPU: the identity the source-side binning kernel (aai_kernels_bin.cu) rests on, checked against the oracle without a GPU.

fastAreaAverageInterpolation (Source.cpp:584-911) averages, per canvas pixel, the expanded-source pixels whose CENTRE lies
in the pixel's footprint.  The footprints of neighbouring canvas pixels are the cells of a rotated square lattice, so every
source pixel belongs to exactly ONE footprint: binning each source pixel by the inverse of the centre map (212-219) and
dividing the per-bin sums by the per-bin counts must reproduce the reference wherever a footprint lies inside the image."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def oracle(built):
    from oracle import port

    return port


@pytest.fixture(scope="module")
def aai(built):
    import area_average_interpolation_b200 as m  # the geometry plan is host logic: no GPU needed

    return m


@pytest.mark.parametrize("w,h,ratio,angle,iso", [
    (96, 80, 0.37, 17.3, (48.0, 40.0)),
    (120, 90, 0.45, 61.0, (60.3, 44.8)),
    (77, 131, 0.3, 33.0, (30.0, 70.0)),
])
def test_binning_by_the_inverse_centre_map_equals_the_reference_fast_mode(oracle, aai, w, h, ratio, angle, iso):
    rng = np.random.default_rng(w * 13 + h)
    src = rng.uniform(0.0, 4096.0, size=(h, w))
    st, want, _ = oracle.run(src, 1.0, ratio, iso, angle, mode=2)
    assert st == 0
    p = aai.make_plan(w, h, 1.0, ratio, iso, angle)
    assert p.status == 0 and p.scale == 1 and p.quadrant == 0
    dh, dw = want.shape
    assert (dh, dw) == (p.dst_h, p.dst_w)
    # forward centre map C(x, y) = C0 + x a + y b with a = L (cos, -sin), b = L (sin, cos); C0 from the
    # reference's own expression of the centre of canvas pixel (0, 0) (212-219)
    L, c, s = p.side, p.cos_t, p.sin_t
    u0 = (p.off_ix * L - p.iso_x) + p.off_x
    v0 = (p.off_iy * L - p.iso_y) + p.off_y
    c0x, c0y = (u0 * c + v0 * s) + p.iso_x, (-u0 * s + v0 * c) + p.iso_y
    # inverse: canvas coordinates (U, V) of every source pixel centre (i, j); pixel X covers |U - X| <= 1/2
    jj, ii = np.mgrid[0:h, 0:w].astype(np.float64)
    U = ((ii - c0x) * c - (jj - c0y) * s) / L
    V = ((ii - c0x) * s + (jj - c0y) * c) / L
    X, Y = np.rint(U).astype(np.int64), np.rint(V).astype(np.int64)
    on_edge = (np.abs(np.abs(U - X) - 0.5) < 1e-9) | (np.abs(np.abs(V - Y) - 0.5) < 1e-9)  # counted by two footprints
    ok = (X >= 0) & (X < dw) & (Y >= 0) & (Y < dh)
    sums, counts = np.zeros((dh, dw)), np.zeros((dh, dw))
    np.add.at(sums, (Y[ok], X[ok]), src[ok])
    np.add.at(counts, (Y[ok], X[ok]), 1.0)
    got = np.where(counts > 0, sums / np.maximum(counts, 1.0), 0.0)
    # compare where the footprint (half diagonal L / sqrt 2) lies inside the image and no source pixel sits on its edge
    yy, xx = np.mgrid[0:dh, 0:dw].astype(np.float64)
    cx, cy = c0x + xx * L * c + yy * L * s, c0y - xx * L * s + yy * L * c
    r = L / np.sqrt(2.0) + 1.0
    inside = (cx - r >= 0) & (cx + r <= w - 1) & (cy - r >= 0) & (cy + r <= h - 1)
    tie = np.zeros((dh, dw), dtype=bool)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            sel = on_edge & (X + dx >= 0) & (X + dx < dw) & (Y + dy >= 0) & (Y + dy < dh)
            tie[(Y + dy)[sel], (X + dx)[sel]] = True
    check = inside & ~tie
    assert check.sum() > 0.3 * dh * dw
    err = np.abs(got - want)[check] / np.maximum(np.abs(want[check]), 1e-300)
    assert err.max() <= 1e-12, float(err.max())
    assert (counts[check] > 0).all()
