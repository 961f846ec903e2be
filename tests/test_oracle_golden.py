"""CPU: the oracle restatement (oracle/aai_oracle.cpp) against the golden vectors produced by the compiled
upstream Source.cpp (tests/golden/make_golden.py), and against the compiled reference itself when present."""
import numpy as np
import pytest

from common import load_golden, golden_source


@pytest.fixture(scope="module")
def port(built):
    from oracle import port

    return port


def test_oracle_matches_every_golden_vector(port):
    z, meta = load_golden()
    assert len(meta["cases"]) >= 25
    for case in meta["cases"]:
        src = golden_source(case)
        st, dst, iso = port.run(src, case["src_res"], case["dst_res"], case["iso"], case["angle"], mode=case["mode"])
        want = z[case["name"]]
        assert st == 0, case["name"]
        assert list(dst.shape) == case["dst_shape"], case["name"]
        assert list(iso) == case["dst_iso"], case["name"]
        # the restatement keeps the reference's operand order: bit-identical results
        assert np.array_equal(dst, want), (case["name"], float(np.abs(dst - want).max()))


def test_oracle_validation_messages_match_reference(port):
    _, meta = load_golden()
    for e in meta["errors"]:
        p = port.plan(e["w"], e["h"], e["src_res"], e["dst_res"], (1.0, 1.0), 10.0)
        assert p["status"] != 0
        assert p["message"] == e["message"], e["name"]
        assert e["dst_iso"] == [-7.0, -9.0]  # the reference leaves dstIsocenter untouched on failure


def test_oracle_sub_rectangle_equals_full_image(port):
    z, meta = load_golden()
    case = next(c for c in meta["cases"] if c["name"] == "cfg4_f32_037_173deg")
    src = golden_source(case)
    st, part, _ = port.run(src, case["src_res"], case["dst_res"], case["iso"], case["angle"], rows=(10, 23),
                           cols=(5, 61))
    assert st == 0
    assert np.array_equal(part, z[case["name"]][10:23, 5:61])


def test_oracle_typed_and_multichannel_sources(port):
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, size=(40, 52, 3), dtype=np.uint8)
    for c in range(3):
        st, a, _ = port.run(rgb, 1.0, 0.37, (26.0, 20.0), 30.0, channel=c)
        st2, b, _ = port.run(rgb[..., c].astype(np.float64), 1.0, 0.37, (26.0, 20.0), 30.0)
        assert st == 0 and st2 == 0 and np.array_equal(a, b)


def test_oracle_against_compiled_reference_sweep(port):
    from oracle import ref

    if not ref.available:
        pytest.skip("oracle/_ref/libaai_ref.so not present (built only where /root/reference exists)")
    rng = np.random.default_rng(11)
    for t in range(24):
        w, h = int(rng.integers(8, 60)), int(rng.integers(8, 60))
        ratio = float(rng.choice([0.2, 0.37, 0.5, 0.9, 1.0, 1.7, rng.uniform(0.1, 2.5)]))
        angle = float(rng.choice([0.0, 90.0, 180.0, 270.0, 30.0, 45.0, 17.3, rng.uniform(-400, 800)]))
        iso = (float(rng.uniform(-10, w + 10)), float(rng.uniform(-10, h + 10)))
        mode = 1 if t % 4 else 2
        src = rng.uniform(0, 4096, size=(h, w))
        ok, msg, want, wiso, _ = ref.run(src, 1.0, ratio, iso, angle, mode=mode)
        st, got, giso = port.run(src, 1.0, ratio, iso, angle, mode=mode)
        assert ok and st == 0
        assert got.shape == want.shape and giso == wiso
        assert np.array_equal(got, want), (w, h, ratio, angle, iso, mode)


def test_exact_mode_checker_properties(port):
    """Row f4's own checker (mode 3: Sutherland-Hodgman clip + shoelace, NOT the reference's algorithm): interior
    footprints have total area L^2, axis-aligned angles equal the reference mode bit for bit in area, and on rotated
    inputs it differs from the reference exactly where the shape-2/4 quirk fires."""
    rng = np.random.default_rng(5)
    src = rng.uniform(0, 4096, size=(96, 96))
    p = port.plan(96, 96, 1.0, 0.37, (48.0, 48.0), 17.3)
    st, exact, _, area = port.run(src, 1.0, 0.37, (48.0, 48.0), 17.3, mode=3, want_area=True)
    st1, quirk, _, area1 = port.run(src, 1.0, 0.37, (48.0, 48.0), 17.3, mode=1, want_area=True)
    assert st == 0 and st1 == 0
    L2 = p["side"] ** 2
    full = area > 0.99 * L2
    interior = full.copy()  # erode by one pixel: footprints well away from the image border
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            interior &= np.roll(np.roll(full, dy, axis=0), dx, axis=1)
    interior[[0, -1], :] = False
    interior[:, [0, -1]] = False
    assert interior.mean() > 0.4
    assert np.abs(area[interior] - L2).max() <= 1e-12 * L2
    assert np.abs(area1[interior] - L2).max() > 1e-3        # the reference's areas do not add up to L^2
    assert (np.abs(exact - quirk) > 1e-6).mean() > 0.3      # and its values differ on rotated inputs
    # axis aligned: no corner cuts -> same result up to summation rounding
    st, e0, _ = port.run(src, 1.0, 0.5, (48.0, 48.0), 90.0, mode=3)
    st1, q0, _ = port.run(src, 1.0, 0.5, (48.0, 48.0), 90.0, mode=1)
    assert np.abs(e0 - q0).max() <= 1e-10
