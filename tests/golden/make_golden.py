"""Generates tests/golden/reference_vectors.npz by running the UNMODIFIED upstream Source.cpp
(oracle/_ref/libaai_ref.so, compiled from /root/reference by oracle/Makefile) on small seeded inputs.

Run it in the build container (the only place /root/reference exists):
    make -C oracle ref && python tests/golden/make_golden.py
The inputs are regenerated from (seed, dtype, shape) by area_average_interpolation_b200.synthetic, so only
the reference OUTPUTS are stored.  The upstream repository ships no tests, golden vectors or sample data
(SURVEY.md §8c); these vectors are the parity pin of the oracle and of the CUDA path.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from area_average_interpolation_b200.synthetic import synthetic_image  # noqa: E402
from oracle import ref  # noqa: E402

# name, w, h, dtype, src_res, dst_res, iso, angle, mode
CASES = [
    # reduced replicas of the five BASELINE.json configurations (same ratio / angle / isocentre rule)
    ("cfg1_u8_half_0deg", 64, 64, "uint8", 1.0, 0.5, (32.0, 32.0), 0.0, 1),
    ("cfg1_u8_half_0deg_halfiso", 64, 64, "uint8", 1.0, 0.5, (31.5, 31.5), 0.0, 1),
    ("cfg2_u8_037_30deg", 128, 128, "uint8", 1.0, 0.37, (64.0, 64.0), 30.0, 1),
    ("cfg3_u8_17_45deg", 40, 40, "uint8", 1.0, 1.7, (19.5, 19.5), 45.0, 1),
    ("cfg4_f32_037_173deg", 160, 160, "float32", 1.0, 0.37, (80.0, 80.0), 17.3, 1),
    ("cfg5_f32_half_0deg", 96, 96, "float32", 1.0, 0.5, (48.0, 48.0), 0.0, 1),
    # every quadrant, non-square, off-centre / negative isocentre, expansion factors 2..4, theta >= 45 branch
    ("q1_1173deg", 50, 70, "float64", 1.0, 0.37, (25.0, 35.0), 117.3, 1),
    ("q2_200deg", 50, 70, "float64", 1.0, 0.37, (25.0, 35.0), 200.0, 1),
    ("q3_3055deg_s2_iso0", 50, 70, "float64", 1.0, 0.9, (0.0, 0.0), 305.5, 1),
    ("s4_61deg", 24, 24, "float64", 1.0, 2.3, (12.0, 12.0), 61.0, 1),
    ("neg_angle_far_iso", 33, 47, "float64", 1.0, 0.2, (100.0, -5.0), -12.0, 1),
    ("deg90", 64, 64, "float64", 1.0, 0.37, (32.0, 32.0), 90.0, 1),
    ("deg180_half", 48, 40, "float64", 1.0, 0.5, (24.0, 20.0), 180.0, 1),
    ("deg270_half", 48, 40, "float64", 1.0, 0.5, (23.5, 20.5), 270.0, 1),
    ("s2_0deg", 32, 32, "float64", 1.0, 1.0, (16.0, 16.0), 0.0, 1),
    ("s3_0deg", 20, 20, "float64", 1.0, 2.0, (10.0, 10.0), 0.0, 1),
    ("near45_lo", 64, 64, "float64", 1.0, 0.37, (32.0, 32.0), 44.999, 1),
    ("near45_hi", 64, 64, "float64", 1.0, 0.37, (32.0, 32.0), 45.001, 1),
    ("dpi_default_15deg", 120, 120, "float64", 150.0, 25.4, (60.0, 60.0), 1.5, 1),
    ("big_downscale_73deg", 200, 150, "float64", 1.0, 0.11, (100.0, 75.0), 73.0, 1),
    ("wrap_7657deg", 40, 56, "float64", 2.0, 0.74, (20.0, 28.0), 765.7, 1),
    # structured / degenerate inputs: footprint vertices and edges exactly on grid lines.  The REFERENCE ITSELF is
    # ill-conditioned on some of these pixels (its answer flips when the isocentre moves by 1e-11),
    # so the CUDA path is compared on them under a conditioning mask (tests/test_gpu_parity.py); the oracle
    # restatement must still reproduce them bit for bit.
    ("degenerate_45deg_integer_iso", 40, 56, "float64", 2.0, 0.74, (20.0, 28.0), 765.0, 1),
    ("degenerate_30deg_half_iso", 64, 64, "float64", 1.0, 0.37, (31.5, 31.5), 30.0, 1),
    # fast mode (Source.cpp:584), incl. the shipped user settings 1528-1534 on a reduced image
    ("fast_dpi_default", 120, 120, "float64", 150.0, 25.4, (60.0, 60.0), 1.5, 2),
    ("fast_037_30deg", 96, 96, "float64", 1.0, 0.37, (48.0, 48.0), 30.0, 2),
    ("fast_17_45deg", 32, 32, "float64", 1.0, 1.7, (15.5, 15.5), 45.0, 2),
    ("fast_half_0deg", 64, 64, "float64", 1.0, 0.5, (32.0, 32.0), 0.0, 2),
    ("fast_q2_200deg", 50, 70, "float64", 1.0, 0.37, (25.0, 35.0), 200.0, 2),
    # round 2: more structured degenerate inputs (appended, so that the seeds of the cases above do not move), compared
    # under the conditioning mask like the two above
    ("degenerate_45deg_integer_iso_upscale", 48, 48, "float64", 1.0, 1.7, (24.0, 24.0), 45.0, 1),  # cfg3's shape, lattice iso
    ("degenerate_45deg_integer_iso_rect", 72, 40, "float64", 1.0, 0.37, (36.0, 20.0), 45.0, 1),
    ("degenerate_30deg_half_iso_rect", 80, 56, "float64", 1.0, 0.5, (39.5, 27.5), 30.0, 1),   # L = 2
    ("degenerate_135deg_integer_iso", 64, 64, "float64", 1.0, 0.5, (32.0, 32.0), 135.0, 1),    # quadrant 1 + exact 45 deg
    ("degenerate_60deg_half_iso_fast", 60, 60, "float64", 1.0, 0.37, (29.5, 29.5), 60.0, 2),   # fast mode: centres on edges
]

# validation failures: (name, w, h, src_res, dst_res) -> message
ERROR_CASES = [
    ("err_xy", 8, 8, (1.0, 2.0), (1.0, 1.0)),
    ("err_dst_xy", 8, 8, (1.0, 1.0), (1.0, 1.5)),
    ("err_zero_res", 8, 8, (0.0, 0.0), (1.0, 1.0)),
    ("err_neg_dst_res", 8, 8, (1.0, 1.0), (-1.0, -1.0)),
    ("err_no_rows", 8, 0, (1.0, 1.0), (1.0, 1.0)),
    ("err_no_cols", 0, 8, (1.0, 1.0), (1.0, 1.0)),
]


def main():
    assert ref.available, "build oracle/_ref first: make -C oracle ref"
    arrays, meta = {}, {"cases": [], "errors": []}
    for i, (name, w, h, dt, sres, dres, iso, ang, mode) in enumerate(CASES):
        seed = 20201 + i
        src = synthetic_image(w, h, dt, seed)
        ok, msg, dst, diso, sec = ref.run(src.astype(np.float64), sres, dres, iso, ang, mode=mode)
        assert ok, (name, msg)
        arrays[name] = dst
        meta["cases"].append(dict(name=name, w=w, h=h, dtype=dt, seed=seed, src_res=sres, dst_res=dres, iso=list(iso),
                                  angle=ang, mode=mode, dst_iso=list(diso), dst_shape=list(dst.shape),
                                  degenerate=name.startswith("degenerate")))
        print(f"{name}: {w}x{h} -> {dst.shape[1]}x{dst.shape[0]} dstIso={diso} {sec*1e3:.0f} ms")
    for name, w, h, sres, dres in ERROR_CASES:
        src = np.ones((h, w)) if h and w else np.zeros((h, w))
        ok, msg, dst, diso, _ = ref.run(src, sres, dres, (1.0, 1.0), 10.0, dst_iso_in=(-7.0, -9.0))
        assert not ok
        meta["errors"].append(dict(name=name, w=w, h=h, src_res=list(sres), dst_res=list(dres), message=msg,
                                   dst_iso=list(diso)))
        print(f"{name}: '{msg}' dstIso untouched={diso}")
    arrays["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    out = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(out, **arrays)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
