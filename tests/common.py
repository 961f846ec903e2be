"""Shared helpers of the test-suite (golden vectors, tolerances)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "reference_vectors.npz")

# north-star tolerances (BASELINE.json): FP64 kernel <= 1e-9 relative; FP32 kernel <= 1e-5 relative on float
# data and <= 0.5/255 absolute on 8-bit data.
TOL_F64_REL = 1e-9
TOL_F32_REL = 1e-5
TOL_U8_ABS = 0.5 / 255.0


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
