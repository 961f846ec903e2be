import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import area_average_interpolation_b200 as aai
if len(sys.argv) > 1:
    aai.LIB_PATH = os.path.abspath(sys.argv[1])
from common import load_golden, golden_source
z, meta = load_golden()
for arith, od in ((aai.ARITH_F64, np.float64), (aai.ARITH_F32, np.float32)):
    op = aai.AreaAverageInterpolation(arith=arith, out_dtype=od)
    for case in meta["cases"]:
        src = golden_source(case)
        if arith == aai.ARITH_F32 and src.dtype == np.float64:
            src = src.astype(np.float32)
        try:
            f = op.areaAverageInterpolation if case["mode"] == 1 else op.fastAreaAverageInterpolation
            r = f(src, case["src_res"], case["dst_res"], case["iso"], case["angle"])
            p = r.plan
            print("ok  ", arith, case["name"], "scale", p.scale, "quadrant", p.quadrant, flush=True)
        except Exception as e:
            print("FAIL", arith, case["name"], str(e)[:150], flush=True)
            sys.exit(1)
