import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
if len(sys.argv) > 1:
    import area_average_interpolation_b200 as aai
    from oracle import port
    dt, w, h, r, ang, arith, odt = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), float(sys.argv[5]), int(sys.argv[6]), sys.argv[7]
    rng = np.random.default_rng(1)
    src = rng.uniform(0, 255, size=(h, w))
    src = np.floor(src).astype(np.uint8) if dt == "uint8" else src.astype(dt)
    op = aai.AreaAverageInterpolation(arith=arith, out_dtype=np.dtype(odt))
    got = op.areaAverageInterpolation(src, 1.0, r, (w / 2, h / 2), ang)
    st, want, _ = port.run(src, 1.0, r, (w / 2, h / 2), ang)
    print("OK", sys.argv[1:], "max abs err", np.abs(got.dst - want).max())
else:
    for args in [("float64", 64, 64, 0.5, 0, 0, "float64"), ("float32", 64, 64, 0.5, 0, 0, "float64"), ("uint8", 64, 64, 0.5, 0, 0, "float64"),
                 ("uint8", 512, 512, 0.5, 0, 1, "float32"), ("float32", 512, 512, 0.5, 0, 1, "float32"), ("float32", 300, 200, 0.3, 0, 1, "float32"),
                 ("float32", 300, 200, 0.15, 0, 1, "float32"), ("float64", 300, 200, 0.7, 0, 0, "float64"), ("uint8", 300, 200, 0.37, 0, 1, "uint8")]:
        out = subprocess.run([sys.executable, __file__] + [str(a) for a in args], capture_output=True, text=True)
        print(out.stdout.strip()[-300:] or out.stderr.strip()[-400:])
