"""CPU (-m "not gpu"): contract points the driver and the judge rely on -- the product never routes through the
oracle, bench.py fails loudly without a GPU, and the reference arm prints the agreed JSON line."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "area_average_interpolation_b200")


def _product_files():
    for base, _dirs, files in os.walk(PKG):
        if os.path.basename(base) in ("build", "gpurun_variants", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                yield os.path.join(base, f)
    for d in ("include", "examples"):
        for f in os.listdir(os.path.join(ROOT, d)):
            if f.endswith((".h", ".hpp", ".cpp", ".c", ".py")):
                yield os.path.join(ROOT, d, f)


def test_product_never_imports_links_or_executes_the_oracle():
    """oracle/ is test infrastructure: nothing shipped may import, include, dlopen or exec anything under it."""
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#\s*include\s*[\"<][^\">]*oracle)|(libaai_oracle|libaai_ref|oracle/)",
                     re.M)
    offenders = [p for p in _product_files() if pat.search(open(p, errors="replace").read())]
    assert offenders == []


def test_only_tests_smoke_and_bench_baseline_legs_use_the_oracle():
    users = []
    for f in os.listdir(ROOT):
        if f.endswith(".py") and re.search(r"^\s*(from|import)\s+oracle\b", open(os.path.join(ROOT, f)).read(), re.M):
            users.append(f)
    assert sorted(users) == ["__graft_entry__.py", "bench.py"]
    src = open(os.path.join(ROOT, "bench.py")).read()
    # in bench.py the oracle is imported inside the cpu_baseline / reference-arm functions only
    for m in re.finditer(r"^(\s*)(from|import)\s+oracle\b", src, re.M):
        assert len(m.group(1)) > 0, "bench.py imports the oracle at module level"
        head = src[:m.start()]
        func = re.findall(r"^def (\w+)\(", head, re.M)[-1]
        # the cpu_baseline leg and the reference arm call only the first two; _verify_rows is the checker of the `verified`
        # flag (oracle rows against the canvas the timed path produced, outside every timed region)
        assert func in ("_ref_worker", "cpu_kind", "_verify_rows"), func


def test_bench_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without a device the default arm must exit non-zero and print no result line."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0",
                          "--no-cpu-baseline"], capture_output=True, text=True, timeout=300)
    import torch

    if torch.cuda.is_available():
        return  # (on a GPU box this is the bench itself; covered by the driver)
    assert out.returncode != 0
    assert '"value"' not in out.stdout


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("output Mpixels/s") and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["workload"].startswith("cfg4")
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d.get("gpu_launches", 0) == 0
