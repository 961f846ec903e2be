"""Row f2 (CSV reader/writer + main driver, Source.cpp:1437-1598).

CPU: the reference's OWN main() (compiled, unmodified, in oracle/_ref) is run on a generated Test_film_dose.csv with
its hard-coded user settings; reading the same file with this repo's CSV reader, evaluating with the oracle and writing
with this repo's CSV writer must give a byte-identical `_mod.csv`.  GPU: the `aai_main` driver itself."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))


def _build(tmp_path, src, name, link_lib):
    exe = str(tmp_path / name)
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", src, "-o", exe]
    if link_lib:
        libdir = os.path.join(ROOT, "area_average_interpolation_b200")
        cmd += ["-L" + libdir, "-laai_b200", "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True)
    return exe


def _write_input(path, w=300, h=260, seed=3):
    rng = np.random.default_rng(seed)
    img = rng.uniform(0, 4096, size=(h, w))
    with open(path, "w") as f:
        for row in img:
            f.write(",".join(repr(float(v)) for v in row) + "\n")
    return img


def _run_reference_main(cwd):
    from oracle import REF_SO

    code = ("import ctypes,sys; lib=ctypes.CDLL(%r); f=getattr(lib,'_Z18aai_reference_mainv'); f.restype=ctypes.c_int; "
            "sys.exit(f() & 0xff)" % REF_SO)
    return subprocess.run([sys.executable, "-c", code], cwd=cwd, capture_output=True, text=True)


def test_csv_reader_writer_reproduce_the_reference_main_byte_for_byte(tmp_path, built):
    from oracle import port, ref

    if not ref.available:
        pytest.skip("oracle/_ref/libaai_ref.so not present")
    img = _write_input(tmp_path / "Test_film_dose.csv")
    out = _run_reference_main(str(tmp_path))
    assert out.returncode == 0 and "Run terminated correctly." in out.stdout, out.stdout[-500:]
    want = (tmp_path / "Test_film_dose_mod.csv").read_bytes()
    harness = _build(tmp_path, os.path.join(HERE, "csv_host.cpp"), "csv_host", False)
    # reader: same doubles as the text holds
    r = subprocess.run([harness, "read", str(tmp_path / "Test_film_dose.csv"), str(tmp_path / "in.f64")],
                       capture_output=True, text=True, check=True)
    w, h = map(int, r.stdout.split())
    data = np.fromfile(tmp_path / "in.f64", dtype=np.float64).reshape(h, w)
    assert np.array_equal(data, img)
    # the reference's shipped settings (Source.cpp:1528-1534): 150 -> 25.4 dpi, iso (455,455), 1.5 deg, mode 2
    st, dst, _ = port.run(data, 150.0, 25.4, (455.0, 455.0), 1.5, mode=2)
    assert st == 0
    dst.tofile(tmp_path / "out.f64")
    subprocess.run([harness, "write", str(tmp_path / "out.f64"), str(dst.shape[1]), str(dst.shape[0]),
                    str(tmp_path / "mine_mod.csv")], check=True)
    assert (tmp_path / "mine_mod.csv").read_bytes() == want
    # path convention <dir><base>_mod<ext>
    r = subprocess.run([harness, "split", "a/b/Test_film_dose.csv"], capture_output=True, text=True, check=True)
    assert r.stdout.strip() == "a/b/|Test_film_dose|.csv"


def test_csv_reader_tolerant_fields_and_rejected_inputs(tmp_path, built):
    harness = _build(tmp_path, os.path.join(HERE, "csv_host.cpp"), "csv_host", False)
    (tmp_path / "a.csv").write_text("1, 2.5e1,abc,3x\\n 4,5,,6\\n".replace("\\n", "\n"))
    r = subprocess.run([harness, "read", str(tmp_path / "a.csv"), str(tmp_path / "a.f64")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split() == ["3", "2"]  # 'abc' and '' are skipped, '3x' reads as 3
    assert np.array_equal(np.fromfile(tmp_path / "a.f64"), [1, 25, 3, 4, 5, 6])
    (tmp_path / "b.csv").write_text("1,2,3\n4,5\n")
    r = subprocess.run([harness, "read", str(tmp_path / "b.csv"), str(tmp_path / "b.f64")], capture_output=True, text=True)
    assert r.returncode == 2 and "different length" in r.stdout
    r = subprocess.run([harness, "read", str(tmp_path / "missing.csv"), str(tmp_path / "c.f64")], capture_output=True, text=True)
    assert r.returncode == 2 and "Failed to read csv file." in r.stdout


@pytest.mark.gpu
def test_aai_main_driver_matches_the_reference_main(tmp_path, built):
    from oracle import ref

    if not ref.available:
        pytest.skip("oracle/_ref/libaai_ref.so not present")
    _write_input(tmp_path / "Test_film_dose.csv")
    out = _run_reference_main(str(tmp_path))
    assert out.returncode == 0
    want_text = (tmp_path / "Test_film_dose_mod.csv").read_text()
    os.rename(tmp_path / "Test_film_dose_mod.csv", tmp_path / "ref_mod.csv")
    exe = _build(tmp_path, os.path.join(ROOT, "examples", "aai_main.cpp"), "aai_main", True)
    mine = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, text=True)  # no arguments: shipped settings
    assert mine.returncode == 0 and "Run terminated correctly." in mine.stdout, mine.stdout[-500:]
    got_text = (tmp_path / "Test_film_dose_mod.csv").read_text()
    want = np.array([[float(v) for v in ln.split(",")] for ln in want_text.strip().split("\n")])
    got = np.array([[float(v) for v in ln.split(",")] for ln in got_text.strip().split("\n")])
    assert want.shape == got.shape
    # fast mode sums the same values in a different order: identical up to the 6-digit text format
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()
    same = sum(a == b for a, b in zip(got_text.split("\n"), want_text.split("\n")))
    assert same >= 0.9 * len(want_text.split("\n"))
    # and the area-average mode through the same driver
    mine = subprocess.run([exe, "Test_film_dose.csv", "150", "25.4", "455", "455", "1.5", "1"], cwd=str(tmp_path),
                          capture_output=True, text=True)
    assert mine.returncode == 0, mine.stdout[-300:]
