"""The C++ host mirror (csrc/aai.hpp): compiles against the C ABI with the reference's signature; on a box without
a GPU it must fail loudly (no CPU fallback); on the GPU box it must reproduce the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "dropin_main.cpp")


def _build(tmp_path, built):
    exe = str(tmp_path / "dropin_main")
    libdir = os.path.join(ROOT, "area_average_interpolation_b200")
    cxx = os.environ.get("CXX", "g++")
    subprocess.run([cxx, "-std=c++17", SRC, "-L" + libdir, "-laai_b200", "-Wl,-rpath," + libdir, "-o", exe], check=True)
    return exe


def test_dropin_compiles_and_fails_loudly_without_gpu(tmp_path, built):
    import area_average_interpolation_b200 as aai

    exe = _build(tmp_path, built)
    if aai.device_count() > 0:
        pytest.skip("a GPU is present")
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode != 0
    assert "no CPU fallback" in out.stdout and "Run terminated abnormally." in out.stdout


@pytest.mark.gpu
def test_dropin_matches_oracle_on_the_reference_user_settings(tmp_path, built):
    from oracle import port

    exe = _build(tmp_path, built)
    dump = str(tmp_path / "dst.f64")
    out = subprocess.run([exe, dump], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Run terminated correctly." in out.stdout
    # same synthetic image as examples/dropin_main.cpp, the reference's shipped settings (Source.cpp:1528-1534)
    y, x = np.mgrid[0:911, 0:911]
    src = ((x * 131 + y * 71) % 4096).astype(np.float64)
    st, want, iso = port.run(src, 150.0, 25.4, (455.0, 455.0), 1.5)
    assert f"dst {want.shape[1]}x{want.shape[0]}" in out.stdout
    assert f"dstIsocenter ({iso[0]:g}, {iso[1]:g})" in out.stdout
    got = float(out.stdout.split("dst[79][79] = ")[1].split()[0])
    assert abs(got - want[79, 79]) <= 1e-9 * abs(want[79, 79])
    # the WHOLE image that the C++ mirror returned (flatten -> aai_run_host -> unflatten), against the oracle
    whole = np.fromfile(dump, dtype=np.float64).reshape(want.shape)
    err = np.abs(whole - want) / np.maximum(np.abs(want), 1e-300)
    err[want == 0] = np.abs(whole[want == 0])
    assert err.max() <= 1e-9, float(err.max())
