"""TEST INFRASTRUCTURE ONLY.

ctypes clients for the two CPU checkers of the area-average hot path:

* ``oracle.ref``  -> ``oracle/_ref/libaai_ref.so``: the UNMODIFIED upstream ``Source.cpp`` (compiled from
  ``/root/reference`` by ``oracle/Makefile``; the prebuilt file travels to the GPU box).
* ``oracle.port`` -> ``oracle/libaai_oracle.so``: this repo's CPU restatement (``oracle/aai_oracle.cpp``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package.  The product (``area_average_interpolation_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libaai_ref.so")
PORT_SO = os.path.join(HERE, "libaai_oracle.so")

STATUS_STRINGS = {
    0: "",
    1: "Assumed X & Y resolution are same.",
    2: "0 or negative resolution is not acceptable.",
    3: "There is no data in src array.",
    4: "There is no data in the second dimension of src array.",
}


def build(verbose: bool = False) -> None:
    """Compile the checkers (the upstream one only where /root/reference exists)."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout, out.stderr)
    out.check_returncode()


def _pair(v):
    if np.isscalar(v):
        return float(v), float(v)
    return float(v[0]), float(v[1])


class _Ref:
    """The compiled upstream operator (Source.cpp:55 / :584), single-threaded as shipped."""

    def __init__(self):
        self._lib = None

    @property
    def available(self) -> bool:
        return os.path.exists(REF_SO)

    def lib(self):
        if self._lib is None:
            lib = C.CDLL(REF_SO)
            lib.aai_ref_run.restype = C.c_int
            lib.aai_ref_run.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_double] * 7 + [
                C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                C.c_char_p, C.c_int, C.POINTER(C.c_double)]
            lib.aai_ref_fetch.argtypes = [C.c_void_p]
            lib.aai_ref_get_area.restype = C.c_double
            lib.aai_ref_get_area.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                             C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
            lib.aai_ref_intersection_type.restype = C.c_int
            lib.aai_ref_intersection_type.argtypes = [C.c_double] * 8 + [C.POINTER(C.c_double)] * 2
            self._lib = lib
        return self._lib

    def run(self, src, src_res, dst_res, iso, angle, mode=1, dst_iso_in=(0.0, 0.0)):
        """Returns (ok, message, dst[h,w] float64, (dstIsoX, dstIsoY), seconds)."""
        lib = self.lib()
        src = np.ascontiguousarray(src, dtype=np.float64)
        if src.ndim != 2:
            raise ValueError("single-channel 2-D image expected")
        h, w = src.shape
        rx, ry = _pair(src_res)
        dx, dy = _pair(dst_res)
        ix, iy = _pair(iso)
        dw, dh = C.c_int(0), C.c_int(0)
        ox, oy = C.c_double(dst_iso_in[0]), C.c_double(dst_iso_in[1])
        sec = C.c_double(0)
        msg = C.create_string_buffer(256)
        ok = lib.aai_ref_run(mode, src.ctypes.data if src.size else None, w, h, rx, ry, dx, dy, ix, iy, float(angle),
                             C.byref(dw), C.byref(dh), C.byref(ox), C.byref(oy), msg, 256, C.byref(sec))
        dst = np.empty((dh.value, dw.value), dtype=np.float64)
        if dst.size:
            lib.aai_ref_fetch(dst.ctypes.data)
        lib.aai_ref_release()
        return bool(ok), msg.value.decode(), dst, (ox.value, oy.value), sec.value

    def get_area(self, xa, xb, ya, yb, centre_in, vertex_in, vx=-1.0, vy=-1.0) -> float:
        lib = self.lib()
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (xa, xb, ya, yb)]
        args = []
        for a in arrs:
            args += [a.ctypes.data if a.size else None, int(a.size)]
        return lib.aai_ref_get_area(*args, int(centre_in), int(vertex_in), float(vx), float(vy))


class _PlanOut(C.Structure):
    _fields_ = [("status", C.c_int), ("scale", C.c_uint), ("quadrant", C.c_int), ("theta", C.c_double),
                ("sn", C.c_double), ("cs", C.c_double), ("side", C.c_double), ("modW", C.c_longlong),
                ("modH", C.c_longlong), ("dstW", C.c_longlong), ("dstH", C.c_longlong), ("dstIsoX", C.c_double),
                ("dstIsoY", C.c_double)]


_DTYPES = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.uint8): 2}


class _Port:
    """This repo's CPU restatement (oracle/aai_oracle.cpp), OpenMP over destination rows."""

    def __init__(self):
        self._lib = None

    @property
    def available(self) -> bool:
        return os.path.exists(PORT_SO)

    def lib(self):
        if self._lib is None:
            if not os.path.exists(PORT_SO):
                build()
            lib = C.CDLL(PORT_SO)
            lib.aai_oracle_plan.restype = C.c_int
            lib.aai_oracle_plan.argtypes = [C.c_longlong, C.c_longlong] + [C.c_double] * 7 + [C.POINTER(_PlanOut)]
            lib.aai_oracle_run.restype = C.c_int
            lib.aai_oracle_run.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int,
                                           C.c_longlong, C.c_longlong] + [C.c_double] * 7 + [C.c_longlong] * 4 + [
                                              C.c_void_p, C.c_void_p, C.c_int]
            lib.aai_oracle_pair_area.restype = C.c_double
            lib.aai_oracle_pair_area.argtypes = [C.c_void_p, C.c_int, C.c_int]
            self._lib = lib
        return self._lib

    def plan(self, w, h, src_res, dst_res, iso, angle) -> dict:
        rx, ry = _pair(src_res)
        dx, dy = _pair(dst_res)
        ix, iy = _pair(iso)
        out = _PlanOut()
        self.lib().aai_oracle_plan(int(w), int(h), rx, ry, dx, dy, ix, iy, float(angle), C.byref(out))
        d = {k: getattr(out, k) for k, _ in _PlanOut._fields_}
        d["message"] = STATUS_STRINGS.get(out.status, "unknown")
        return d

    def run(self, src, src_res, dst_res, iso, angle, mode=1, rows=None, cols=None, channel=0, threads=0,
            want_area=False):
        """src: [h,w] or [h,w,c] array of float64 / float32 / uint8.  Returns (status, dst, (isoX, isoY)[, area])."""
        src = np.asarray(src)
        if src.dtype not in _DTYPES:
            src = src.astype(np.float64)
        if src.ndim == 2:
            h, w = src.shape
            ch = 1
        else:
            h, w, ch = src.shape
        if not src.flags.c_contiguous:
            src = np.ascontiguousarray(src)
        p = self.plan(w, h, src_res, dst_res, iso, angle)
        if p["status"]:
            return p["status"], np.empty((0, 0)), (0.0, 0.0)
        y0, y1 = rows if rows is not None else (0, p["dstH"])
        x0, x1 = cols if cols is not None else (0, p["dstW"])
        y0, y1 = max(0, y0), min(p["dstH"], y1)
        x0, x1 = max(0, x0), min(p["dstW"], x1)
        out = np.zeros((max(0, y1 - y0), max(0, x1 - x0)), dtype=np.float64)
        area = np.zeros_like(out) if want_area else None
        rx, ry = _pair(src_res)
        dx, dy = _pair(dst_res)
        ix, iy = _pair(iso)
        st = self.lib().aai_oracle_run(mode, src.ctypes.data, _DTYPES[src.dtype], src.strides[0], ch, channel, w, h,
                                       rx, ry, dx, dy, ix, iy, float(angle), x0, x1, y0, y1, out.ctypes.data,
                                       area.ctypes.data if want_area else None, int(threads))
        res = (st, out, (p["dstIsoX"], p["dstIsoY"]))
        return res + (area,) if want_area else res

    def pair_area(self, vertices, sx, sy) -> float:
        v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(8)
        return self.lib().aai_oracle_pair_area(v.ctypes.data, int(sx), int(sy))


ref = _Ref()
port = _Port()
