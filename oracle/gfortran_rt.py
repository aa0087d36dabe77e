"""The reference's Fortran RUNTIME as a checker (test infrastructure only).

No Fortran compiler exists in this image, but libgfortran.so.5 -- what a gfortran build of the reference links and
what executes its `write(65,"(3(f16.4,1x))") ...`, `write(*,*) ...` and `read(52,*) x, y, z, poro_val` statements --
ships inside numpy / scipy.  oracle/gfortran_rt.c issues the same runtime calls gfortran generates; this module finds
the library and wraps the three entry points.  `available()` is False when no libgfortran.so.5 can be loaded; the
tests then fall back to the committed outputs in tests/golden/gfortran_io.npz (tests/golden/make_gfortran_io.py).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_STATE = None


def find_libgfortran() -> str | None:
    cands = []
    for mod in ("numpy", "scipy"):
        try:
            m = __import__(mod)
            cands += sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(m.__file__)), mod + ".libs",
                                                   "libgfortran*.so.5*")))
        except Exception:
            pass
    cands += sorted(glob.glob("/usr/lib/x86_64-linux-gnu/libgfortran.so.5*")) + sorted(glob.glob("/usr/lib64/libgfortran.so.5*"))
    return cands[0] if cands else None


def _load():
    global _LIB, _STATE
    if _STATE is not None:
        return _LIB
    _STATE = False
    so = os.path.join(_HERE, "libgfrt.so")
    src = os.path.join(_HERE, "gfortran_rt.c")
    try:
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s", "libgfrt.so"])
        path = find_libgfortran()
        if path is None:
            return None
        L = C.CDLL(so)
        L.gfrt_open.argtypes = [C.c_char_p]
        if L.gfrt_open(path.encode()) != 0:
            return None
        L.gfrt_formatted_write.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.gfrt_list_write.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int),
                                      C.POINTER(C.c_double), C.c_char_p, C.c_int]
        L.gfrt_list_read_record.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.gfrt_read_settings.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_char_p, C.c_char_p,
                                         C.POINTER(C.c_int)]
        _LIB, _STATE = L, True
    except Exception:
        _LIB = None
    return _LIB


def available() -> bool:
    return _load() is not None


def formatted_write(fmt: str, values, per_record: int) -> bytes:
    """the bytes a file receives from `write(u, fmt) v(1:per_record)` repeated over `values`"""
    L = _load()
    vals = [float(v) for v in values]
    assert len(vals) % per_record == 0
    arr = (C.c_double * len(vals))(*vals)
    cap = 64 * len(vals) + 1024
    out = C.create_string_buffer(cap)
    n = L.gfrt_formatted_write(fmt.encode(), arr, per_record, len(vals) // per_record, out, cap)
    if n < 0:
        raise RuntimeError("gfrt_formatted_write failed")
    return out.raw[:n]


def list_write(*items) -> bytes:
    """one `write(u,*) items...` record (str -> character, bool -> logical, int -> integer(4), float -> real(8))"""
    L = _load()
    kinds, strs, ints, reals = [], [], [], []
    for it in items:
        if isinstance(it, str):
            kinds.append(0); strs.append(it.encode())
        elif isinstance(it, bool):
            kinds.append(3); ints.append(1 if it else 0)
        elif isinstance(it, int):
            kinds.append(1); ints.append(it)
        else:
            kinds.append(2); reals.append(float(it))
    ck = (C.c_int * max(len(kinds), 1))(*kinds)
    cs = (C.c_char_p * max(len(strs), 1))(*strs)
    ci = (C.c_int * max(len(ints), 1))(*ints)
    cr = (C.c_double * max(len(reals), 1))(*reals)
    out = C.create_string_buffer(4096)
    n = L.gfrt_list_write(ck, len(kinds), cs, ci, cr, out, 4096)
    if n < 0:
        raise RuntimeError("gfrt_list_write failed")
    return out.raw[:n]


def list_read_record(line: str):
    """`read(line,*) x, y, z, v` -> (iostat, x, y, z, v)"""
    L = _load()
    xyz = (C.c_int * 3)(-1, -1, -1)
    v = C.c_double(float("nan"))
    b = line.encode()
    ios = L.gfrt_list_read_record(b, len(b), xyz, C.byref(v))
    return ios, xyz[0], xyz[1], xyz[2], v.value


REAL_NAMES = ("xnue", "xlambda", "density", "width", "height", "depth", "time", "inlet_velocity", "outlet_pressure", "AoA",
              "thickness", "threshold", "radius", "center_x", "center_y", "center_z", "relux_factor")
INT_NAMES = ("istep_out", "istep_max", "nonslip", "iter_max")


def read_settings(path: str, defaults: dict | None = None) -> dict:
    """the reference's read_settings (lib/global.f90:47-62) executed by the runtime's namelist reader on `path`.
    Objects a group does not mention keep `defaults` (the reference leaves them uninitialised)."""
    L = _load()
    defaults = defaults or {}
    reals = (C.c_double * 19)(*[float(defaults.get(n, 0.0)) for n in REAL_NAMES], 0.0, 0.0)
    ints = (C.c_int * 4)(*[int(defaults.get(n, 0)) for n in INT_NAMES])
    folder = C.create_string_buffer(b" " * 50, 51)
    csv = C.create_string_buffer(b" " * 50, 51)
    ios = C.c_int(0)
    rc = L.gfrt_read_settings(path.encode(), reals, ints, folder, csv, C.byref(ios))
    out = {"rc": rc, "iostat": ios.value}
    out.update({n: reals[q] for q, n in enumerate(REAL_NAMES)})
    out.update({n: ints[q] for q, n in enumerate(INT_NAMES)})
    out["nonslip"] = bool(ints[2])
    out["output_folder"] = folder.raw[:50].decode().rstrip()
    out["csv_file"] = csv.raw[:50].decode().rstrip()
    return out
