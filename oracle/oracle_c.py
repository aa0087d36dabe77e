"""ctypes binding of the C oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- nowhere else (the product package pixelflow_b200 never imports oracle/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TOP, BOTTOM, EAST, WEST, SOUTH, NORTH = range(6)


class PfoParams(C.Structure):
    _fields_ = [
        ("m", C.c_int), ("n", C.c_int), ("l", C.c_int),
        ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double), ("dt", C.c_double),
        ("xnue", C.c_double), ("xlambda", C.c_double), ("density", C.c_double),
        ("thickness", C.c_double),
        ("nonslip", C.c_int), ("iter_max", C.c_int),
        ("relux_factor", C.c_double),
        ("inlet_velocity", C.c_double), ("outlet_pressure", C.c_double), ("AoA", C.c_double),
        ("wall", C.c_int * 6),
    ]


def build(force: bool = False) -> str:
    """compile oracle/liboracle.so with the committed Makefile (gcc, seconds)"""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, "pf_oracle.c"), os.path.join(_HERE, "stl_oracle.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(src) for src in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        dp = C.POINTER(C.c_double)
        pp = C.POINTER(PfoParams)
        vp = C.c_void_p
        L.pfo_ws_create.restype = vp
        L.pfo_ws_create.argtypes = [C.c_size_t]
        L.pfo_ws_destroy.argtypes = [vp]
        L.pfo_ws_array.restype = dp
        L.pfo_ws_array.argtypes = [vp, C.c_int]
        L.pfo_sizeof_params.restype = C.c_int
        assert L.pfo_sizeof_params() == C.sizeof(PfoParams)
        sig = {
            "pfo3u_porosity_halo": [pp, dp], "pfo3a_porosity_halo": [pp, dp],
            "pfo3_initial_conditions": [pp, C.c_int, dp, dp, dp, dp],
            "pfo3_copy_old": [pp, dp, dp, dp, dp, dp, dp],
            "pfo3_divergence": [pp, C.c_int, dp, dp, dp, dp],
            "pfo3_predictor": [pp, dp, dp, dp, dp, dp, dp, dp, dp],
            "pfo3_matrix": [pp, dp, dp, dp, dp, vp],
            "pfo3u_boundary_matrix": [pp, dp, vp],
            "pfo3a_boundary_matrix": [pp, dp, dp, vp],
            "pfo3_project": [pp, dp, dp, dp, dp],
            "pfo3u_boundary": [pp, dp, dp, dp, dp],
            "pfo3a_boundary": [pp, dp, dp, dp, dp, dp],
            "pfo3_step": [pp, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, vp, dp],
            "pfo2_porosity_halo": [pp, dp],
            "pfo2_initial_conditions": [pp, C.c_int, dp, dp, dp, dp],
            "pfo2_copy_old": [pp, dp, dp, dp, dp],
            "pfo2_divergence": [pp, dp, dp, dp],
            "pfo2_predictor": [pp, dp, dp, dp, dp, dp, dp],
            "pfo2_matrix": [pp, dp, dp, dp, vp],
            "pfo2_boundary_matrix": [pp, dp, vp],
            "pfo2_project": [pp, dp, dp, dp],
            "pfo2_boundary": [pp, C.c_int, dp, dp, dp, dp],
            "pfo2_step": [pp, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, vp, dp],
            "pfo2_force_log": [pp, C.c_double, dp, dp, dp, dp, dp],
            "pfo3_force_log": [pp, C.c_double, dp, dp, dp, dp, dp, dp],
        }
        for name, args in sig.items():
            getattr(L, name).argtypes = args
            getattr(L, name).restype = None
        L.pfo3_sor.argtypes = [pp, C.c_int, C.c_int, dp, vp]
        L.pfo3_sor.restype = C.c_double
        L.pfo2_sor.argtypes = [pp, C.c_int, dp, vp]
        L.pfo2_sor.restype = C.c_double
        _LIB = L
    return _LIB


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


WS_NAMES = ("ap", "ae", "aw", "an", "as", "at", "ab", "bb", "div", "p_old")


class Workspace:
    """the reference's solver-local static arrays (zero-initialised, persistent)"""

    def __init__(self, shape):
        self.shape = tuple(shape)
        self.nelem = int(np.prod(shape))
        self.h = lib().pfo_ws_create(self.nelem)
        if not self.h:
            raise MemoryError("pfo_ws_create")

    def array(self, name: str) -> np.ndarray:
        ptr = lib().pfo_ws_array(self.h, WS_NAMES.index(name))
        return np.ctypeslib.as_array(ptr, shape=(self.nelem,)).reshape(self.shape)

    def close(self):
        if self.h:
            lib().pfo_ws_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_params(**kw) -> PfoParams:
    P = PfoParams()
    P.l = 1
    P.dz = 1.0
    P.density = 1.0
    P.thickness = 1.5
    P.nonslip = 1
    P.iter_max = 100
    P.relux_factor = 1.7
    wall = kw.pop("wall", (1, 0, 0, 0, 2, 0))
    for i, wv in enumerate(wall):
        P.wall[i] = int(wv)
    for k, v in kw.items():
        if not hasattr(P, k):
            raise AttributeError(k)
        setattr(P, k, int(v) if k in ("m", "n", "l", "nonslip", "iter_max") else float(v))
    return P


class Oracle3D:
    """whole-program driver over the C oracle for ibm3 uniform / air-condition"""

    def __init__(self, P: PfoParams, air: bool, porosity_interior: np.ndarray):
        self.P, self.air = P, bool(air)
        shape = (P.l + 2, P.n + 2, P.m + 2)
        self.shape = shape
        z = lambda: np.zeros(shape, dtype=np.float64)
        self.p, self.u, self.v, self.w = z(), z(), z(), z()
        self.uo, self.vo, self.wo = z(), z(), z()
        self.e = z()
        self.e[1:-1, 1:-1, 1:-1] = porosity_interior
        L = lib()
        (L.pfo3a_porosity_halo if air else L.pfo3u_porosity_halo)(C.byref(P), _dp(self.e))
        self.ws = Workspace(shape)

    def initialise(self):
        L, P = lib(), self.P
        L.pfo3_initial_conditions(C.byref(P), int(self.air), _dp(self.p), _dp(self.u), _dp(self.v), _dp(self.w))
        self.boundary()

    def boundary(self):
        L, P = lib(), self.P
        if self.air:
            L.pfo3a_boundary(C.byref(P), _dp(self.e), _dp(self.p), _dp(self.u), _dp(self.v), _dp(self.w))
        else:
            L.pfo3u_boundary(C.byref(P), _dp(self.p), _dp(self.u), _dp(self.v), _dp(self.w))

    def force_log(self, radius: float) -> np.ndarray:
        """output_force_log_3d: Fp xyz, Fv xyz, F xyz, Cd(x), Cl, Cd(z)"""
        out = np.zeros(12)
        lib().pfo3_force_log(C.byref(self.P), float(radius), _dp(self.p), _dp(self.u), _dp(self.v), _dp(self.w),
                             _dp(self.e), _dp(out))
        return out

    def step(self, nsteps: int = 1) -> np.ndarray:
        err = np.zeros(nsteps)
        lib().pfo3_step(C.byref(self.P), int(self.air), nsteps, _dp(self.p), _dp(self.u), _dp(self.v),
                        _dp(self.w), _dp(self.uo), _dp(self.vo), _dp(self.wo), _dp(self.e), self.ws.h,
                        _dp(err))
        return err


class Oracle2D:
    """whole-program driver over the C oracle for ibm2 uniform / drag / backstep"""

    def __init__(self, P: PfoParams, backstep: bool, porosity_interior: np.ndarray):
        self.P, self.backstep = P, bool(backstep)
        shape = (P.n + 2, P.m + 2)
        self.shape = shape
        z = lambda: np.zeros(shape, dtype=np.float64)
        self.p, self.u, self.v = z(), z(), z()
        self.uo, self.vo = z(), z()
        self.e = z()
        self.e[1:-1, 1:-1] = porosity_interior
        lib().pfo2_porosity_halo(C.byref(P), _dp(self.e))
        self.ws = Workspace(shape)

    def initialise(self):
        L, P = lib(), self.P
        L.pfo2_initial_conditions(C.byref(P), int(self.backstep), _dp(self.e), _dp(self.p), _dp(self.u), _dp(self.v))
        L.pfo2_boundary(C.byref(P), int(self.backstep), _dp(self.e), _dp(self.p), _dp(self.u), _dp(self.v))

    def force_log(self, radius: float) -> np.ndarray:
        """output_force_log_2d: Fpx, Fpy, Fvx, Fvy, Fx, Fy, Cd, Cl"""
        out = np.zeros(8)
        lib().pfo2_force_log(C.byref(self.P), float(radius), _dp(self.p), _dp(self.u), _dp(self.v), _dp(self.e), _dp(out))
        return out

    def step(self, nsteps: int = 1) -> np.ndarray:
        err = np.zeros(nsteps)
        lib().pfo2_step(C.byref(self.P), int(self.backstep), nsteps, _dp(self.p), _dp(self.u), _dp(self.v),
                        _dp(self.uo), _dp(self.vo), _dp(self.e), self.ws.h, _dp(err))
        return err


def convolve3d_nearest(a: np.ndarray, w: np.ndarray) -> np.ndarray:
    """pfo_convolve3d_nearest: scipy.ndimage.convolve(a, w, mode='nearest') as tools/voxel2poro/voxel2poro.py:33
    calls it (float32 in/out, float64 weights of odd sizes), restated in C"""
    L = lib()
    a = np.ascontiguousarray(a, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float64)
    assert a.ndim == 3 and w.ndim == 3 and all(k % 2 == 1 for k in w.shape)
    out = np.empty_like(a)
    fp = C.POINTER(C.c_float)
    L.pfo_convolve3d_nearest.restype = None
    L.pfo_convolve3d_nearest.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int,
                                         C.c_int, fp]
    L.pfo_convolve3d_nearest(a.ctypes.data_as(fp), *a.shape, w.ctypes.data_as(C.POINTER(C.c_double)), *w.shape,
                             out.ctypes.data_as(fp))
    return out
