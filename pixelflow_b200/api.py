"""Host-side mirror of the C ABI (include/pixelflow_gpu.h) over ctypes.

This is plumbing: it marshals numpy arrays into the `extern "C"` entry points a Fortran driver
binds with iso_c_binding.  All arithmetic happens in libpixelflow_gpu.so (hand-written sm_100a
CUDA); there is no Python or CPU fallback -- if the library is missing or no GPU is usable, calls
raise.

Array convention: numpy float64, C-contiguous, shape (l+2, n+2, m+2) for 3D and (n+2, m+2) for 2D,
index order [k, j, i] -- byte-identical to the Fortran `dimension(0:m+1,0:n+1,0:l+1)` arrays.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# PIXELFLOW_GPU_LIB: an alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("PIXELFLOW_GPU_LIB") or os.path.join(_PKG, "libpixelflow_gpu.so")

IBM2_UNIFORM, IBM2_BACKSTEP, IBM2_DRAG, IBM3_UNIFORM, IBM3_AIRCOND = range(5)
CASE_NAMES = {
    "ibm2_uniform": IBM2_UNIFORM, "ibm2_backstep": IBM2_BACKSTEP, "ibm2_drag": IBM2_DRAG,
    "ibm3_uniform": IBM3_UNIFORM, "ibm3_air_condition": IBM3_AIRCOND,
}
TOP, BOTTOM, EAST, WEST, SOUTH, NORTH = range(6)
FIELDS = {name: i for i, name in enumerate(
    ["u", "v", "w", "p", "u_old", "v_old", "w_old", "porosity", "div",
     "ap", "ae", "aw", "an", "as", "at", "ab", "bb"])}


class PfConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int), ("solver_case", C.c_int),
        ("m", C.c_int), ("n", C.c_int), ("l", C.c_int),
        ("host_ldx", C.c_int), ("host_ldy", C.c_int), ("host_is_slab", C.c_int),
        ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double), ("dt", C.c_double),
        ("xnue", C.c_double), ("xlambda", C.c_double), ("density", C.c_double), ("thickness", C.c_double),
        ("nonslip", C.c_int), ("iter_max", C.c_int),
        ("relux_factor", C.c_double),
        ("inlet_velocity", C.c_double), ("outlet_pressure", C.c_double), ("AoA", C.c_double),
        ("wall", C.c_int * 6),
        ("device", C.c_int), ("rank", C.c_int), ("nranks", C.c_int),
        ("nccl_unique_id", C.c_void_p),
        ("sor_variant", C.c_int), ("use_graph", C.c_int), ("halo_transport", C.c_int),
    ]


_lib = None


class PixelFlowError(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """dlopen libpixelflow_gpu.so; fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PixelFlowError(
            f"{LIB_PATH} is missing: build it with `python -m pixelflow_b200.build` "
            "(there is no CPU fallback for the hot path)")
    L = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    L.pf_abi_version.restype = C.c_int
    L.pf_config_init.argtypes = [C.POINTER(PfConfig)]
    L.pf_config_init.restype = None
    L.pf_create.argtypes = [C.POINTER(vp), C.POINTER(PfConfig)]
    L.pf_destroy.argtypes = [vp]
    L.pf_destroy.restype = None
    L.pf_last_error.argtypes = [vp]
    L.pf_last_error.restype = C.c_char_p
    L.pf_comm_unique_id.argtypes = [vp]
    L.pf_local_slab.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.pf_set_porosity.argtypes = [vp, dp]
    L.pf_upload.argtypes = [vp, dp, dp, dp, dp]
    L.pf_download.argtypes = [vp, dp, dp, dp, dp]
    L.pf_get_field.argtypes = [vp, C.c_int, dp]
    L.pf_set_field.argtypes = [vp, C.c_int, dp]
    L.pf_step.argtypes = [vp, C.c_int, dp]
    L.pf_step_host.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp]
    for name in ("pf_initial_conditions", "pf_copy_old", "pf_divergence", "pf_predictor",
                 "pf_build_poisson", "pf_project", "pf_boundary", "pf_sync"):
        getattr(L, name).argtypes = [vp]
    L.pf_sor.argtypes = [vp, C.c_int, dp]
    L.pf_last_timing.argtypes = [vp, dp, dp, C.POINTER(C.c_longlong)]
    L.pf_stream.argtypes = [vp]
    L.pf_stream.restype = vp
    if L.pf_abi_version() != 2:
        raise PixelFlowError("libpixelflow_gpu.so ABI version mismatch")
    _lib = L
    return L


EXPORTS = [
    "pf_abi_version", "pf_config_init", "pf_create", "pf_destroy", "pf_last_error", "pf_comm_unique_id",
    "pf_local_slab", "pf_set_porosity", "pf_upload", "pf_download", "pf_get_field", "pf_set_field",
    "pf_step", "pf_step_host", "pf_initial_conditions", "pf_copy_old", "pf_divergence", "pf_predictor",
    "pf_build_poisson", "pf_sor", "pf_project", "pf_boundary", "pf_sync", "pf_last_timing", "pf_stream",
    "pf_debug_fastdiv_mismatches", "pf_get_sor_variant", "pf_force_log_2d", "pf_get_halo_transport",
    "pf_convolve3d_nearest", "pf_force_log_3d", "pf_vtk_section_bytes", "pf_vtk_section",
    "pf_parse_porosity_csv", "pf_debug_quot_mismatches", "pf_gather", "pf_ranks_launch", "pf_ranks_rank",
    "pf_ranks_count", "pf_ranks_unique_id", "pf_ranks_barrier", "pf_ranks_finish", "pf_stl_signed_distance",
]


def fastdiv_mismatches(d: float, n: int = 1 << 24, seed: int = 1) -> int:
    """GPU self-check of the exact reciprocal division used for loop-invariant divisors (must be 0)."""
    L = load_library()
    L.pf_debug_fastdiv_mismatches.argtypes = [C.c_double, C.c_longlong, C.c_ulonglong, C.POINTER(C.c_longlong)]
    out = C.c_longlong(-1)
    if L.pf_debug_fastdiv_mismatches(float(d), int(n), int(seed), C.byref(out)):
        raise PixelFlowError("pf_debug_fastdiv_mismatches failed")
    return out.value


def quot_mismatches(n: int = 1 << 24, seed: int = 1, exp_range: int = 60):
    """GPU self-check of the branch-free division of the fused SOR kernel (variant 6): (mismatches inside the guard -- must be 0,
    operand pairs outside the guard)."""
    L = load_library()
    L.pf_debug_quot_mismatches.argtypes = [C.c_longlong, C.c_ulonglong, C.c_int, C.POINTER(C.c_longlong),
                                           C.POINTER(C.c_longlong)]
    bad, outside = C.c_longlong(-1), C.c_longlong(-1)
    if L.pf_debug_quot_mismatches(int(n), int(seed), int(exp_range), C.byref(bad), C.byref(outside)):
        raise PixelFlowError("pf_debug_quot_mismatches failed")
    return bad.value, outside.value


def parse_porosity_csv(text: bytes, m: int, n: int, l: int = 0, threshold: float = 1.0e-6, out=None, device: int = -1):
    """Records of a porosity CSV (everything after the `m,n,l` header line) parsed on the GPU into the array
    [l+2][n+2][m+2] (2D, l = 0: [n+2][m+2]) like lib/grid.f90:281-294 does: porosity(x,y,z) = max(value, threshold).
    Returns (array, records stored)."""
    L = load_library()
    shape = (l + 2, n + 2, m + 2) if l > 0 else (n + 2, m + 2)
    out = np.zeros(shape) if out is None else out
    if tuple(out.shape) != shape:
        raise ValueError(f"array shape {out.shape} != {shape}")
    L.pf_parse_porosity_csv.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_double,
                                        C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]
    nrec = C.c_longlong(0)
    if L.pf_parse_porosity_csv(text, len(text), int(m), int(n), int(l), float(threshold), _dp(out), C.byref(nrec),
                               int(device)):
        raise PixelFlowError(L.pf_last_error(None).decode())
    return out, nrec.value


def comm_unique_id() -> bytes:
    """128-byte NCCL id (rank 0 calls this and broadcasts the bytes to the other ranks)."""
    buf = C.create_string_buffer(128)
    if load_library().pf_comm_unique_id(C.cast(buf, C.c_void_p)):
        raise PixelFlowError(load_library().pf_last_error(None).decode())
    return buf.raw


def _dp(a):
    if a is None:
        return None
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]):
        raise TypeError("expected a C-contiguous float64 numpy array")
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Solver:
    """One GPU's view of a PixelFlow run: mirrors the body of the reference's time loop.

    Parameters follow the reference's namelists (&physical, &grid_control, &porosity_control,
    &solver_control); dx,dy,dz,dt are what lib/grid.f90:297-300 derives from them.
    """

    def __init__(self, case, m, n, l=1, *, dx, dy, dz=1.0, dt, xnue, xlambda=0.0, density=1.0,
                 thickness=1.5, nonslip=True, iter_max=100, relux_factor=1.7, inlet_velocity=1.0,
                 outlet_pressure=0.0, AoA=0.0, wall=(1, 0, 0, 0, 2, 0), device=-1, rank=0, nranks=1,
                 nccl_unique_id: bytes | None = None, host_is_slab=False, sor_variant=0, use_graph=1,
                 halo_transport=0):
        L = load_library()
        cfg = PfConfig()
        L.pf_config_init(C.byref(cfg))
        cfg.solver_case = CASE_NAMES[case] if isinstance(case, str) else int(case)
        cfg.m, cfg.n, cfg.l = int(m), int(n), int(l)
        cfg.dx, cfg.dy, cfg.dz, cfg.dt = float(dx), float(dy), float(dz), float(dt)
        cfg.xnue, cfg.xlambda, cfg.density, cfg.thickness = float(xnue), float(xlambda), float(density), float(thickness)
        cfg.nonslip = 1 if nonslip else 0
        cfg.iter_max = int(iter_max)
        cfg.relux_factor = float(relux_factor)
        cfg.inlet_velocity, cfg.outlet_pressure, cfg.AoA = float(inlet_velocity), float(outlet_pressure), float(AoA)
        for i, wv in enumerate(wall):
            cfg.wall[i] = int(wv)
        cfg.device, cfg.rank, cfg.nranks = int(device), int(rank), int(nranks)
        cfg.host_is_slab = 1 if host_is_slab else 0
        cfg.sor_variant, cfg.use_graph = int(sor_variant), int(use_graph)
        cfg.halo_transport = int(halo_transport)
        self._uid = None
        if nranks > 1:
            if nccl_unique_id is None or len(nccl_unique_id) != 128:
                raise ValueError("nranks > 1 needs the 128-byte nccl_unique_id")
            self._uid = C.create_string_buffer(nccl_unique_id, 128)
            cfg.nccl_unique_id = C.cast(self._uid, C.c_void_p)
        self.cfg = cfg
        self.dim = 3 if cfg.solver_case >= IBM3_UNIFORM else 2
        self._L = L
        h = C.c_void_p()
        if L.pf_create(C.byref(h), C.byref(cfg)):
            raise PixelFlowError("pf_create: " + L.pf_last_error(None).decode())
        self._h = h
        kf, kc = C.c_int(), C.c_int()
        L.pf_local_slab(h, C.byref(kf), C.byref(kc))
        self.k_first, self.k_count = kf.value, kc.value
        if self.dim == 2:
            self.shape = (n + 2, m + 2)
        elif host_is_slab:
            self.shape = (self.k_count + 2, n + 2, m + 2)
        else:
            self.shape = (l + 2, n + 2, m + 2)

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc:
            raise PixelFlowError(f"{what}: {self._L.pf_last_error(self._h).decode()}")

    def _arr(self, a):
        if a is not None and tuple(a.shape) != self.shape:
            raise ValueError(f"array shape {a.shape} != {self.shape}")
        return _dp(a)

    def close(self):
        if getattr(self, "_h", None):
            self._L.pf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def zeros(self):
        return np.zeros(self.shape, dtype=np.float64)

    # -- data movement ------------------------------------------------------------------------
    def set_porosity(self, porosity):
        self._check(self._L.pf_set_porosity(self._h, self._arr(porosity)), "pf_set_porosity")

    def upload(self, u, v, w, p):
        self._check(self._L.pf_upload(self._h, self._arr(u), self._arr(v), self._arr(w), self._arr(p)), "pf_upload")

    def download(self, u=None, v=None, w=None, p=None):
        u = self.zeros() if u is None else u
        v = self.zeros() if v is None else v
        w = (self.zeros() if w is None else w) if self.dim == 3 else None
        p = self.zeros() if p is None else p
        self._check(self._L.pf_download(self._h, self._arr(u), self._arr(v), self._arr(w), self._arr(p)), "pf_download")
        return u, v, w, p

    def get_field(self, name, out=None):
        out = self.zeros() if out is None else out
        self._check(self._L.pf_get_field(self._h, FIELDS[name], self._arr(out)), "pf_get_field")
        return out

    def set_field(self, name, a):
        self._check(self._L.pf_set_field(self._h, FIELDS[name], self._arr(a)), "pf_set_field")

    # -- hot path ---------------------------------------------------------------------------------
    def step(self, nsteps=1):
        err = np.zeros(max(nsteps, 1))
        self._check(self._L.pf_step(self._h, int(nsteps), _dp(err)), "pf_step")
        return err[:nsteps]

    def step_host(self, nsteps, u, v, w, p):
        err = np.zeros(max(nsteps, 1))
        self._check(self._L.pf_step_host(self._h, int(nsteps), self._arr(u), self._arr(v), self._arr(w),
                                         self._arr(p), _dp(err)), "pf_step_host")
        return err[:nsteps]

    def initial_conditions(self):
        self._check(self._L.pf_initial_conditions(self._h), "pf_initial_conditions")

    def copy_old(self):
        self._check(self._L.pf_copy_old(self._h), "pf_copy_old")

    def divergence(self):
        self._check(self._L.pf_divergence(self._h), "pf_divergence")

    def predictor(self):
        self._check(self._L.pf_predictor(self._h), "pf_predictor")

    def build_poisson(self):
        self._check(self._L.pf_build_poisson(self._h), "pf_build_poisson")

    def sor(self, iters):
        err = np.zeros(1)
        self._check(self._L.pf_sor(self._h, int(iters), _dp(err)), "pf_sor")
        return float(err[0])

    def project(self):
        self._check(self._L.pf_project(self._h), "pf_project")

    def boundary(self):
        self._check(self._L.pf_boundary(self._h), "pf_boundary")

    def sync(self):
        self._check(self._L.pf_sync(self._h), "pf_sync")

    def force_log_2d(self, radius: float):
        """output_force_log_2d (lib/output.f90:244-305): dict with Fp, Fv, F (x,y pairs), Cd, Cl"""
        out = np.zeros(8)
        self._L.pf_force_log_2d.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
        self._check(self._L.pf_force_log_2d(self._h, float(radius), _dp(out)), "pf_force_log_2d")
        return {"Fp": (out[0], out[1]), "Fv": (out[2], out[3]), "F": (out[4], out[5]), "Cd": out[6], "Cl": out[7],
                "raw": out}

    def force_log_3d(self, radius: float):
        """output_force_log_3d (lib/output.f90:1090-1165): Fp, Fv, F (x,y,z triples), Cd(x), Cl, Cd(z)"""
        out = np.zeros(12)
        self._L.pf_force_log_3d.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_double)]
        self._check(self._L.pf_force_log_3d(self._h, float(radius), _dp(out)), "pf_force_log_3d")
        return {"Fp": tuple(out[0:3]), "Fv": tuple(out[3:6]), "F": tuple(out[6:9]), "Cdx": out[9], "Cl": out[10],
                "Cdz": out[11], "raw": out}

    VTK_SECTIONS = {"points": 0, "velocity": 1, "velocityInFluid": 2, "dimless_v": 3, "porosity": 4, "pressure": 5,
                    "VelocityDivergent": 6, "abs_dimless_v": 7}

    def vtk_section(self, section, xp, yp, zp=None, k_local0=1, nplanes=None) -> bytes:
        """body of one section of the reference's ASCII VTK snapshot (lib/output.f90:968-1088 / :421-537),
        formatted on the GPU with "(3(f16.4,1x))" for this rank's planes"""
        sec = self.VTK_SECTIONS[section] if isinstance(section, str) else int(section)
        if self.dim == 2:
            k_local0, nplanes = 0, 1
        elif nplanes is None:
            nplanes = self.k_count - k_local0 + 1
        L = self._L
        L.pf_vtk_section_bytes.restype = C.c_size_t
        L.pf_vtk_section_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int]
        dp = C.POINTER(C.c_double)
        L.pf_vtk_section.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, dp, dp, C.c_char_p]
        n = L.pf_vtk_section_bytes(self._h, sec, int(nplanes))
        buf = C.create_string_buffer(max(n, 1))
        xp, yp = np.ascontiguousarray(xp, dtype=np.float64), np.ascontiguousarray(yp, dtype=np.float64)
        zp = None if zp is None else np.ascontiguousarray(zp, dtype=np.float64)
        self._check(L.pf_vtk_section(self._h, sec, int(k_local0), int(nplanes), _dp(xp), _dp(yp), _dp(zp), buf),
                    "pf_vtk_section")
        return buf.raw[:n]

    @property
    def sor_variant(self) -> int:
        """the SOR kernel in use after auto-selection (1 half-sweeps, 3/4 fused, 6 fused + TMA; opt-in: 7 persistent
        half-sweeps, 8 temporally blocked 2D tiles)"""
        self._L.pf_get_sor_variant.argtypes = [C.c_void_p]
        return int(self._L.pf_get_sor_variant(self._h))

    @property
    def halo_transport(self) -> int:
        """slab-face transport of the fused SOR kernels: 0 single rank, 1 NCCL groups, 2 peer stores over NVLink +
        barrier kernel, 3 peer stores + handshake inside the TMA kernel"""
        self._L.pf_get_halo_transport.argtypes = [C.c_void_p]
        return int(self._L.pf_get_halo_transport(self._h))

    def last_timing(self):
        a, b, n = C.c_double(), C.c_double(), C.c_longlong()
        self._L.pf_last_timing(self._h, C.byref(a), C.byref(b), C.byref(n))
        return {"ms_total": a.value, "ms_sor": b.value, "launches": n.value}
