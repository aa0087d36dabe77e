"""Python harness over the translated reference programs (oracle/_ref/<program>_<flavour>.so, built by
oracle/build_ref.py from the reference's own Fortran with oracle/f90toc.py).  TEST INFRASTRUCTURE ONLY.

A run is what a user of the reference does: a project directory with `config/controlDict.txt` and the porosity CSV
it names, then the program (`program main`, all of it: read_settings, the grid routine incl. the CSV read and the
porosity halos, initial_conditions, boundary, the time loop).  Only the output_* subroutines, get_now_time and
`call system` are stubs.  Afterwards the program's arrays are read straight out of its static storage.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from oracle import build_ref

PROGRAM_OF_CASE = {
    "ibm2_uniform": "ibm_2d_uniform_omp_cpu",
    "ibm2_backstep": "ibm_2d_backstep_omp_cpu",
    "ibm2_drag": "ibm_2d_drag_omp_cpu",
    "ibm3_uniform": "ibm_3d_uniform_omp_cpu",
    "ibm3_air_condition": "ibm_3d_air_condition_omp_cpu",
}


class RtVar(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ptr", C.c_void_p), ("type", C.c_char), ("rank", C.c_int),
                ("lo", C.c_int * 3), ("hi", C.c_int * 3)]


def have(case_or_program: str, flavour: str = "serial", size: str = "s") -> bool:
    program = PROGRAM_OF_CASE.get(case_or_program, case_or_program)
    return os.path.exists(build_ref.lib_path(program, flavour, size)) or build_ref.available()


CONTROLDICT_KEYS = {
    "physical": ("xnue", "xlambda", "density", "width", "height", "depth", "time", "inlet_velocity",
                 "outlet_pressure", "AoA"),
    "file_control": ("istep_out",),
    "grid_control": ("istep_max",),
    "porosity_control": ("thickness", "threshold", "radius", "center_x", "center_y", "center_z"),
    "calculation_method": ("nonslip",),
    "directory_control": ("output_folder", "csv_file"),
    "solver_control": ("iter_max", "relux_factor"),
}
DEFAULTS = dict(xnue=1e-3, xlambda=0.0, density=1.0, width=1.0, height=1.0, depth=1.0, time=1.0, inlet_velocity=1.0,
                outlet_pressure=0.0, AoA=0.0, istep_out=1000000, istep_max=100, thickness=1.5, threshold=1e-6,
                radius=0.1, center_x=0.5, center_y=0.5, center_z=0.5, nonslip=True, output_folder="output",
                csv_file="data/porosity.csv", iter_max=100, relux_factor=1.7)


def controldict_text(**settings) -> str:
    """the seven namelist groups in the order lib/global.f90:56-62 reads them; reals with 17 significant digits"""
    s = dict(DEFAULTS)
    for k in settings:
        if k not in s:
            raise KeyError(k)
    s.update(settings)
    out = []
    for grp, keys in CONTROLDICT_KEYS.items():
        out.append(f"&{grp}")
        for k in keys:
            v = s[k]
            if isinstance(v, bool):
                out.append(f"{k} = {'.true.' if v else '.false.'}")
            elif isinstance(v, (int, np.integer)) and k in ("istep_out", "istep_max", "iter_max"):
                out.append(f"{k} = {int(v)}")
            elif isinstance(v, str):
                out.append(f'{k} = "{v}"')
            else:
                out.append(f"{k} = {float(v)!r}")
        out.append("/")
    return "\n".join(out) + "\n"


_HELPER = None


def _helper_lib():
    """any built library of oracle/_ref: they all carry the runtime (ref_write_csv)"""
    global _HELPER
    if _HELPER is None:
        import glob
        libs = sorted(glob.glob(os.path.join(build_ref.OUT, "*_s_serial.so")))
        if not libs:
            return None
        _HELPER = C.CDLL(libs[0])
        _HELPER.ref_write_csv.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int]
        _HELPER.ref_write_csv.restype = C.c_int
    return _HELPER


def write_porosity_csv(path: str, eps: np.ndarray) -> None:
    """`m,n,l` then one record `i, j, k, value` per cell, i fastest (template/data/.porosity:1-3); values with 17
    significant digits so that the list-directed read returns exactly the array's doubles.  eps is [l,n,m] or [n,m]."""
    if eps.ndim == 2:
        eps = eps[None]
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    l, n, m = eps.shape
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    H = _helper_lib()
    if H is not None:
        if H.ref_write_csv(os.fsencode(path), eps.ctypes.data_as(C.POINTER(C.c_double)), m, n, l):
            raise OSError(f"cannot write {path}")
        return
    kk, jj, ii = np.meshgrid(np.arange(1, l + 1), np.arange(1, n + 1), np.arange(1, m + 1), indexing="ij")
    with open(path, "w") as f:
        f.write(f"{m},{n},{l}\n")
        for i, j, k, v in zip(ii.ravel(), jj.ravel(), kk.ravel(), eps.ravel()):
            f.write(f"{i}, {j}, {k}, {float(v)!r}\n")


def write_deck(dirpath: str, eps: np.ndarray, **settings) -> None:
    os.makedirs(os.path.join(dirpath, "config"), exist_ok=True)
    csv = settings.get("csv_file", DEFAULTS["csv_file"])
    with open(os.path.join(dirpath, "config", "controlDict.txt"), "w") as f:
        f.write(controldict_text(**settings))
    write_porosity_csv(os.path.join(dirpath, csv), eps)


class RefProgram:
    """one translated program; not re-entrant (the program's arrays are static, like the Fortran's).
    flavour "serial" / "omp": output_* are stubs; "gf": lib/output.f90 is translated as well and all I/O (open, list-directed and
    namelist reads, writes) is executed by libgfortran — the run directory then holds what a gfortran build of the reference leaves there
    (`stdout.log` = unit *, `etc/*.dat`, `<output_folder>/*.vtk`), byte for byte except the TIME stamps (stub)."""

    def __init__(self, case_or_program: str, flavour: str = "serial", size: str = "s", lib: str | None = None):
        """size "s": static bounds 160x160 / 72^3; "b": 2304x1500 / 260^3 (oracle/build_ref.py:BOUNDS);
        lib: an explicit library (build_ref.build_variant: other `parameter` values)"""
        self.program = PROGRAM_OF_CASE.get(case_or_program, case_or_program)
        self.flavour = flavour
        path = lib or build_ref.lib_path(self.program, flavour, size)
        if not os.path.exists(path):
            if build_ref.available():
                build_ref.build(programs=[self.program])
            else:
                raise FileNotFoundError(f"{path}: translated reference not built and /root/reference is absent")
        self.path = path
        self.L = None
        self.d3 = "_3d_" in self.program
        self._load()

    def _load(self):
        """bind the library.  A Fortran program starts from zero-initialised static storage (`-fno-automatic`) and
        the programs rely on it (halo corners, solver-local arrays); ref_run() therefore zeroes every static
        variable of the translated program before it calls `program main` (rt_reset_statics, generated)."""
        path = self.path
        L = C.CDLL(path)
        L.ref_run.argtypes = [C.c_char_p]
        L.ref_run.restype = C.c_int
        L.ref_lookup.argtypes = [C.c_char_p]
        L.ref_lookup.restype = C.POINTER(RtVar)
        L.ref_perr_count.restype = C.c_int
        L.ref_perr.argtypes = [C.c_int]
        L.ref_perr.restype = C.c_double
        L.ref_log.restype = C.c_char_p
        L.ref_error.restype = C.c_char_p
        L.ref_stub_count.argtypes = [C.c_char_p]
        L.ref_stub_count.restype = C.c_int
        L.ref_set_verbose.argtypes = [C.c_int]
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_max_threads.restype = C.c_int
        L.ref_use_libgfortran.argtypes = [C.c_char_p]
        L.ref_use_libgfortran.restype = C.c_int
        if self.flavour == "gf":
            # the output routines are translated too; their WRITE statements are executed by libgfortran itself
            from oracle import gfortran_rt
            path = gfortran_rt.find_libgfortran()
            if path is None:
                raise FileNotFoundError("libgfortran.so.5 not found (it ships inside numpy / scipy)")
            rc = L.ref_use_libgfortran(os.fsencode(path))
            if rc:
                raise OSError(f"ref_use_libgfortran({path}) -> {rc}")
        L.ref_set_step_limit.argtypes = [C.c_int]
        L.ref_step_count.restype = C.c_int
        L.ref_step_time.argtypes = [C.c_int]
        L.ref_step_time.restype = C.c_double
        L.ref_end_time.restype = C.c_double
        self.L = L

    def run(self, deck_dir: str, verbose: bool = False, step_limit: int = 0) -> np.ndarray:
        """run `program main` in deck_dir; returns the per-step 'p error' values the program logged.
        step_limit=n leaves the time loop after n steps (an unmodified shipped deck would run 2000-5000)."""
        self.L.ref_set_verbose(int(verbose))
        self.L.ref_set_step_limit(int(step_limit))
        rc = self.L.ref_run(os.fsencode(deck_dir))
        if rc != 0:
            raise RuntimeError(f"{self.program}: {self.L.ref_error().decode()}")
        return np.array([self.L.ref_perr(i) for i in range(self.L.ref_perr_count())])

    def call(self, subroutine: str, *args):
        """call a translated subroutine directly (Fortran convention: everything by reference).  An argument is the
        name of a program / module variable (str) or a Python int / float / bool passed through a temporary."""
        fn = getattr(self.L, "f_" + subroutine)
        fn.restype = None
        keep, cargs = [], []
        for a in args:
            if isinstance(a, str):
                v = self.L.ref_lookup(a.encode())
                if not v:
                    raise KeyError(a)
                cargs.append(C.c_void_p(v.contents.ptr))
            elif isinstance(a, (bool, int, np.integer)):
                keep.append(C.c_int(int(a)))
                cargs.append(C.byref(keep[-1]))
            else:
                keep.append(C.c_double(float(a)))
                cargs.append(C.byref(keep[-1]))
        fn(*cargs)

    def set_threads(self, n: int) -> int:
        """OpenMP flavour: omp_set_num_threads(n); returns omp_get_max_threads() (1 for the serial flavours)"""
        self.L.ref_set_threads(int(n))
        return int(self.L.ref_max_threads())

    def step_seconds(self) -> np.ndarray:
        """wall seconds of each completed time step of the last run (from one '--- time_steps=' line to the next;
        the last one up to the program's return, which then includes the stubbed end-of-run output calls)"""
        n = self.L.ref_step_count()
        t = np.array([self.L.ref_step_time(i) for i in range(n)] + [self.L.ref_end_time()])
        d = np.diff(t)
        return d[d > 0] if n else d

    def log(self) -> str:
        return self.L.ref_log().decode()

    def stub_count(self, name: str) -> int:
        return self.L.ref_stub_count(name.encode())

    def scalar(self, name: str):
        v = self.L.ref_lookup(name.encode())
        if not v:
            raise KeyError(name)
        v = v.contents
        if v.type == b"d":
            return C.cast(v.ptr, C.POINTER(C.c_double))[0]
        if v.type == b"f":
            return C.cast(v.ptr, C.POINTER(C.c_float))[0]
        if v.type in (b"i", b"l"):
            return C.cast(v.ptr, C.POINTER(C.c_int))[0]
        raise TypeError(name)

    def array(self, name: str) -> np.ndarray:
        """the used part (0..m+1, 0..n+1[, 0..l+1]) of a field array, as a C-ordered [k,j,i] copy"""
        v = self.L.ref_lookup(name.encode())
        if not v:
            raise KeyError(name)
        v = v.contents
        ext = [v.hi[d] - v.lo[d] + 1 for d in range(v.rank)]
        n = int(np.prod(ext))
        ct = C.c_float if v.type == b"f" else C.c_double
        flat = np.ctypeslib.as_array(C.cast(v.ptr, C.POINTER(ct)), shape=(n,))
        full = flat.reshape(ext[::-1])     # column-major (i fastest) == C order [k,j,i]
        m, nn = self.scalar("m"), self.scalar("n")
        if v.rank == 3:
            l = self.scalar("l")
            return full[:l + 2, :nn + 2, :m + 2].copy()
        if v.rank == 2:
            return full[:nn + 2, :m + 2].copy()
        return full.copy()

    def fields(self) -> dict:
        names = ("u", "v", "w", "p", "porosity") if self.d3 else ("u", "v", "p", "porosity")
        return {k: self.array(k) for k in names}
