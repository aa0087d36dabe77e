"""Reader for PixelFlow's config/controlDict.txt (seven Fortran namelist groups).

Mirrors `read_settings`, src/omp_parallel/lib/global.f90:28-63: groups &physical, &file_control,
&grid_control, &porosity_control, &calculation_method, &directory_control, &solver_control, read in
that order from one unit.  `!` starts a comment; numbers are parsed with Python's float() (correctly
rounded decimal -> double, like gfortran's list-directed read at -fdefault-real-8).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

GROUP_ORDER = ["physical", "file_control", "grid_control", "porosity_control", "calculation_method",
               "directory_control", "solver_control"]

_FLOAT_KEYS = {"xnue", "xlambda", "density", "width", "height", "depth", "time", "inlet_velocity",
               "outlet_pressure", "aoa", "thickness", "threshold", "radius", "center_x", "center_y",
               "center_z", "relux_factor"}
_INT_KEYS = {"istep_out", "istep_max", "iter_max"}
_BOOL_KEYS = {"nonslip"}
_STR_KEYS = {"output_folder", "csv_file"}


@dataclass
class ControlDict:
    xnue: float = 0.0
    xlambda: float = 0.0
    density: float = 0.0
    width: float = 0.0
    height: float = 0.0
    depth: float = 0.0
    time: float = 0.0
    inlet_velocity: float = 0.0
    outlet_pressure: float = 0.0
    AoA: float = 0.0
    istep_out: int = 0
    istep_max: int = 0
    thickness: float = 0.0
    threshold: float = 0.0
    radius: float = 0.0
    center_x: float = 0.0
    center_y: float = 0.0
    center_z: float = 0.0
    nonslip: bool = False
    output_folder: str = ""
    csv_file: str = ""
    iter_max: int = 0
    relux_factor: float = 0.0
    groups_seen: list = field(default_factory=list)


def _fortran_float(tok: str) -> float:
    return float(tok.strip().lower().replace("d", "e"))


def _fortran_bool(tok: str) -> bool:
    t = tok.strip().lower().strip(".")
    if t.startswith("t"):
        return True
    if t.startswith("f"):
        return False
    raise ValueError(f"bad logical {tok!r}")


def parse_controldict(text: str) -> ControlDict:
    cd = ControlDict()
    group = None
    for raw in text.splitlines():
        line = raw
        # strip comments that are outside quotes
        out, q = [], None
        for ch in line:
            if q:
                out.append(ch)
                if ch == q:
                    q = None
            elif ch in "\"'":
                q = ch
                out.append(ch)
            elif ch == "!":
                break
            else:
                out.append(ch)
        line = "".join(out).strip()
        if not line:
            continue
        if line.startswith("&"):
            group = line[1:].split()[0].lower()
            cd.groups_seen.append(group)
            line = line[1 + len(group):].strip()
            if not line:
                continue
        if line.startswith("/"):
            group = None
            continue
        if group is None:
            continue
        for key, val in re.findall(r"([A-Za-z_][A-Za-z_0-9]*)\s*=\s*(\"[^\"]*\"|'[^']*'|[^,\s/]+)", line):
            k = key.lower()
            if k in _FLOAT_KEYS:
                setattr(cd, "AoA" if k == "aoa" else k, _fortran_float(val))
            elif k in _INT_KEYS:
                setattr(cd, k, int(val))
            elif k in _BOOL_KEYS:
                setattr(cd, k, _fortran_bool(val))
            elif k in _STR_KEYS:
                setattr(cd, k, val.strip("\"'")[:50])
            else:
                raise KeyError(f"unknown namelist variable {key!r} in &{group}")
        if line.endswith("/"):
            group = None
    # the reference reads the groups in a fixed order from one unit: a group that appears before an
    # earlier-listed one would not be found (SURVEY.md 5 "config / flags")
    order = [g for g in cd.groups_seen if g in GROUP_ORDER]
    if order != sorted(order, key=GROUP_ORDER.index):
        raise ValueError("namelist groups out of order: " + ", ".join(order))
    return cd
