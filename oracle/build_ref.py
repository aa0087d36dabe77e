"""Build the translated reference: /root/reference/src/omp_parallel/*.f90  --f90toc-->  oracle/_ref/*.c  --gcc-->
oracle/_ref/*.so.  TEST INFRASTRUCTURE.

The Fortran sources are read where they lie (they are never copied into the repository); everything generated lands in
`oracle/_ref/`, which is git-ignored but travels to the GPU box with the gpurun snapshot.  On a machine without
/root/reference (the GPU box) nothing can be rebuilt; the prebuilt libraries are used as they are.

One library per (program, flavour):
    <program>_serial.so   `!$omp` directives ignored — the deterministic run the parity tests compare with
    <program>_omp.so      `!$omp` directives translated to `#pragma omp` — the timed CPU baseline ("reference" kind)
Compile flags follow scripts/build/build_omp.sh:38 (`-O3 -fopenmp -fno-automatic -mcmodel=medium`), plus
`-ffp-contract=off` (baseline x86-64 gfortran has no FMA) — `-fno-automatic` is what the translator's `static`
locals are; `-mcmodel=medium` is needed as soon as the static arrays pass 2 GB.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_SRC = os.environ.get("PF_REFERENCE_SRC", "/root/reference/src/omp_parallel")

# program -> grid routine it calls, extra library routines to translate.
PROGRAMS = {
    "ibm_2d_uniform_omp_cpu": dict(grid="grid_conditions", extra=()),
    "ibm_2d_backstep_omp_cpu": dict(grid="grid_conditions", extra=()),
    "ibm_2d_drag_omp_cpu": dict(grid="grid_conditions", extra=("output_force_log_2d",)),
    # output_force_log_3d is defined in lib/output.f90 but called by none of the programs; the harness calls it
    "ibm_3d_uniform_omp_cpu": dict(grid="grid_conditions_yz_periodic", extra=("output_force_log_3d",)),
    "ibm_3d_air_condition_omp_cpu": dict(grid="grid_conditions_wall", extra=()),
}
# Static bounds md, nd, ld (lib/global.f90:6,12; a grid needs m+1 <= md ...).  The shipped values (2D 1500x1500,
# 3D 256x180x80) are too small for backstep (2251x411) and for 256^3, so they are set here — the only
# edit to the reference's text, a parameter value.  "s": small grids and the 64^3 room deck (a run zeroes all static
# storage first, so small bounds keep the tests fast); "b": the shipped 2D decks and the 256^3 bench sample.
BOUNDS = {
    "s": {2: dict(md=160, nd=160), 3: dict(md=72, nd=72, ld=72)},
    "b": {2: dict(md=2304, nd=1500), 3: dict(md=260, nd=260, ld=260)},
}
FLAVOURS = ("serial", "omp", "gf", "r4")
# "r4": serial, default `real` = 32 bit — the reference exactly as shipped (no -fdefault-real-8 anywhere in its build
# files).  Not a parity target (the north star fixes fp64); it measures how far the shipped precision is
# from the fp64 evaluation.
# "gf": serial, and lib/output.f90's routines are translated too instead of being stubs; their formatted writes need
# the libgfortran backend of the runtime (ref_translated.RefProgram(..., "gf")).  Small bounds only (the VTK files of
# the shipped 2D decks are 100 MB each).
OUTPUT_ROUTINES = {
    2: ("output_grid_2d", "output_solution_post_2d", "output_paraview_2d", "output_divergent_2d",
        "output_paraview_temp_2d", "output_force_log_2d"),
    3: ("output_grid_3d", "output_solution_post_3d", "output_paraview_3d", "output_divergent_3d",
        "output_paraview_temp_3d", "output_force_log_3d"),
}
CFLAGS = ["-O3", "-ffp-contract=off", "-fPIC", "-shared", "-mcmodel=medium", "-fno-strict-aliasing"]


def dim_of(program: str) -> int:
    return 3 if "_3d_" in program else 2


def available() -> bool:
    return os.path.isdir(REF_SRC)


def lib_path(program: str, flavour: str = "serial", size: str = "s") -> str:
    return os.path.join(OUT, f"{program}_{size}_{flavour}.so")


def generate(program: str, flavour: str, size: str = "s", extra_overrides: dict | None = None) -> str:
    from oracle import f90toc
    cfg = PROGRAMS[program]
    files = [
        (os.path.join(REF_SRC, "lib", "global.f90"), None, None),
        (os.path.join(REF_SRC, "lib", "grid.f90"), {cfg["grid"]}, None),
    ]
    if flavour == "gf":
        files.append((os.path.join(REF_SRC, "lib", "output.f90"), set(OUTPUT_ROUTINES[dim_of(program)]), None))
    elif cfg["extra"]:
        files.append((os.path.join(REF_SRC, "lib", "output.f90"), set(cfg["extra"]), None))
    files.append((os.path.join(REF_SRC, program + ".f90"), None, None))
    overrides = dict(BOUNDS[size][dim_of(program)])
    overrides.update(extra_overrides or {})
    return f90toc.translate(files, omp=(flavour == "omp"), overrides=overrides,
                            real_kind=4 if flavour == "r4" else 8)


def build(force: bool = False, programs=None, verbose: bool = False) -> list[str]:
    """translate + compile; returns the libraries that exist afterwards"""
    every = [(p, z, f) for p in PROGRAMS for z in BOUNDS for f in FLAVOURS if not (f in ("gf", "r4") and z != "s")]
    if not available():
        return [lib_path(p, f, z) for p, z, f in every if os.path.exists(lib_path(p, f, z))]
    os.makedirs(OUT, exist_ok=True)
    deps = [os.path.join(HERE, n) for n in ("f90toc.py", "ref_runtime.c", "ref_runtime.h", "build_ref.py")]
    procs, libs = [], []
    for prog, size, flav in every:
        if programs and prog not in programs:
            continue
        srcs = [os.path.join(REF_SRC, prog + ".f90")] + [os.path.join(REF_SRC, "lib", n) for n in ("global.f90", "grid.f90", "output.f90")]
        lib = lib_path(prog, flav, size)
        libs.append(lib)
        if not force and os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps + srcs):
            continue
        csrc = lib[:-3] + ".c"
        with open(csrc, "w") as f:
            f.write(f"/* GENERATED by oracle/f90toc.py from {REF_SRC}/{prog}.f90 (+ lib/global.f90, lib/grid.f90) — not committed */\n")
            f.write(generate(prog, flav, size))
        cmd = ["gcc"] + CFLAGS + (["-fopenmp"] if flav == "omp" else []) + \
              ["-I", HERE, csrc, os.path.join(HERE, "ref_runtime.c"), "-o", lib, "-lm", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("gcc failed: " + " ".join(cmd))
        if verbose and out.strip():
            print(out)
    return libs


def build_variant(program: str, tag: str, extra_overrides: dict, flavour: str = "serial", size: str = "s",
                  cflags: list | None = None) -> str:
    """the same program with other `parameter` values — e.g. module wall_conditions of the air-condition program
    (top_wall ... north_wall are compile-time parameters in the reference: a user edits them and rebuilds).
    Returns oracle/_ref/<program>_<size>_<flavour>_<tag>.so"""
    lib = os.path.join(OUT, f"{program}_{size}_{flavour}_{tag}.so")
    deps = [os.path.join(HERE, n) for n in ("f90toc.py", "ref_runtime.c", "ref_runtime.h", "build_ref.py")]
    if os.path.exists(lib) and (not available() or all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps)):
        return lib
    if not available():
        raise FileNotFoundError(f"{lib}: not built and /root/reference is absent")
    os.makedirs(OUT, exist_ok=True)
    csrc = lib[:-3] + ".c"
    with open(csrc, "w") as f:
        f.write(f"/* GENERATED by oracle/f90toc.py from {REF_SRC}/{program}.f90, parameters {extra_overrides} — not committed */\n")
        f.write(generate(program, flavour, size, extra_overrides))
    subprocess.check_call(["gcc"] + (cflags or CFLAGS) + (["-fopenmp"] if flavour == "omp" else []) +
                          ["-I", HERE, csrc, os.path.join(HERE, "ref_runtime.c"), "-o", lib, "-lm", "-ldl"])
    return lib


FORTRAN_DIR = os.path.join(os.path.dirname(HERE), "pixelflow_b200", "fortran")


# the product's Fortran drivers: name -> (source file, grid routine of the reference it calls, dimension)
FORTRAN_DRIVERS = {
    "ibm3_uniform": ("ibm3_uniform_gpu.f90", "grid_conditions_yz_periodic", 3),
    "ibm3_air_condition": ("ibm3_air_condition_gpu.f90", "grid_conditions_wall", 3),
    "ibm2_uniform": ("ibm2_uniform_gpu.f90", "grid_conditions", 2),
    "ibm2_backstep": ("ibm2_backstep_gpu.f90", "grid_conditions", 2),
    "ibm2_drag": ("ibm2_drag_gpu.f90", "grid_conditions", 2),
}


def build_fortran_driver(backend: str = "double", case: str = "ibm3_uniform") -> str:
    """A Fortran driver of the product (pixelflow_b200/fortran/<case>_gpu.f90, `use pixelflow_gpu`), translated with
    the reference's support library (lib/global.f90, lib/grid.f90, lib/output.f90) and linked against
      backend "double": oracle/_ref/libpf_abi_double.so — the CPU test double of the C ABI (oracle/abi_double.c)
      backend "gpu":    pixelflow_b200/libpixelflow_gpu.so — the product library (needs a GPU at run time)
    -> oracle/_ref/fdriver_<case>_<backend>.so, run with ref_translated.RefProgram(..., "gf", lib=...)."""
    from oracle import f90_cmodule, f90toc
    fname, grid, dim = FORTRAN_DRIVERS[case]
    lib = os.path.join(OUT, f"fdriver_{case}_{backend}.so")
    mod = os.path.join(FORTRAN_DIR, "pixelflow_gpu_mod.f90")
    drv = os.path.join(FORTRAN_DIR, fname)
    deps = [os.path.join(HERE, n) for n in ("f90toc.py", "f90_cmodule.py", "ref_runtime.c", "ref_runtime.h",
                                            "build_ref.py", "fortran_helpers.c", "abi_double.c", "pf_oracle.c")] + [mod, drv]
    if os.path.exists(lib) and (not available() or all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps)):
        return lib
    if not available():
        raise FileNotFoundError(f"{lib}: not built and /root/reference is absent")
    os.makedirs(OUT, exist_ok=True)
    root = os.path.dirname(HERE)
    if backend == "double":
        from oracle import oracle_c
        oracle_c.build()
        dbl = os.path.join(OUT, "libpf_abi_double.so")
        if not os.path.exists(dbl) or any(os.path.getmtime(d) > os.path.getmtime(dbl) for d in deps[6:8]):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-I",
                                   os.path.join(root, "include"), os.path.join(HERE, "abi_double.c"), "-o", dbl,
                                   "-L", HERE, "-loracle", "-Wl,-rpath,$ORIGIN/.."])
        link = ["-L", OUT, "-lpf_abi_double", "-Wl,-rpath,$ORIGIN"]
    else:
        pkg = os.path.join(root, "pixelflow_b200")
        link = ["-L", pkg, "-lpixelflow_gpu", "-Wl,-rpath,$ORIGIN/../../pixelflow_b200"]
    tr = f90toc.Translator(overrides=BOUNDS["s"][dim])
    tr.add_c_module(f90_cmodule.describe(mod))
    routines = set(OUTPUT_ROUTINES[dim]) - {"output_force_log_2d", "output_force_log_3d"}
    for path, only in ((os.path.join(REF_SRC, "lib", "global.f90"), None),
                       (os.path.join(REF_SRC, "lib", "grid.f90"), {grid}),
                       (os.path.join(REF_SRC, "lib", "output.f90"), routines),
                       (drv, None)):
        with open(path) as f:
            tr.add_source(f.read(), path, only, None)
    csrc = lib[:-3] + ".c"
    with open(csrc, "w") as f:
        f.write(f"/* GENERATED by oracle/f90toc.py from {drv} (+ the reference's lib/*.f90) — not committed */\n")
        f.write(tr.emit())
    subprocess.check_call(["gcc"] + CFLAGS + ["-I", HERE, csrc, os.path.join(HERE, "ref_runtime.c"),
                                              os.path.join(HERE, "fortran_helpers.c"), "-o", lib, "-lm", "-ldl"] + link)
    return lib


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(HERE))
    for lib in build(force="--force" in sys.argv, verbose=True):
        print(lib)
