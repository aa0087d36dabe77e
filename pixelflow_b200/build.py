"""In-tree nvcc build of libpixelflow_gpu.so for sm_100a (and of the C++ twin drivers).

`python -m pixelflow_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU.
The .so stays inside the package directory so that it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libpixelflow_gpu.so")

SOURCES = ["pf_api.cu", "pf_kernels.cu", "pf_sor.cu", "pf_sor_fused.cu", "pf_sor_tma.cu", "pf_sor_persistent.cu", "pf_sor_tb2d.cu", "pf_comm.cu", "pf_ranks.cu", "pf_voxel.cu", "pf_stl.cu", "pf_output.cu", "pf_ingest.cu"]
DRIVER_NAMES = ["ibm2_uniform_omp", "ibm2_omp", "ibm2_drag_omp", "ibm2_backstep_omp", "ibm3_uniform_omp",
                "ibm3_omp", "ibm3_air_condition_omp"]

# -fmad=false: the reference (gfortran, baseline x86-64) has no fused multiply-add; parity is bit-exact
#              only if a*b+c stays two roundings.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "-Xptxas", "-v",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, "pf_internal.cuh"), os.path.join(CSRC, "pf_tma_common.cuh"), os.path.join(ROOT, "include", "pixelflow_gpu.h"),
                   os.path.abspath(__file__)]
    if not force and not _stale(LIB, deps):
        return LIB
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = ([nvcc()] + NVCC_FLAGS + os.environ.get("PF_NVCC_EXTRA", "").split() +
               ["-I", os.path.join(ROOT, "include"), "-c", s, "-o", o])
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for cmd, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                    "-Xcompiler", "-fPIC", "-ldl"]
    subprocess.check_call(link)
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


def build_drivers(force: bool = False) -> list[str]:
    """C++ twin drivers (pixelflow_b200/driver); built only if the sources exist."""
    drv = os.path.join(PKG, "driver")
    src = os.path.join(drv, "pixelflow_driver.cpp")
    if not os.path.exists(src):
        return []
    out = os.path.join(drv, "pixelflow_driver")
    if force or _stale(out, [src, LIB]):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-o", out,
                               "-L", PKG, "-lpixelflow_gpu", "-Wl,-rpath,$ORIGIN/..", "-ldl", "-lpthread", "-lrt"])
    # the reference's executable names (scripts/build/buildAll.sh:20-24, README.md:176-179)
    bindir = os.path.join(drv, "bin")
    os.makedirs(bindir, exist_ok=True)
    for name in DRIVER_NAMES:
        link = os.path.join(bindir, name)
        if not os.path.islink(link):
            if os.path.exists(link):
                os.remove(link)
            os.symlink(os.path.join("..", "pixelflow_driver"), link)
    return [out]


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    for d in build_drivers(force="--force" in sys.argv):
        print(d)
