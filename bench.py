#!/usr/bin/env python
"""bench.py -- throughput of the PixelFlow per-timestep hot path on B200 (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload s1|s2|s3] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one whole time step of the ibm3 (3D uniform) solver -- momentum predictor, Poisson
source, iter_max=100 red-black SOR iterations with the residual, projection, boundary conditions --
over the synthetic porous channel.  Default workload = BASELINE.json configs[4]
(1024x512x512, fp64), z-slab sharded over the N ranks (total work fixed -> "strong" scaling).

Printed (rank 0, one JSON line): metric = cell-updates/s of the whole job with all inputs resident
in HBM; `e2e` = the same metric through pf_step_host with pinned HOST buffers (H2D of u,v,w,p and
D2H of u,v,w,p inside the timed region, every step); `roofline` for the SOR half-sweep kernel;
`cpu_baseline` = the reference's own OpenMP program on this box's host cores on a bounded sample: its
Fortran source machine-translated to C (oracle/f90toc.py -> oracle/_ref, kind "reference"; there is no
Fortran compiler in the image) or, if that is not built, the hand-written restatement (kind "port").
`--impl reference` times only that CPU arm, on the same workload/metric.
At N=1 the line also carries, measured after and outside every timed region, each in its own process with a timeout
(tools/decks_probe.py): `decks` = the reference's three shipped decks (BASELINE configs[0..2]) — bit-identical to the
reference's own outputs or not, and ms/step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pixelflow_b200 import workloads as wl  # noqa: E402

ALGO_BYTES_PER_CELL_SWEEP = 88.0   # 8 coefficient doubles + p read twice + p written once
ALGO_BYTES_PER_CELL_STEP_FIXED = 208.0
# What each kernel's OWN data layout has to move per cell per red+black sweep (its algorithmic bytes; tile-ring
# overlap and halo traffic excluded): the half-sweep kernels stream the survey's eight per-cell arrays (88 B); the
# fused kernels store three face-coefficient arrays instead of seven per-cell ones (aw(i) = ae(i-1) ... hold bit for
# bit) and touch p once per iteration: bb + cx, cy, cz + p in + p out = 48 B; variant 2 rebuilds the coefficients from
# the porosity: bb + eps + p twice + p out = 48 B.
KERNEL_BYTES_PER_CELL_SWEEP = {1: 88.0, 5: 88.0, 7: 88.0, 8: 88.0, 2: 48.0, 3: 48.0, 4: 48.0, 6: 48.0}
FALLBACK_HBM_GBS = 6650.0          # B200_PROFILING.md fallback, used only if MEASURED_PEAKS.json is absent

WORKLOAD_ALIASES = {"s1": "s1_1024x512x512", "s2": "s2_256", "s3": "s3_64", "dragon": "dragon_256", "s4": "dragon_256",
                    "dragon_stl": "dragon_stl_256"}
KERNEL_NAMES = {1: "sor_sweep_kernel (one colour half-sweep per launch)",
                5: "sor_sweep_kernel (one colour half-sweep per launch)",
                2: "sor_sweep_eps_kernel (one colour half-sweep per launch)",
                3: "sor_fused_kernel<32,16> (red+black iteration per launch)",
                4: "sor_fused_kernel<32,8> (red+black iteration per launch)",
                6: "sor_tma_kernel (red+black iteration per launch, TMA pipeline)",
                7: "sor_persistent_kernel (all half-sweeps of a solve in one cooperative launch)",
                8: "sor_tb2d_kernel (2D: four red-black iterations per launch on shared-memory tiles)"}
CPU_SAMPLE = (256, 128, 128)       # sub-block of the workload the CPU restatement is timed on


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="s1")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--iter-max", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary 256^3 measurement at N=1")
    ap.add_argument("--no-decks", action="store_true", help="skip the shipped-deck probe (tools/decks_probe.py)")
    ap.add_argument("--no-parity", action="store_true", help="skip the field hash and the variant-1 cross-check")
    ap.add_argument("--sor-variant", type=int, default=0)
    ap.add_argument("--use-graph", type=int, default=1)
    ap.add_argument("--halo-transport", type=int, default=0,
                    help="z-slab ranks with a fused SOR kernel: 0 auto, 1 NCCL groups, 2 peer stores over NVLink + barrier kernel, "
                         "3 peer stores + handshake inside the TMA kernel")
    return ap.parse_args()


def workload_params(name, iter_max):
    name = WORKLOAD_ALIASES.get(name, name)
    m, n, l, width, height, depth = wl.WORKLOADS[name]
    ph = dict(wl.CHANNEL_PHYSICS)
    ph["iter_max"] = iter_max
    dx, dy, dz, dt = wl.grid_spacing(width, height, depth, ph["time"], ph["istep_max"], m, n, l)
    return name, (m, n, l), dict(dx=dx, dy=dy, dz=dz, dt=dt, xnue=ph["xnue"], xlambda=ph["xlambda"],
                                 density=ph["density"], thickness=ph["thickness"], nonslip=ph["nonslip"],
                                 iter_max=ph["iter_max"], relux_factor=ph["relux_factor"],
                                 inlet_velocity=ph["inlet_velocity"], outlet_pressure=ph["outlet_pressure"],
                                 AoA=ph["AoA"])


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index=0):
        self.dev = device_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            t = [x.strip() for x in line.split(",")]
            if len(t) < 9 or t[0] != str(self.dev):
                continue
            try:
                sm.append(float(t[1])); smax.append(float(t[2])); power.append(float(t[3]))
            except ValueError:
                continue
            for nm, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU restatement (the reference arm / cpu_baseline)
# ------------------------------------------------------------------------------------------------
def _cpu_model():
    try:
        return [ln.split(":", 1)[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")][0]
    except Exception:
        return "unknown"


def cpu_reference_run(workload, iter_max, steps, warmup, full=False):
    """Times the reference's own CPU implementation of the path with all host threads: on a bounded sub-block of the
    workload (CPU_SAMPLE), or -- full=True, for workloads inside the translated program's static bounds (256^3) -- on
    the WHOLE workload, the same configuration the GPU arm runs.  Returns (cell_updates_per_s, ms_per_step, info).

    kind "reference": `program main` of src/omp_parallel/ibm_3d_uniform_omp_cpu.f90, machine-translated to C with its
    `!$omp` directives (oracle/f90toc.py -> oracle/_ref/*_b_omp.so, built where /root/reference exists and shipped
    prebuilt to the GPU box; there is no Fortran compiler in the image), run on a project directory like the
    reference is; steps are timed between its own '--- time_steps=' log lines.
    kind "port": the hand-written restatement oracle/pf_oracle.c — used only if oracle/_ref is missing."""
    if FULL_AFFINITY:   # the GPU arm may have bound this process to one NUMA node: the CPU arm gets every core back
        os.sched_setaffinity(0, FULL_AFFINITY)
    name, (m, n, l), kw = workload_params(workload, iter_max)
    sm, sn, sl = (m, n, l) if full else (min(CPU_SAMPLE[0], m), min(CPU_SAMPLE[1], n), min(CPU_SAMPLE[2], l))
    same = (sm, sn, sl) == (m, n, l)
    where = f"the whole {name} grid ({sm}x{sn}x{sl})" if same else f"the {sm}x{sn}x{sl} leading sub-block of {name}"
    cores = os.cpu_count() or 1
    # all host threads: torchrun exports OMP_NUM_THREADS=1 to its workers, and in the reference arm rank 0 works alone
    threads = int(os.environ.get("PF_CPU_THREADS", cores))
    os.environ["OMP_NUM_THREADS"] = str(threads)
    cells = sm * sn * sl
    eps = wl.porous_channel(sm, sn, sl)
    from oracle import build_ref, ref_translated  # bench.py touches oracle/ only here: as the measured CPU baseline
    prog = "ibm_3d_uniform_omp_cpu"
    if os.path.exists(build_ref.lib_path(prog, "omp", "b")) and max(sm, sn, sl) < build_ref.BOUNDS["b"][3]["md"]:
        ph = dict(wl.CHANNEL_PHYSICS)
        R = ref_translated.RefProgram(prog, "omp", "b")
        threads = R.set_threads(threads)
        with tempfile.TemporaryDirectory() as d:
            # same dx, dy, dz, dt as the full workload: width = dx*(m-1) ..., time/istep_max unchanged
            ref_translated.write_deck(
                d, eps[1:-1, 1:-1, 1:-1], xnue=ph["xnue"], xlambda=ph["xlambda"], density=ph["density"],
                width=kw["dx"] * (sm - 1), height=kw["dy"] * (sn - 1), depth=kw["dz"] * (sl - 1), time=ph["time"],
                istep_max=ph["istep_max"], inlet_velocity=ph["inlet_velocity"], outlet_pressure=ph["outlet_pressure"],
                AoA=ph["AoA"], thickness=ph["thickness"], nonslip=bool(ph["nonslip"]), iter_max=iter_max,
                relux_factor=ph["relux_factor"])
            R.run(d, step_limit=warmup + steps)
        secs = R.step_seconds()
        if len(secs) != warmup + steps:
            raise RuntimeError(f"translated reference ran {len(secs)} steps, expected {warmup + steps}")
        dt = float(secs[warmup:].sum())
        info = {"kind": "reference", "cores": threads, "cpu": _cpu_model(),
                "same_config": same, "warmup_steps": warmup,
                "sample": f"{steps} step(s) after {warmup} warm-up x iter_max={iter_max} on {where}: "
                          "the reference's src/omp_parallel/ibm_3d_uniform_omp_cpu.f90 (program main, OpenMP) "
                          "machine-translated to C (oracle/f90toc.py; no Fortran compiler in the image), "
                          "gcc -O3 -fopenmp -ffp-contract=off"}
        return cells * steps / dt, dt / steps * 1e3, info
    from oracle import oracle_c
    oracle_c.build()
    P = oracle_c.make_params(m=sm, n=sn, l=sl, **{k: v for k, v in kw.items()})
    oc = oracle_c.Oracle3D(P, False, eps[1:-1, 1:-1, 1:-1])
    oc.initialise()
    if warmup > 0:
        oc.step(warmup)
    t0 = time.perf_counter()
    oc.step(steps)
    dt = time.perf_counter() - t0
    info = {"kind": "port", "cores": threads, "cpu": _cpu_model(),
            "same_config": same, "warmup_steps": warmup,
            "sample": f"{steps} step(s) after {warmup} warm-up x iter_max={iter_max} on {where} "
                      "(restated reference, C/OpenMP; oracle/_ref not built)"}
    return cells * steps / dt, dt / steps * 1e3, info


def run_decks_probe(sor_variant, timeout_s):
    """tools/decks_probe.py in a subprocess; a failure or a timeout is reported, never raised"""
    cmd = [sys.executable, os.path.join(ROOT, "tools", "decks_probe.py"), "--sor-variant", str(sor_variant)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return {"error": f"timeout after {timeout_s} s"}
    except Exception as e:   # noqa: BLE001
        return {"error": repr(e)}
    rows = []
    for ln in r.stdout.splitlines():
        if ln.startswith("{"):
            try:
                rows.append(json.loads(ln))
            except ValueError:
                pass
    if r.returncode != 0:
        return {"error": f"exit code {r.returncode}", "stderr": r.stderr[-400:], "rows": rows}
    return rows


# ------------------------------------------------------------------------------------------------
# parity key: a checksum of checksums that does not depend on how the grid is cut into z-slabs
# ------------------------------------------------------------------------------------------------
def plane_digests(fields, k_first, k_count, rank, nranks):
    """{global plane k: SHA-256 over that plane of u, v, w, p (x/y halos included)} for the planes this rank OWNS
    (k_first .. k_first+k_count-1) plus the global ghost planes 0 (rank 0) and l+1 (last rank).  `fields` are slab
    arrays [k_count+2][n+2][m+2]; the inner ghost planes (copies of a neighbour's cells) are not hashed."""
    import hashlib
    from concurrent.futures import ThreadPoolExecutor
    lo = 0 if rank == 0 else 1
    hi = k_count + 1 if rank == nranks - 1 else k_count

    def one(kl):
        h = hashlib.sha256()
        for a in fields:
            h.update(np.ascontiguousarray(a[kl]).data)
        return k_first - 1 + kl, h.hexdigest()

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:   # hashlib releases the GIL
        return dict(ex.map(one, range(lo, hi + 1)))


def combine_plane_digests(parts, l):
    """SHA-256 over the per-plane digests in global plane order 0 .. l+1; every plane must come from exactly one rank"""
    import hashlib
    merged = {}
    for d in parts:
        for k, v in d.items():
            k = int(k)
            if k in merged:
                raise ValueError(f"plane {k} hashed by two ranks")
            merged[k] = v
    if sorted(merged) != list(range(l + 2)):
        raise ValueError("plane digests do not cover the grid")
    h = hashlib.sha256()
    for k in range(l + 2):
        h.update(bytes.fromhex(merged[k]))
    return h.hexdigest()


def parity_record(s, errs, l, rank, nranks, dist, steps_from_init):
    """the fields after `steps_from_init` steps from the initial conditions, as one hash, and the logged p errors of
    the timed steps -- equal across N = 1, 2, 4, 8 and across SOR kernels iff the paths are bit-identical"""
    u, v, w, p = s.download()
    mine = plane_digests([u, v, w, p], s.k_first, s.k_count, rank, nranks)
    del u, v, w, p
    if dist is not None:
        parts = [None] * nranks
        dist.all_gather_object(parts, mine)
    else:
        parts = [mine]
    if rank != 0:
        return None
    import hashlib
    errs = np.ascontiguousarray(errs, dtype=np.float64)
    return {"fields_sha256": combine_plane_digests(parts, l), "what": "SHA-256 over the per-plane SHA-256 digests of "
            "u,v,w,p (halos included, global planes 0..l+1 in order) after steps_from_init steps from the initial "
            "conditions; independent of the z-slab decomposition", "steps_from_init": steps_from_init,
            "p_error": [float(e).hex() for e in errs], "p_error_last": float(errs[-1]) if len(errs) else None,
            "p_error_sha256": hashlib.sha256(errs.tobytes()).hexdigest()}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
SETUP = {}   # one-off input preparation timings of the last make_solver call


def make_solver(workload, iter_max, rank, nranks, uid, sor_variant, use_graph, halo_transport=0):
    from pixelflow_b200 import Solver
    name, (m, n, l), kw = workload_params(workload, iter_max)
    import torch
    s = Solver("ibm3_uniform", m, n, l, device=torch.cuda.current_device(), rank=rank, nranks=nranks,
               nccl_unique_id=uid, host_is_slab=True,
               sor_variant=sor_variant, use_graph=use_graph, halo_transport=halo_transport, **kw)
    if name.startswith("dragon_stl"):
        # the same dragon through the STL route: signed distance to the 67,116 triangles at every cell centre, on the GPU
        tri = wl.dragon_triangles_in_cells(np.load(os.path.join(ROOT, "tests", "golden", "stl_meshes.npz"))["dragon"], m)
        t0 = time.perf_counter()
        eps = wl.porosity_from_stl(tri, m, k_first=s.k_first, k_count=s.k_count, device=torch.cuda.current_device())
        SETUP["stl2poro_s"] = time.perf_counter() - t0
    elif name.startswith("dragon"):
        # BASELINE configs[3]: voxel model -> porosity by the GPU tanh filter (the reference: scipy, hours at 256^3)
        occ = wl.load_occupancy(os.path.join(ROOT, "tests", "golden", f"dragon_voxels_{m}.npz"))
        t0 = time.perf_counter()
        eps = wl.porosity_from_occupancy(occ, k_first=s.k_first, k_count=s.k_count, device=torch.cuda.current_device())
        SETUP["voxel2poro_s"] = time.perf_counter() - t0
    else:
        eps = wl.porous_channel(m, n, l, k_first=s.k_first, k_count=s.k_count)
    s.set_porosity(eps)
    del eps
    s.initial_conditions()
    return name, (m, n, l), s


NUMA = {}   # what bind_to_gpu_numa_node did, reported in the e2e key
NUMA_CPUS = set()
try:
    FULL_AFFINITY = os.sched_getaffinity(0)
except (AttributeError, OSError):
    FULL_AFFINITY = None


def bind_to_gpu_numa_node(local_rank):
    """Run this process on the cores of the NUMA node its GPU hangs off, so that the pinned host buffers of the e2e leg
    (first touched by this process) are local to the GPU's PCIe root: with 8 ranks the host copies otherwise cross the
    socket interconnect.  Best effort: sysfs or NVML missing -> nothing happens."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(vis.split(",")[local_rank]) if vis and vis.split(",")[local_rank].strip().isdigit() else local_rank
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        node = int(open(dev + "/numa_node").read())
        cpus = set()
        for part in open(dev + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if node >= 0 and cpus:
            os.sched_setaffinity(0, cpus)
            NUMA.update({"node": node, "cpus": len(cpus)})
            NUMA_CPUS.update(cpus)
        else:
            NUMA["skipped"] = f"the host exposes no NUMA placement for {bus} (numa_node = {node})"
    except Exception as e:   # noqa: BLE001
        NUMA["skipped"] = repr(e)[:120]


def gpu_measure(args, workload, rank, nranks, dist, uid, with_e2e, with_parity=False):
    import torch
    if NUMA_CPUS:   # (again, after a CPU-baseline leg gave the process every core back)
        os.sched_setaffinity(0, NUMA_CPUS)
    name, (m, n, l), s = make_solver(workload, args.iter_max, rank, nranks, uid, args.sor_variant, args.use_graph,
                                     args.halo_transport)
    cells = m * n * l
    K, W = args.steps, args.warmup

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if W > 0:
        s.step(W)
    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    t0 = time.perf_counter()
    errs = s.step(K)               # CUDA events on the solver's stream bracket exactly these K steps
    torch.cuda.synchronize()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    t = s.last_timing()
    times = torch.tensor([t["ms_total"], t["ms_sor"], wall * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_sor, ms_wall = (float(x) for x in times.cpu())
    variant = s.sor_variant
    res = {
        "workload": name, "dims": (m, n, l), "cells": cells, "sor_variant": variant,
        "halo_transport": {0: "none (one rank)", 1: "nccl send/recv", 2: "peer stores over NVLink (CUDA IPC), flag-barrier kernel per iteration",
                           3: "peer stores over NVLink (CUDA IPC), neighbour handshake inside the sweep kernel"}.get(
            s.halo_transport if variant in (3, 4, 6) else (1 if nranks > 1 else 0)),
        "ms_per_step": ms_total / K, "ms_sor_per_step": ms_sor / K, "ms_wall_per_step": ms_wall / K,
        "value": cells * K / (ms_total * 1e-3),
        "sweeps_per_s": K * args.iter_max / (ms_sor * 1e-3) if ms_sor > 0 else None,
        "launches": t["launches"], "clocks": clocks, "k_count": s.k_count, "setup": dict(SETUP),
    }
    # roofline of the dominant kernel (SOR half-sweep): algorithmic bytes per launch / mean launch time
    local_cells = m * n * s.k_count
    fused = variant in (3, 4, 6)          # one launch = one whole red+black iteration
    n_launch = (1 if fused else 2) * args.iter_max * K
    bytes_per_launch = ALGO_BYTES_PER_CELL_SWEEP / (1.0 if fused else 2.0) * local_cells
    res["sor_launch_ms"] = ms_sor / n_launch if n_launch else None
    res["sor_gbs"] = bytes_per_launch / (ms_sor / n_launch * 1e-3) / 1e9 if n_launch and ms_sor > 0 else None
    res["local_cells"] = local_cells
    res["sweeps_per_launch"] = 1.0 if fused else 0.5
    if with_parity:
        res["parity"] = parity_record(s, errs, l, rank, nranks, dist, W + K)
    if with_e2e:
        Ke = max(1, min(K, 10))
        shape = s.shape
        bufs = [torch.zeros(shape, dtype=torch.float64).pin_memory().numpy() for _ in range(4)]
        s.download(*bufs)
        s.step_host(1, *bufs)      # warm-up of the host path
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            s.step_host(1, *bufs)
        torch.cuda.synchronize()
        barrier()
        we = time.perf_counter() - t0
        tw = torch.tensor([we], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        we = float(tw.cpu()[0])
        nbytes = 4 * int(np.prod(shape)) * 8
        copy_s = max(we / Ke - res["ms_per_step"] * 1e-3, 1e-9)
        res["e2e"] = {"value": cells * Ke / we, "unit": "cell-updates/s", "h2d_bytes_per_step": nbytes * nranks,
                      "d2h_bytes_per_step": nbytes * nranks, "steps": Ke, "ms_per_step": we / Ke * 1e3,
                      "api": "pf_step_host (pinned host u,v,w,p in and out every step; transfers chunked along z and "
                             "overlapped with the first and last phases of the step)",
                      "exposed_copy_ms_per_step": copy_s * 1e3,
                      "pcie_gbs_per_gpu_if_serial": 2 * nbytes / copy_s / 1e9, "host_numa_binding": dict(NUMA)}
        del bufs
    s.close()
    return res


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nranks = world if world > 1 else 1
    if args.gpus != nranks and world > 1:
        args.gpus = nranks

    if args.impl == "reference":
        # CPU arm: rank 0 alone runs and prints; the others exit 0 without work
        if rank != 0:
            return 0
        # the whole workload where the translated program's static bounds hold it (256^3), else a bounded sub-block
        wname, wdims, _ = workload_params(args.workload, args.iter_max)
        value, ms, info = cpu_reference_run(args.workload, args.iter_max, args.steps, args.warmup, full=max(wdims) <= 256)
        name, dims, _ = workload_params(args.workload, args.iter_max)
        info["value"] = value
        info["unit"] = "cell-updates/s"
        print(json.dumps({
            "impl": "reference", "metric": "cell_updates_per_s", "value": value, "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": name, "solver": "ibm3_uniform", "iter_max": args.iter_max},
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return 0

    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the hot path has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(local_rank)
    dist = None
    uid = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
        from pixelflow_b200 import comm_unique_id
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, src=0)
        uid = bytes(buf.cpu().numpy().tobytes())

    res = gpu_measure(args, args.workload, rank, nranks, dist, uid, with_e2e=not args.no_e2e,
                      with_parity=not args.no_parity)

    # N = 1: the same steps again with the half-sweep kernel (variant 1, the plainest path) must leave the same hash
    if nranks == 1 and not args.no_parity and res["sor_variant"] != 1:
        a1 = argparse.Namespace(**{**vars(args), "sor_variant": 1})
        r1 = gpu_measure(a1, args.workload, 0, 1, None, None, with_e2e=False, with_parity=True)
        pa, pb = res["parity"], r1["parity"]
        res["parity"]["crosscheck"] = {
            "against": "sor_variant 1 (sor_sweep_kernel half-sweeps), same workload, same steps",
            "fields_identical": pa["fields_sha256"] == pb["fields_sha256"],
            "p_error_identical": pa["p_error_sha256"] == pb["p_error_sha256"],
            "ms_per_step": r1["ms_per_step"]}

    also = None
    if nranks == 1 and not args.no_also and WORKLOAD_ALIASES.get(args.workload, args.workload) != "s2_256":
        # the north-star single-GPU target is quoted on a 256^3 case: measure it beside the headline
        also = gpu_measure(args, "s2_256", 0, 1, None, None, with_e2e=not args.no_e2e, with_parity=not args.no_parity)

    cpu = cpu_also = None
    if rank == 0 and nranks == 1 and not args.no_cpu_baseline:
        # one protocol for every CPU number of this line and of `--impl reference`: 1 warm-up step, then timed steps
        v, ms, info = cpu_reference_run(args.workload, args.iter_max, 2, 1)
        info.update({"value": v, "unit": "cell-updates/s", "ms_per_step": ms})
        cpu = info
        if also:
            try:   # the 256^3 case fits the translated program's static bounds: the SAME configuration on the CPU
                v, ms, info = cpu_reference_run("s2_256", args.iter_max, 2, 1, full=True)
                info.update({"value": v, "unit": "cell-updates/s", "ms_per_step": ms})
                cpu_also = info
            except Exception as e:   # noqa: BLE001
                cpu_also = {"error": repr(e)}

    decks = None
    if rank == 0 and nranks == 1 and not args.no_decks:
        # outside every timed region, in its own process with a timeout: the reference's shipped decks (BASELINE
        # configs[0..2]): bit-identical to the reference's own outputs after 3 steps or not, and ms/step
        decks = run_decks_probe(0, 300)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        else:
            peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"

        def roofline(r):
            """the SOR kernel of run `r` against the HBM peak.  achieved = the bytes the kernel's data layout has to
            move per launch (its algorithmic bytes: KERNEL_BYTES_PER_CELL_SWEEP x the cells one launch sweeps) / the mean
            launch time; never above 1 unless the timing is wrong.  `survey_88` = the same time charged with SURVEY 8(d)'s
            common currency of 88 B/cell/sweep, for comparison across kernels (it may exceed 1: the fused kernels move
            fewer bytes than the model).  `traffic` = ncu dram bytes of one launch of this rank (profiles/)."""
            v = r["sor_variant"]
            bpc = KERNEL_BYTES_PER_CELL_SWEEP.get(v, ALGO_BYTES_PER_CELL_SWEEP)
            t = (r["sor_launch_ms"] or 0) * 1e-3
            units = r["local_cells"] * r["sweeps_per_launch"]
            ach = bpc * units / t / 1e9 if t > 0 else None
            ach88 = ALGO_BYTES_PER_CELL_SWEEP * units / t / 1e9 if t > 0 else None
            traffic = tsrc = None
            tp = os.path.join(ROOT, "profiles", "sor_traffic.json")
            if os.path.exists(tp):
                try:
                    tj = json.load(open(tp)).get(f"variant_{v}", {})
                    if tj.get("dram_bytes_per_cell_sweep"):
                        traffic = tj["dram_bytes_per_cell_sweep"] * units
                        tsrc = tj.get("source")
                except Exception:   # noqa: BLE001
                    pass
            out = {"bound": "hbm", "kernel": KERNEL_NAMES.get(v, "sor_sweep_kernel"), "achieved": ach, "peak": peak,
                   "unit": "GB/s", "frac": ach / peak if ach else None, "traffic": traffic,
                   "traffic_source": tsrc, "peak_source": peak_src, "kernel_bytes_per_cell_sweep": bpc,
                   "launch_ms": r["sor_launch_ms"], "cells_per_launch": r["local_cells"],
                   "survey_88": {"bytes_per_cell_sweep": ALGO_BYTES_PER_CELL_SWEEP, "achieved": ach88,
                                 "frac": ach88 / peak if ach88 else None},
                   "dram": {"achieved": traffic / t / 1e9, "frac": traffic / t / 1e9 / peak} if traffic and t > 0 else None,
                   "note": "launch time = SOR-phase CUDA-event time / launches of the sweep kernel (includes the layout "
                           "conversion and halo kernels of the solve); per rank"}
            if out["frac"] and out["frac"] > 1.0:
                out["warning"] = "frac > 1: the kernel cannot move its own bytes faster than the copy peak -- check the timing"
            return out

        m, n, l = res["dims"]
        out = {
            "metric": "cell_updates_per_s", "value": res["value"], "unit": "cell-updates/s",
            "n_gpus": nranks, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": res["workload"], "solver": "ibm3_uniform", "grid": [m, n, l],
                       "iter_max": args.iter_max, "relux_factor": 1.7, "parallelism": f"z-slab x{nranks}",
                       "sor_variant": res["sor_variant"], "halo_transport": res["halo_transport"],
                       **({"setup": res["setup"]} if res.get("setup") else {}),
                       "l2": "inputs larger than L2 (no flush needed)" if res["cells"] * 8 * 10 > 126e6 * 4 else
                             "working set comparable to L2"},
            "sor_sweeps_per_s": res["sweeps_per_s"], "ms_sor_per_step": res["ms_sor_per_step"],
            "ms_wall_per_step": res["ms_wall_per_step"],
            "roofline": roofline(res),
            "step_bytes_model": {"survey_bytes_per_cell_step": ALGO_BYTES_PER_CELL_STEP_FIXED + ALGO_BYTES_PER_CELL_SWEEP * args.iter_max,
                                 "frac_of_peak": (ALGO_BYTES_PER_CELL_STEP_FIXED + ALGO_BYTES_PER_CELL_SWEEP * args.iter_max)
                                 * res["cells"] / nranks / (res["ms_per_step"] * 1e-3) / 1e9 / peak},
            "parity": res.get("parity"),
            "cpu_baseline": cpu, "e2e": res.get("e2e"), "gpu_launches": res["launches"], "clocks": res["clocks"],
        }
        if decks is not None:
            out["decks"] = decks
        if also:
            out["also"] = {"workload": also["workload"], "sor_variant": also["sor_variant"],
                           "value": also["value"], "ms_per_step": also["ms_per_step"],
                           "sor_sweeps_per_s": also["sweeps_per_s"], "roofline": roofline(also),
                           "e2e": also.get("e2e"), "parity": also.get("parity"), "clocks": also["clocks"],
                           "cpu_baseline": cpu_also}
            if cpu_also and cpu_also.get("value"):
                out["also"]["vs_reference"] = {
                    "same_config": bool(cpu_also.get("same_config")),
                    "ratio": also["value"] / cpu_also["value"],
                    "e2e_ratio": (also["e2e"]["value"] / cpu_also["value"]) if also.get("e2e") else None}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
