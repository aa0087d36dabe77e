"""Host-side z-slab decomposition logic (mirrors `build()` in csrc/pf_api.cu) and the halo-exchange
schedule of one time step, stated as data so that it can be tested without a GPU.

The reference is a single-address-space code; the decomposition is this framework's addition.  The
schedule below is what keeps a slab run bit-identical to the single-domain run, including the
reference's stale-halo behaviour.
"""
from __future__ import annotations

from dataclasses import dataclass


def slab_range(l: int, rank: int, nranks: int):
    """(k_first, k_count): 1-based first global plane and number of planes owned by `rank`.
    Remainder planes go to the lowest ranks."""
    if not (0 <= rank < nranks):
        raise ValueError("bad rank")
    base, rem = divmod(l, nranks)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem) + 1
    return first, count


def neighbours(rank: int, nranks: int, wrap: bool):
    """(prev, next) ranks of the z ring (wrap=True: periodic z) or open chain (None at the ends)."""
    prev = (rank - 1) % nranks if (wrap or rank > 0) else None
    nxt = (rank + 1) % nranks if (wrap or rank < nranks - 1) else None
    return prev, nxt


@dataclass(frozen=True)
class Exchange:
    """one plane exchange: `what` array, after `phase`; wrap=True crosses the periodic seam"""
    phase: str
    what: str
    wrap: bool
    note: str


def step_schedule(air: bool):
    """Exchanges of one time step, in order, for nranks > 1 (pf_api.cu: do_divergence, do_predictor,
    sor_refresh, do_boundary)."""
    periodic = not air
    return [
        Exchange("divergence", "div", periodic,
                 "div(i,j,k+-1) of the w predictor; periodic seam = the reference's div(i,j,0)=div(i,j,l) (:215-222)"),
        Exchange("predictor", "w", False,
                 "bb reads w(i,j,k+-1) (:408-409): fresh across slab interfaces, STALE across the periodic seam "
                 "(the predictor writes only k=1..l, the halo keeps the previous step's value)"),
        Exchange("sor_half_sweep", "p[colour just updated]", periodic,
                 "before every half-sweep (2*iter_max per step) and once after the last (:473-480,:530-537,:598-605); "
                 "across the seam the colour flips iff l is odd"),
        Exchange("boundary", "u,v,w,p", periodic,
                 "after the x faces and periodic-y rows of the own planes; the seam copy is the reference's "
                 "u(i,j,0)=u(i,j,l) (:735-748)"),
    ]


def opposite_face_transfers(wall, nranks: int):
    """Air-condition on z-slabs: the two places where the reference reads the OPPOSITE z face, as (phase, what, src
    rank, dst rank, how often).  wall = (top, bottom, east, west, south, north), 0 wall / 1 inlet / 2 outlet.
    A top outlet starts its Dirichlet fold from bb(i,j,1) (ibm_3d_air_condition_omp_cpu.f90:702); a bottom inlet tests
    porosity(i,j,l) (:948).  pf_api.cu: do_rhs, pf_set_porosity."""
    out = []
    if nranks > 1 and wall[0] == 2:
        out.append(("poisson_source", "raw bb of global plane 1", 0, nranks - 1, "every step"))
    if nranks > 1 and wall[1] == 1:
        out.append(("set_porosity", "porosity of global plane l", nranks - 1, 0, "once"))
    return out


def tma_schedule(m: int, n: int, lz: int, sms: int = 148, tw: int = 32, tr: int = 16):
    """Mirror of pf_tma_schedule() in csrc/pf_sor_tma.cu: the z-chunk schedule of the TMA sweep kernel for a rank that
    owns `lz` planes of an m x n grid on a GPU with `sms` SMs.  Returns (tiles, tA, nzA, nzB, blocks): the first tA
    tiles (x-fastest order) are cut into nzA z-chunks each, the others into nzB; a block costs (planes of its chunk + 3)
    steps and blocks reach SMs in launch order as SMs become free; the schedule with the shortest simulated makespan
    wins.  Stated in Python so that the host logic can be tested without a GPU (tests/test_host_logic.py); the kernel's
    own grid sizes in profiles/r02_final_sor_tma_ncu_raw_*.csv (launch__grid_size 740 and 295) are this function's
    answers for 1024x512x512 and 256^3."""
    import heapq
    cols = ((m + 1) >> 1) + 2
    tiles = ((cols + tw - 3) // (tw - 2)) * ((n + tr - 3) // (tr - 2))

    def norm(nz):
        cz = -(-lz // nz)
        return -(-lz // cz)

    def makespan(tA, nzA, nzB):
        heap = [0.0] * sms
        for nblocks, cost in ((tA * nzA, -(-lz // nzA) + 3.0), ((tiles - tA) * nzB, -(-lz // nzB) + 3.0)):
            for _ in range(nblocks):
                heapq.heapreplace(heap, heap[0] + cost)
        return max(heap)

    best = None
    nzmax = max(1, min(16, lz // 8))
    for nzA in range(1, nzmax + 1):
        if norm(nzA) != nzA:
            continue
        for nzB in range(nzA, nzmax + 1):
            if norm(nzB) != nzB:
                continue
            for tA in (tiles, tiles * nzA // sms * sms // nzA):
                if tA < 0 or tA > tiles or (nzB == nzA) != (tA == tiles):
                    continue
                c = makespan(tA, nzA, nzB)
                if best is None or c < best[0] - 1e-9:
                    best = (c, tA, nzA, nzB)
    _, tA, nzA, nzB = best
    return tiles, tA, nzA, nzB, tA * nzA + (tiles - tA) * nzB


def tma_block_chunk(block: int, lz: int, tA: int, nzA: int, nzB: int):
    """Mirror of the kernel's block -> (tile, first plane, last plane) decoding (sor_tma_kernel, csrc/pf_sor_tma.cu)"""
    nA = tA * nzA
    if block < nA:
        nzc, tile = nzA, block // nzA
        zc = block - tile * nzA
    else:
        b = block - nA
        nzc, t = nzB, b // nzB
        zc, tile = b - t * nzB, tA + t
    czp = -(-lz // nzc)
    k0 = zc * czp + 1
    return tile, k0, min(k0 + czp - 1, lz)
