"""The z-slab exchange schedule (pixelflow_b200/slab.py == csrc/pf_api.cu) on CPU, world_size 2, gloo.

Each rank runs the numpy restatement of the hot path on its own slab and exchanges ghost planes with
torch.distributed (gloo) exactly where the CUDA library calls NCCL.  The gathered result must be
BIT-IDENTICAL to the single-domain run -- this is what proves the schedule (which planes, when,
across the periodic seam or not) independent of any GPU.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle_np as onp  # noqa: E402
from pixelflow_b200.slab import neighbours, opposite_face_transfers, slab_range, step_schedule  # noqa: E402


def test_slab_range_partitions_exactly():
    for l in (4, 7, 8, 33, 512):
        for P in (1, 2, 3, 4, 8):
            if l // P < 1:
                continue
            planes = []
            for r in range(P):
                first, cnt = slab_range(l, r, P)
                planes += list(range(first, first + cnt))
            assert planes == list(range(1, l + 1))
            cnts = [slab_range(l, r, P)[1] for r in range(P)]
            assert max(cnts) - min(cnts) <= 1


def test_neighbours_and_schedule():
    assert neighbours(0, 4, True) == (3, 1) and neighbours(3, 4, True) == (2, 0)
    assert neighbours(0, 4, False) == (None, 1) and neighbours(3, 4, False) == (2, None)
    assert neighbours(0, 2, True) == (1, 1)
    uni = {e.what: e.wrap for e in step_schedule(air=False)}
    assert uni["w"] is False and uni["div"] is True and uni["u,v,w,p"] is True
    assert all(e.wrap is False for e in step_schedule(air=True))
    assert opposite_face_transfers((2, 1, 0, 0, 0, 0), 1) == [] and opposite_face_transfers((1, 0, 2, 2, 2, 2), 4) == []
    t = opposite_face_transfers((2, 1, 0, 0, 0, 0), 4)
    assert [(x[2], x[3]) for x in t] == [(0, 3), (3, 0)]


# ------------------------------------------------------------------------------------------------
def _exchange(dist, rank, nranks, a, lz, wrap):
    """a: local array [lz+2, ...]; planes 1 and lz go out, planes 0 and lz+1 come in."""
    import torch
    prev, nxt = neighbours(rank, nranks, wrap)
    reqs, bufs = [], []
    if nxt is not None:
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[lz])), nxt, tag=1))
    if prev is not None:
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[1])), prev, tag=2))
    if prev is not None:
        b = torch.empty(a[0].shape, dtype=torch.float64)
        reqs.append(dist.irecv(b, prev, tag=1))
        bufs.append((0, b))
    if nxt is not None:
        b = torch.empty(a[0].shape, dtype=torch.float64)
        reqs.append(dist.irecv(b, nxt, tag=2))
        bufs.append((lz + 1, b))
    for r in reqs:
        r.wait()
    for k, b in bufs:
        a[k] = b.numpy()


def _slab_step(dist, rank, nranks, Pg, koff, lz, e, p, u, v, w, c):
    """one time step on a slab; mirrors run_steps() in csrc/pf_api.cu"""
    Pl = onp.Params(**{**Pg.__dict__, "l": lz})
    m, n = Pg.m, Pg.n
    ex = lambda a, wrap: _exchange(dist, rank, nranks, a, lz, wrap)
    uo, vo, wo = u.copy(), v.copy(), w.copy()
    # divergence + y halo locally, z by exchange (wrap)
    zsave = c["div"][[0, lz + 1]].copy()
    onp.divergence_3d(Pl, False, uo, vo, wo, c["div"])
    c["div"][[0, lz + 1]] = zsave           # undo the local periodic-z copy of the single-domain routine
    ex(c["div"], True)
    onp.predictor_3d(Pl, uo, vo, wo, e, c["div"], u, v, w)
    ex(w, False)                            # interfaces only: the seam keeps the stale plane
    onp.matrix_3d(Pl, u, v, w, e, c)
    onp.boundary_matrix_3d_uniform(Pl, p, c)
    # SOR with global-k colouring
    s = lambda a, di=0, dj=0, dk=0: onp._sh(a, Pl, di, dj, dk)
    k, j, i = np.meshgrid(np.arange(1, lz + 1) + koff, np.arange(1, n + 1), np.arange(1, m + 1), indexing="ij")
    masks = (((i + j + k) % 2) == 0, ((i + j + k) % 2) == 1)
    om = Pg.relux_factor
    err = 0.0

    def halo():
        p[1:lz + 1, 0, 1:m + 1] = p[1:lz + 1, n, 1:m + 1]
        p[1:lz + 1, n + 1, 1:m + 1] = p[1:lz + 1, 1, 1:m + 1]
        ex(p, True)

    for _ in range(Pg.iter_max):
        for half in (0, 1):
            halo()
            po = p.copy()
            new = ((s(c["bb"]) - s(c["ae"]) * s(po, 1) - s(c["aw"]) * s(po, -1) - s(c["an"]) * s(po, 0, 1)
                    - s(c["as"]) * s(po, 0, -1) - s(c["at"]) * s(po, 0, 0, 1) - s(c["ab"]) * s(po, 0, 0, -1))
                   / s(c["ap"]) * om + s(po) * (1. - om))
            s(p)[masks[half]] = new[masks[half]]
        err = max(err, float(np.max(np.abs(s(p) - s(po)))))
    halo()
    onp.project_3d(Pl, p, u, v, w)
    # boundary: x faces + periodic y on the own planes, then whole-plane exchange (wrap)
    J, K = slice(1, n + 1), slice(1, lz + 1)
    import math
    u[K, J, 1] = Pg.inlet_velocity * math.cos(Pg.AoA / 1300. * onp.PI)
    v[K, J, 1] = Pg.inlet_velocity * math.sin(Pg.AoA / 1300. * onp.PI)
    w[K, J, 1] = 0.0
    for a in (u, v, w):
        a[K, J, 0] = a[K, J, 1]
    p[K, J, 0] = p[K, J, 2]
    for a in (u, v, w):
        a[K, J, m + 1] = a[K, J, m - 1]
    p[K, J, m + 1] = Pg.outlet_pressure
    for a in (u, v, w, p):
        a[K, 0, :] = a[K, n, :]
        a[K, n + 1, :] = a[K, 1, :]
        ex(a, True)
    return err


def _worker(rank, nranks, port, l, seed, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=nranks)
    try:
        m, n = 7, 6
        Pg = onp.Params(m=m, n=n, l=l, dx=0.01, dy=0.011, dz=0.009, dt=2e-4, xnue=1e-3, xlambda=0.1,
                        iter_max=5, relux_factor=1.7, inlet_velocity=1.0, outlet_pressure=0.2, AoA=5.0)
        rng = np.random.default_rng(seed)
        shape = (l + 2, n + 2, m + 2)
        e = np.zeros(shape)
        e[1:-1, 1:-1, 1:-1] = np.clip((rng.random((l, n, m)) - 0.2) / 0.6, 1e-6, 1.0)
        onp.porosity_halo_3d_uniform(Pg, e)
        fields = {nm: 0.1 * rng.standard_normal(shape) for nm in ("p", "u", "v", "w")}
        onp.boundary_3d_uniform(Pg, fields["p"], fields["u"], fields["v"], fields["w"])
        # single-domain reference (every rank computes it; cheap)
        ref = onp.State3D(Pg, False, e.copy(), **{k: a.copy() for k, a in fields.items()})
        ref_err = [ref.step() for _ in range(2)]
        # slab run
        first, lz = slab_range(l, rank, nranks)
        koff = first - 1
        sl = slice(koff, koff + lz + 2)
        loc = {k: a[sl].copy() for k, a in fields.items()}
        el = e[sl].copy()
        c = {nm: np.zeros(el.shape) for nm in ("ap", "ae", "aw", "an", "as", "at", "ab", "bb", "div")}
        errs = []
        for _ in range(2):
            err = _slab_step(dist, rank, nranks, Pg, koff, lz, el, loc["p"], loc["u"], loc["v"], loc["w"], c)
            t = torch.tensor([err], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            errs.append(float(t[0]))
        ok = errs == ref_err
        for nm in ("p", "u", "v", "w"):
            full = getattr(ref, nm)
            k0 = 0 if rank == 0 else 1
            k1 = lz + 1 if rank == nranks - 1 else lz
            ok = ok and np.array_equal(loc[nm][k0:k1 + 1], full[koff + k0:koff + k1 + 1])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# Air-condition on z-slabs (open chain, no seam), including the two places where the reference reads the OPPOSITE z
# face (csrc/pf_api.cu: do_rhs, pf_set_porosity): a top outlet starts its Dirichlet fold from bb(i,j,1)
# (ibm_3d_air_condition_omp_cpu.f90:702), a bottom inlet tests porosity(i,j,l) (:948).
def _air_slab_step(dist, rank, nranks, Pg, koff, lz, e, p, u, v, w, c, e_top):
    import torch
    Pl = onp.Params(**{**Pg.__dict__, "l": lz})
    m, n = Pg.m, Pg.n
    own_top, own_bottom = rank == nranks - 1, rank == 0
    ex = lambda a: _exchange(dist, rank, nranks, a, lz, False)
    uo, vo, wo = u.copy(), v.copy(), w.copy()
    onp.divergence_3d(Pl, True, uo, vo, wo, c["div"])      # every halo 0 ... except the slab interfaces:
    ex(c["div"])
    onp.predictor_3d(Pl, uo, vo, wo, e, c["div"], u, v, w)
    ex(w)
    onp.matrix_3d(Pl, u, v, w, e, c)
    bb1 = None
    if Pg.wall[0] == 2 and nranks > 1:                     # top outlet: the raw bb of global plane 1 goes to the last rank
        if rank == 0:
            dist.send(torch.from_numpy(np.ascontiguousarray(c["bb"][1])), nranks - 1, tag=7)
        elif own_top:
            t = torch.empty(c["bb"][1].shape, dtype=torch.float64)
            dist.recv(t, 0, tag=7)
            bb1 = t.numpy()
    onp.boundary_matrix_3d_air(Pl, p, e, c, own_top=own_top, own_bottom=own_bottom, bb1=bb1)
    s = lambda a, di=0, dj=0, dk=0: onp._sh(a, Pl, di, dj, dk)
    k, j, i = np.meshgrid(np.arange(1, lz + 1) + koff, np.arange(1, n + 1), np.arange(1, m + 1), indexing="ij")
    masks = (((i + j + k) % 2) == 0, ((i + j + k) % 2) == 1)
    om = Pg.relux_factor
    err = 0.0
    for _ in range(Pg.iter_max):
        for half in (0, 1):
            ex(p)                                           # no halo refresh inside the solve (:509-527), only the interfaces
            po = p.copy()
            new = ((s(c["bb"]) - s(c["ae"]) * s(po, 1) - s(c["aw"]) * s(po, -1) - s(c["an"]) * s(po, 0, 1)
                    - s(c["as"]) * s(po, 0, -1) - s(c["at"]) * s(po, 0, 0, 1) - s(c["ab"]) * s(po, 0, 0, -1))
                   / s(c["ap"]) * om + s(po) * (1. - om))
            s(p)[masks[half]] = new[masks[half]]
        err = max(err, float(np.max(np.abs(s(p) - s(po)))))
    ex(p)
    onp.project_3d(Pl, p, u, v, w)
    onp.boundary_3d_air(Pl, e, p, u, v, w, own_top=own_top, own_bottom=own_bottom, e_top=e_top)
    for a in (u, v, w, p):
        ex(a)
    return err


def _air_worker(rank, nranks, port, l, wall, seed, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=nranks)
    try:
        m, n = 7, 6
        Pg = onp.Params(m=m, n=n, l=l, dx=0.01, dy=0.011, dz=0.009, dt=2e-4, xnue=1e-3, xlambda=0.1,
                        iter_max=5, relux_factor=1.7, inlet_velocity=1.5, outlet_pressure=0.2, AoA=0.0, wall=wall)
        rng = np.random.default_rng(seed)
        shape = (l + 2, n + 2, m + 2)
        e = np.zeros(shape)
        e[1:-1, 1:-1, 1:-1] = np.clip((rng.random((l, n, m)) - 0.2) / 0.6, 1e-6, 1.0)
        e[l, 1::2, 1:-1] = 0.95          # fluid cells on the top plane: the outlet / the bottom inlet's test see both kinds
        e[1, 1:-1, 1::2] = 1.0
        onp.porosity_halo_3d_wall(Pg, e)
        fields = {nm: 0.1 * rng.standard_normal(shape) for nm in ("p", "u", "v", "w")}
        onp.boundary_3d_air(Pg, e, fields["p"], fields["u"], fields["v"], fields["w"])
        ref = onp.State3D(Pg, True, e.copy(), **{k: a.copy() for k, a in fields.items()})
        ref_err = [ref.step() for _ in range(2)]
        first, lz = slab_range(l, rank, nranks)
        koff = first - 1
        sl = slice(koff, koff + lz + 2)
        loc = {k: a[sl].copy() for k, a in fields.items()}
        el = e[sl].copy()
        # the bottom inlet's fluid test reads the porosity of global plane l: sent once from the last rank to rank 0
        e_top = None
        if wall[1] == 1 and nranks > 1:
            if rank == nranks - 1:
                dist.send(torch.from_numpy(np.ascontiguousarray(el[lz])), 0, tag=8)
            elif rank == 0:
                t = torch.empty(el[1].shape, dtype=torch.float64)
                dist.recv(t, nranks - 1, tag=8)
                e_top = t.numpy()
        c = {nm: np.zeros(el.shape) for nm in ("ap", "ae", "aw", "an", "as", "at", "ab", "bb", "div")}
        errs = []
        for _ in range(2):
            err = _air_slab_step(dist, rank, nranks, Pg, koff, lz, el, loc["p"], loc["u"], loc["v"], loc["w"], c, e_top)
            t = torch.tensor([err], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            errs.append(float(t[0]))
        ok = errs == ref_err
        for nm in ("p", "u", "v", "w"):
            full = getattr(ref, nm)
            k0 = 0 if rank == 0 else 1
            k1 = lz + 1 if rank == nranks - 1 else lz
            ok = ok and np.array_equal(loc[nm][k0:k1 + 1], full[koff + k0:koff + k1 + 1])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# The fused red+black kernels on z-slabs (csrc/pf_sor_fused.cu, pf_sor_tma.cu; sor_iterations() in pf_api.cu):
# depth-2 ghost planes, ONE launch per iteration that recomputes the red values of the two ghost planes next
# to the slab, and one two-plane hand-over to each neighbour per iteration (peer stores or an NCCL group).
def _fused_slab_sor(dist, rank, nranks, Pg, koff, lz, C, p, iters):
    """C[name], p: arrays [lz+4, n+2, m+2] holding planes k = -1 .. lz+2 (index k+1); ghosts = the neighbours' planes"""
    import torch
    m, n, om = Pg.m, Pg.n, Pg.relux_factor
    prev, nxt = (rank - 1) % nranks, (rank + 1) % nranks

    def v(a, k0, k1, di=0, dj=0, dk=0):   # interior cells of planes k0..k1, shifted
        return a[k0 + 1 + dk:k1 + 2 + dk, 1 + dj:n + 1 + dj, 1 + di:m + 1 + di]

    def mask(k0, k1, parity):
        k, j, i = np.meshgrid(np.arange(k0, k1 + 1) + koff, np.arange(1, n + 1), np.arange(1, m + 1), indexing="ij")
        return ((i + j + k) % 2) == parity

    def update(k0, k1, parity):
        new = ((v(C["bb"], k0, k1) - v(C["ae"], k0, k1) * v(p, k0, k1, 1) - v(C["aw"], k0, k1) * v(p, k0, k1, -1)
                - v(C["an"], k0, k1) * v(p, k0, k1, 0, 1) - v(C["as"], k0, k1) * v(p, k0, k1, 0, -1)
                - v(C["at"], k0, k1) * v(p, k0, k1, 0, 0, 1) - v(C["ab"], k0, k1) * v(p, k0, k1, 0, 0, -1))
               / v(C["ap"], k0, k1) * om + v(p, k0, k1) * (1. - om))
        mk = mask(k0, k1, parity)
        old = v(p, k0, k1)[mk].copy()
        v(p, k0, k1)[mk] = new[mk]
        # the row images the kernels store with every cell
        p[k0 + 1:k1 + 2, 0, 1:m + 1] = p[k0 + 1:k1 + 2, n, 1:m + 1]
        p[k0 + 1:k1 + 2, n + 1, 1:m + 1] = p[k0 + 1:k1 + 2, 1, 1:m + 1]
        return float(np.max(np.abs(v(p, k0, k1)[mk] - old))) if mk.any() else 0.0

    def hand_over():
        lo = torch.from_numpy(np.ascontiguousarray(p[2:4]))            # planes 1, 2      -> prev's lz+1, lz+2
        hi = torch.from_numpy(np.ascontiguousarray(p[lz:lz + 2]))      # planes lz-1, lz  -> next's -1, 0
        from_prev, from_next = torch.empty_like(hi), torch.empty_like(lo)
        reqs = [dist.isend(hi, nxt, tag=1), dist.isend(lo, prev, tag=2),
                dist.irecv(from_prev, prev, tag=1), dist.irecv(from_next, nxt, tag=2)]
        for r in reqs:
            r.wait()
        p[0:2] = from_prev.numpy()
        p[lz + 2:lz + 4] = from_next.numpy()

    err = 0.0
    for _ in range(iters):
        update(0, lz + 1, 0)                  # red, the slab and one ghost plane each side (from old black, depth 2)
        err = max(err, update(1, lz, 1))      # black, the slab
        hand_over()
    return err


def _fused_worker(rank, nranks, port, l, seed, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=nranks)
    try:
        m, n, iters = 9, 6, 5
        Pg = onp.Params(m=m, n=n, l=l, dx=0.01, dy=0.011, dz=0.009, dt=2e-4, xnue=1e-3, xlambda=0.1,
                        iter_max=iters, relux_factor=1.7, inlet_velocity=1.0, outlet_pressure=0.2, AoA=5.0)
        rng = np.random.default_rng(seed)
        shape = (l + 2, n + 2, m + 2)
        e = np.zeros(shape)
        e[1:-1, 1:-1, 1:-1] = np.clip((rng.random((l, n, m)) - 0.2) / 0.6, 1e-6, 1.0)
        onp.porosity_halo_3d_uniform(Pg, e)
        f = {nm: 0.1 * rng.standard_normal(shape) for nm in ("p", "u", "v", "w")}
        onp.boundary_3d_uniform(Pg, f["p"], f["u"], f["v"], f["w"])
        c = {nm: np.zeros(shape) for nm in ("ap", "ae", "aw", "an", "as", "at", "ab", "bb")}
        onp.matrix_3d(Pg, f["u"], f["v"], f["w"], e, c)
        onp.boundary_matrix_3d_uniform(Pg, f["p"], c)
        first, lz = slab_range(l, rank, nranks)
        koff = first - 1
        planes = (np.arange(-1, lz + 3) + koff - 1) % l + 1          # global plane of local k = -1 .. lz+2
        C = {nm: a[planes].copy() for nm, a in c.items()}
        p_loc = f["p"][planes].copy()
        p_loc[:, 0, 1:m + 1] = p_loc[:, n, 1:m + 1]
        p_loc[:, n + 1, 1:m + 1] = p_loc[:, 1, 1:m + 1]
        p_ref = f["p"].copy()
        err_ref = onp.sor_3d(Pg, True, iters, p_ref, c)
        err = _fused_slab_sor(dist, rank, nranks, Pg, koff, lz, C, p_loc, iters)
        t = torch.tensor([err], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = float(t[0]) == err_ref
        ok = ok and np.array_equal(p_loc[2:lz + 2, 1:n + 1, 1:m + 1], p_ref[first:first + lz, 1:n + 1, 1:m + 1])
        # the ghost planes the next phase (projection) reads: one plane each side, as the neighbours left them
        ok = ok and np.array_equal(p_loc[[1, lz + 2], 1:n + 1, 1:m + 1],
                                   p_ref[[(koff - 1) % l + 1, (koff + lz) % l + 1], 1:n + 1, 1:m + 1])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("l", [8, 6, 7])   # even koff, odd koff, odd l (colour flips across the seam)
def test_slab_schedule_matches_single_domain(l):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, l, 42 + l, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)], results


@pytest.mark.parametrize("l", [8, 10, 12])   # 4, 5 (odd colour offset) and 6 planes per rank
def test_fused_slab_schedule_matches_single_domain_sor(l):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fused_worker, args=(r, 2, port, l, 7 + l, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)], results


# wall codes: top, bottom, east, west, south, north (0 wall, 1 inlet, 2 outlet)
@pytest.mark.parametrize("l,wall", [(8, (1, 0, 0, 0, 2, 0)), (7, (2, 1, 2, 1, 1, 2)), (6, (2, 1, 0, 0, 2, 1)), (8, (0, 2, 2, 1, 1, 2))])
def test_air_condition_slab_schedule_matches_single_domain(l, wall):
    """open chain of slabs; (2, 1, ...) = top outlet + bottom inlet, where planes of the opposite z face travel"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_air_worker, args=(r, 2, port, l, wall, 99 + l, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)], results
